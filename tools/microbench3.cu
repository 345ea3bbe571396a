// launch -> host-visible completion latency of small kernels (mapped pinned flag, as Team::wait_result polls it)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
struct Big { char b[3300]; };
__global__ void k_flag(volatile unsigned long long *flag, unsigned long long v) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); *flag = v; }
}
__global__ void k_flag_big(const __grid_constant__ Big p, volatile unsigned long long *flag, unsigned long long v) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); *flag = v + p.b[0]; }
}
__global__ void k_flag_smem(volatile unsigned long long *flag, unsigned long long v) {
  extern __shared__ unsigned char sm[];
  if (threadIdx.x == 0) sm[0] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); *flag = v; }
}
template <class F> double run(F launch, volatile unsigned long long *h, int iters) {
  double tot = 0;
  for (int i = 1; i <= iters + 20; ++i) {
    auto t0 = std::chrono::steady_clock::now();
    launch((unsigned long long)i);
    while (*h != (unsigned long long)i) {}
    if (i > 20) tot += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  return tot / iters * 1e6;
}
int main() {
  unsigned long long *h, *d;
  cudaHostAlloc(&h, 64, cudaHostAllocMapped);
  cudaHostGetDevicePointer(&d, h, 0);
  cudaStream_t s; cudaStreamCreate(&s);
  Big big{};
  const int smem = 216 * 1024;
  cudaFuncSetAttribute(k_flag_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  *h = 0; printf("regular 148x256            : %.1f us\n", run([&](unsigned long long v) { k_flag<<<148, 256, 0, s>>>(d, v); }, h, 300));
  *h = 0; printf("regular 10x256             : %.1f us\n", run([&](unsigned long long v) { k_flag<<<10, 256, 0, s>>>(d, v); }, h, 300));
  *h = 0; printf("regular 148x256, 3.3KB arg : %.1f us\n", run([&](unsigned long long v) { k_flag_big<<<148, 256, 0, s>>>(big, d, v); }, h, 300));
  *h = 0; printf("regular 148x256, 216KB smem: %.1f us\n", run([&](unsigned long long v) { k_flag_smem<<<148, 256, smem, s>>>(d, v); }, h, 300));
  *h = 0; printf("cooperative 148x256        : %.1f us\n", run([&](unsigned long long v) { void *a[] = {&d, &v}; cudaLaunchCooperativeKernel((void *)k_flag, dim3(148), dim3(256), a, 0, s); }, h, 300));
  *h = 0; printf("cooperative 148, 216KB smem: %.1f us\n", run([&](unsigned long long v) { void *a[] = {&d, &v}; cudaLaunchCooperativeKernel((void *)k_flag_smem, dim3(148), dim3(256), a, smem, s); }, h, 300));
  // alternate smem configs (carve-out switch between kernels)
  *h = 0; printf("alternating 0 / 216KB smem : %.1f us per pair\n", run([&](unsigned long long v) { k_flag_smem<<<148, 256, smem, s>>>(d, 0); k_flag<<<148, 256, 0, s>>>(d, v); }, h, 300));
  return 0;
}
