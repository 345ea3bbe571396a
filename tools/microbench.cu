// B200 micro-latencies that the RBCD phases are built from (diagnostics only).
#include <cstdio>
#include <cuda_runtime.h>
#include "../dpgo_ros_b200/csrc/device.cuh"
using namespace dpgo;

__global__ void k(long long *out, double *buf, int *chain, double seed) {
  const int a = threadIdx.x & 7;
  double m[4] = {(a == 0) + 0.01 * seed * a, (a == 1) - 0.02 * seed, (a == 2) + 0.015 * seed, 1.0};
  long long t0, t1;
  // (0) DFMA dependent chain x 128
  double x = seed;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) x = fma(x, 1.0000001, 0.5);
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  buf[threadIdx.x] = x;
  // (1) DADD dependent chain x 128
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i) x = x + 1.25;
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;
  buf[threadIdx.x] += x;
  // (2) one gsum8
  t0 = clock64();
  double s = gsum8(m[0] * m[1]);
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = t1 - t0;
  buf[threadIdx.x] += s;
  // (3) six independent gsum8 (Gram matrix)
  t0 = clock64();
  Sym3 A;
  A.a00 = gsum8(m[0] * m[0]); A.a01 = gsum8(m[0] * m[1]); A.a02 = gsum8(m[0] * m[2]);
  A.a11 = gsum8(m[1] * m[1]); A.a12 = gsum8(m[1] * m[2]); A.a22 = gsum8(m[2] * m[2]);
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = t1 - t0;
  // (4) sym3_invsqrt near identity
  t0 = clock64();
  Sym3 B = sym3_invsqrt(A);
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = t1 - t0;
  buf[threadIdx.x] += B.a00 + B.a12;
  // (5) full stiefel_project_row
  t0 = clock64();
  stiefel_project_row(m);
  t1 = clock64();
  if (threadIdx.x == 0) out[5] = t1 - t0;
  buf[threadIdx.x] += m[0] + m[1] + m[2];
  // (6) qf_row
  m[0] += 0.01; m[1] -= 0.02;
  t0 = clock64();
  qf_row(m);
  t1 = clock64();
  if (threadIdx.x == 0) out[6] = t1 - t0;
  buf[threadIdx.x] += m[0] + m[1] + m[2];
  // (7) tangent_project_row
  double z[4] = {seed, 2 * seed, -seed, 1};
  t0 = clock64();
  tangent_project_row(m, z);
  t1 = clock64();
  if (threadIdx.x == 0) out[7] = t1 - t0;
  buf[threadIdx.x] += z[0] + z[1] + z[2];
  // (8) dependent global loads (L2 hits: chain was written by the host; bypass L1 with ld.cg)
  int p = threadIdx.x & 31;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 16; ++i) p = __ldcg(chain + p);
  t1 = clock64();
  if (threadIdx.x == 0) out[8] = (t1 - t0) / 16;
  buf[threadIdx.x] += p;
  // (9) dependent global loads with default caching (L1 hits after first)
  p = threadIdx.x & 31;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 16; ++i) p = chain[p];
  t1 = clock64();
  if (threadIdx.x == 0) out[9] = (t1 - t0) / 16;
  buf[threadIdx.x] += p;
  // (10) jacobi path
  Sym3 C = {2.0 + seed, 0.3, -0.2, 1.5, 0.1, 0.7};
  t0 = clock64();
  Sym3 D = sym3_invsqrt(C);
  t1 = clock64();
  if (threadIdx.x == 0) out[10] = t1 - t0;
  buf[threadIdx.x] += D.a00 + D.a12;
  // (11) 64-bit shuffle latency chain x 16
  double sv = seed + threadIdx.x;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 16; ++i) sv = __shfl_xor_sync(0xffffffffu, sv, 1) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) out[11] = (t1 - t0) / 16;
  buf[threadIdx.x] += sv;
}

int main() {
  long long *out; double *buf; int *chain;
  cudaMalloc(&out, 16 * 8); cudaMalloc(&buf, 1024 * 8); cudaMalloc(&chain, 4096 * 4);
  int h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (i * 37 + 11) % 4096;
  cudaMemcpy(chain, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 3; ++rep) k<<<1, 32>>>(out, buf, chain, 0.001 * (rep + 1));
  cudaDeviceSynchronize();
  long long ho[16];
  cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
  const char *names[] = {"dfma_chain128", "dadd_chain128", "gsum8_x1", "gsum8_x6", "invsqrt_ns", "stiefel_project", "qf_row",
                         "tangent_project", "ldcg_L2_latency", "ld_default_latency", "invsqrt_jacobi", "shfl64+dadd"};
  for (int i = 0; i < 12; ++i) printf("%-20s %lld cycles\n", names[i], ho[i]);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
