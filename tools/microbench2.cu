// FP64 throughput per SM on B200: DFMA (CUDA cores) vs DMMA m8n8k4 (tensor cores).  Diagnostics only.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, long long *cyc, double s) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = s + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], 1.0000001, 0.5);
  }
  __syncthreads();
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dmma(double *out, long long *cyc, double s) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = s;
  double a = s + threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  __syncthreads();
  long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
  for (int threads : {32, 128, 256, 512, 1024}) {
    for (int grid : {1, 148}) {
      long long h[148];
      k_dfma<<<grid, threads>>>(out, cyc, 0.5); cudaDeviceSynchronize();
      k_dfma<<<grid, threads>>>(out, cyc, 0.5); cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
      double fma_per_clk = 256.0 * 16 * threads / h[0];
      k_dmma<<<grid, threads>>>(out, cyc, 0.5); cudaDeviceSynchronize();
      k_dmma<<<grid, threads>>>(out, cyc, 0.5); cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
      double mma_fma_per_clk = 256.0 * 8 * 256 * (threads / 32) / h[0];
      printf("threads %4d grid %3d: DFMA %.1f FMA/clk/SM   DMMA %.1f FMA/clk/SM\n", threads, grid, fma_per_clk, mma_fma_per_clk);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
