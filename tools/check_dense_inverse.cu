// Stand-alone check + timing of dpgo::spd_inverse (dpgo_ros_b200/csrc/dense_inverse.cu) on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dpgo_ros_b200/csrc tools/check_dense_inverse.cu \
//        dpgo_ros_b200/csrc/build/dense_inverse.o -o /tmp/check_dense_inverse && /tmp/check_dense_inverse
// Residual |A (P v) - v| / |v| over random v for strictly diagonally dominant random symmetric A; N^3 / 2 FMA rate.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "kernels.h"

int main(int argc, char **argv) {
  std::vector<int> sizes = {32, 64, 96, 160, 448, 1248, 2528, 5024};
  if (argc > 1) {
    sizes.clear();
    for (int i = 1; i < argc; ++i) sizes.push_back(atoi(argv[i]));
  }
  int bad = 0;
  for (int N : sizes) {
    const size_t NN = (size_t)N * N;
    std::vector<double> A(NN);
    std::mt19937_64 rng(N);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    for (int j = 0; j < N; ++j)
      for (int i = j; i < N; ++i) {
        const double v = (i == j) ? 0.6 * N + 1.0 : U(rng);
        A[(size_t)j * N + i] = v;
        A[(size_t)i * N + j] = v;
      }
    double *dA, *dW;
    int *dinfo;
    cudaMalloc(&dA, NN * 8);
    cudaMalloc(&dW, NN * 8);
    cudaMalloc(&dinfo, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    const int reps = getenv("REPS") ? atoi(getenv("REPS")) : 3;
    for (int rep = 0; rep < reps; ++rep) {
      cudaMemcpy(dA, A.data(), NN * 8, cudaMemcpyHostToDevice);
      cudaEventRecord(e0);
      cudaError_t err = dpgo::spd_inverse(dA, dW, N, dinfo, 0);
      cudaEventRecord(e1);
      cudaError_t e2 = cudaDeviceSynchronize();
      if (err != cudaSuccess || e2 != cudaSuccess) {
        printf("N=%d CUDA error %s / %s\n", N, cudaGetErrorString(err), cudaGetErrorString(e2));
        return 2;
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    int info = -1;
    cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    std::vector<double> P(NN);
    cudaMemcpy(P.data(), dA, NN * 8, cudaMemcpyDeviceToHost);
    double worst = 0, asym = 0;
    for (int t = 0; t < 3; ++t) {
      std::vector<double> v(N), pv(N, 0.0), apv(N, 0.0);
      for (auto &x : v) x = U(rng);
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) pv[i] += P[(size_t)j * N + i] * v[j];
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) apv[i] += A[(size_t)j * N + i] * pv[j];
      double num = 0, den = 0;
      for (int i = 0; i < N; ++i) {
        num += (apv[i] - v[i]) * (apv[i] - v[i]);
        den += v[i] * v[i];
      }
      worst = std::fmax(worst, std::sqrt(num / den));
    }
    for (int j = 0; j < N; j += 7)
      for (int i = 0; i < N; i += 3) asym = std::fmax(asym, std::fabs(P[(size_t)j * N + i] - P[(size_t)i * N + j]));
    const double fma = 0.5 * (double)N * N * N;
    printf("N=%5d  info=%d  residual=%.2e  asym=%.1e  %.3f ms  %.2f TFMA/s  launches so far %lld\n", N, info, worst, asym,
           best, fma / (best * 1e-3) / 1e12, dpgo::dense_inverse_launch_count());
    if (!(worst < 1e-10) || info != 0 || asym != 0.0) ++bad;
    cudaFree(dA);
    cudaFree(dW);
    cudaFree(dinfo);
  }
  // a matrix that is not positive definite must be reported
  {
    const int N = 96;
    std::vector<double> A((size_t)N * N, 0.0);
    for (int i = 0; i < N; ++i) A[(size_t)i * N + i] = (i == 70) ? -1.0 : 2.0;
    double *dA, *dW;
    int *dinfo;
    cudaMalloc(&dA, A.size() * 8);
    cudaMalloc(&dW, A.size() * 8);
    cudaMalloc(&dinfo, 4);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    dpgo::spd_inverse(dA, dW, N, dinfo, 0);
    int info = 0;
    cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    printf("indefinite matrix: info=%d (expected 71)\n", info);
    if (info != 71) ++bad;
  }
  printf(bad ? "FAILED (%d)\n" : "ok\n", bad);
  return bad ? 1 : 0;
}
