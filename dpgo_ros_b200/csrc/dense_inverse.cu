// Dense SPD inverse on the device: builds (Q + lambda I)^-1 for the
// preconditioner (a6).  The reference factors Q + lambda I with CHOLMOD and
// runs two sparse triangular solves per application -- sequential work that a
// GPU cannot hide.  Here the factorisation is paid once per Q change as a
// blocked dense Cholesky + triangular inverse + L^-T L^-1 product, and each
// application becomes one streaming (r x 4n)(4n x 4n) product (phases.cuh).
//
// Blocked with 32 x 32 tiles; N is padded to a multiple of 32 with identity.
#include "kernels.h"

namespace dpgo {

constexpr int NB = 32;

// C(32x32) += A(32x32) * B(32x32) with both already in shared memory as
// As[row][k], Bs[k][col]; thread (tx, ty) owns rows tx, cols ty + 8 m.
__device__ __forceinline__ void tile_mma(const double (*As)[NB + 1], const double (*Bs)[NB + 1], double (&c)[4]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 8
  for (int k = 0; k < NB; ++k) {
    const double a = As[tx][k];
#pragma unroll
    for (int m = 0; m < 4; ++m) c[m] = fma(a, Bs[k][ty + 8 * m], c[m]);
  }
}

// load a 32x32 tile at (row0, col0) of column-major M (ld) into S[row][col] or transposed S[col][row]
__device__ __forceinline__ void tile_load(const double *M, size_t ld, int row0, int col0, double (*S)[NB + 1],
                                          bool transpose) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int c = ty + 8 * m;
    const double v = M[(size_t)(col0 + c) * ld + row0 + tx];
    if (transpose)
      S[c][tx] = v;
    else
      S[tx][c] = v;
  }
}

__global__ void k_scatter_blocks(double *P, size_t ld, const int *rowptr, const int *col, const double *val, int n,
                                 double lambda, int npad) {
  // block-CSR by output pose j (block column j): entry e = (row block col[e], col block j)
  const int j = blockIdx.x;
  if (j < n) {
    for (int e = rowptr[j] + (threadIdx.x >> 4); e < rowptr[j + 1]; e += blockDim.x >> 4) {
      const int i = col[e];
      const int q = threadIdx.x & 15;  // element of the 4x4 block, column-major
      double v = val[(size_t)e * 16 + q];
      const int rr = 4 * i + (q & 3), cc = 4 * j + (q >> 2);
      if (rr == cc) v += lambda;
      P[(size_t)cc * ld + rr] = v;
    }
  } else {
    // identity on the padding
    const int d = 4 * n + (j - n) * blockDim.x + threadIdx.x;
    if (d < npad) P[(size_t)d * ld + d] = 1.0;
  }
}

// factor the diagonal block k: L_kk (lower, written back, upper zeroed) and its inverse
__global__ void k_potrf_diag(double *A, size_t ld, int k, double *dinv, int *info) {
  __shared__ double s[NB][NB + 1];
  __shared__ double li[NB][NB + 1];
  const int t = threadIdx.x;  // 32 threads
  const int o = k * NB;
  for (int c = 0; c < NB; ++c) s[t][c] = A[(size_t)(o + c) * ld + o + t];
  __syncwarp();
  for (int j = 0; j < NB; ++j) {
    if (t == j) {
      const double d = s[j][j];
      if (!(d > 0.0)) atomicExch(info, k * NB + j + 1);
      s[j][j] = sqrt(d);
    }
    __syncwarp();
    if (t > j) s[t][j] /= s[j][j];
    __syncwarp();
    if (t > j) {
      const double l = s[t][j];
      for (int c = j + 1; c <= t; ++c) s[t][c] -= l * s[c][j];
    }
    __syncwarp();
  }
  // inverse: thread t solves L x = e_t
  for (int i = 0; i < NB; ++i) {
    double v = (i == t) ? 1.0 : 0.0;
    for (int q = t; q < i; ++q) v -= s[i][q] * li[q][t];
    li[i][t] = (i >= t) ? v / s[i][i] : 0.0;
  }
  __syncwarp();
  for (int c = 0; c < NB; ++c) {
    A[(size_t)(o + c) * ld + o + t] = (t >= c) ? s[t][c] : 0.0;
    dinv[(size_t)k * NB * NB + c * NB + t] = li[t][c];
  }
}

// A_ik <- A_ik * Linv_kk^T  for row blocks i > k
__global__ void __launch_bounds__(256) k_trsm_panel(double *A, size_t ld, int k, const double *dinv) {
  __shared__ double As[NB][NB + 1], Bs[NB][NB + 1];
  const int i = k + 1 + blockIdx.x;
  tile_load(A, ld, i * NB, k * NB, As, false);
  tile_load(dinv + (size_t)k * NB * NB, NB, 0, 0, Bs, true);  // Bs[kk][c] = Linv[c][kk]
  __syncthreads();
  double c[4] = {0, 0, 0, 0};
  tile_mma(As, Bs, c);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < 4; ++m) A[(size_t)(k * NB + ty + 8 * m) * ld + i * NB + tx] = c[m];
}

// trailing update: A_ij -= A_ik A_jk^T for i >= j > k
__global__ void __launch_bounds__(256) k_syrk_update(double *A, size_t ld, int k) {
  const int i = k + 1 + blockIdx.x, j = k + 1 + blockIdx.y;
  if (j > i) return;
  __shared__ double As[NB][NB + 1], Bs[NB][NB + 1];
  tile_load(A, ld, i * NB, k * NB, As, false);
  tile_load(A, ld, j * NB, k * NB, Bs, true);  // Bs[kk][c] = A_jk[c][kk]
  __syncthreads();
  double c[4] = {0, 0, 0, 0};
  tile_mma(As, Bs, c);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < 4; ++m) A[(size_t)(j * NB + ty + 8 * m) * ld + i * NB + tx] -= c[m];
}

// row block i of W = L^-1:  W_ii = Linv_ii;  W_ij = -Linv_ii * sum_{k=j}^{i-1} L_ik W_kj
__global__ void __launch_bounds__(256) k_trtri_row(const double *L, double *W, size_t ld, int i, const double *dinv) {
  const int j = blockIdx.x;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __shared__ double As[NB][NB + 1], Bs[NB][NB + 1];
  if (j == i) {
#pragma unroll
    for (int m = 0; m < 4; ++m)
      W[(size_t)(i * NB + ty + 8 * m) * ld + i * NB + tx] = dinv[(size_t)i * NB * NB + (ty + 8 * m) * NB + tx];
    return;
  }
  double c[4] = {0, 0, 0, 0};
  for (int k = j; k < i; ++k) {
    __syncthreads();
    tile_load(L, ld, i * NB, k * NB, As, false);
    tile_load(W, ld, k * NB, j * NB, Bs, false);
    __syncthreads();
    tile_mma(As, Bs, c);
  }
  __syncthreads();
  // T = sum (32x32) -> As ; result = -Linv_ii * T
#pragma unroll
  for (int m = 0; m < 4; ++m) Bs[tx][ty + 8 * m] = c[m];
  tile_load(dinv + (size_t)i * NB * NB, NB, 0, 0, As, false);
  __syncthreads();
  double d[4] = {0, 0, 0, 0};
  tile_mma(As, Bs, d);
#pragma unroll
  for (int m = 0; m < 4; ++m) W[(size_t)(j * NB + ty + 8 * m) * ld + i * NB + tx] = -d[m];
}

// out = W^T W (W lower triangular), tile (i, j) with i >= j, mirrored
__global__ void __launch_bounds__(256) k_lauum(const double *W, double *out, size_t ld, int nb) {
  const int i = blockIdx.x, j = blockIdx.y;
  if (j > i) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __shared__ double As[NB][NB + 1], Bs[NB][NB + 1];
  double c[4] = {0, 0, 0, 0};
  for (int k = i; k < nb; ++k) {
    __syncthreads();
    tile_load(W, ld, k * NB, i * NB, As, true);   // As[r][kk] = W_ki[kk][r]
    tile_load(W, ld, k * NB, j * NB, Bs, false);  // Bs[kk][c] = W_kj[kk][c]
    __syncthreads();
    tile_mma(As, Bs, c);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int rr = i * NB + tx, cc = j * NB + ty + 8 * m;
    out[(size_t)cc * ld + rr] = c[m];
    out[(size_t)rr * ld + cc] = c[m];
  }
}


static long long g_inv_launches = 0;
long long dense_inverse_launch_count() { return g_inv_launches; }

cudaError_t launch_scatter_blocks(double *P, size_t ld, const int *rowptr, const int *col, const double *val,
                                  int n, double lambda, int npad, cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(P, 0, sizeof(double) * ld * ld, s);
  if (err != cudaSuccess) return err;
  const int extra = (npad - 4 * n + 63) / 64;
  ++g_inv_launches;
  k_scatter_blocks<<<n + extra, 64, 0, s>>>(P, ld, rowptr, col, val, n, lambda, npad);
  return cudaGetLastError();
}

cudaError_t spd_inverse(double *A, double *work, double *dinv, int N, int *d_info, cudaStream_t s) {
  const int nb = N / NB;
  const size_t ld = N;
  cudaError_t err = cudaMemsetAsync(d_info, 0, sizeof(int), s);
  if (err != cudaSuccess) return err;
  for (int k = 0; k < nb; ++k) {
    k_potrf_diag<<<1, 32, 0, s>>>(A, ld, k, dinv, d_info);
    ++g_inv_launches;
    const int rest = nb - k - 1;
    if (rest > 0) {
      k_trsm_panel<<<rest, 256, 0, s>>>(A, ld, k, dinv);
      k_syrk_update<<<dim3(rest, rest), 256, 0, s>>>(A, ld, k);
      g_inv_launches += 2;
    }
  }
  err = cudaMemsetAsync(work, 0, sizeof(double) * ld * ld, s);
  if (err != cudaSuccess) return err;
  for (int i = 0; i < nb; ++i) {
    k_trtri_row<<<i + 1, 256, 0, s>>>(A, work, ld, i, dinv);
    ++g_inv_launches;
  }
  k_lauum<<<dim3(nb, nb), 256, 0, s>>>(work, A, ld, nb);
  ++g_inv_launches;
  return cudaGetLastError();
}

}  // namespace dpgo
