// Dense preconditioner application of LARGE agents with HALF the HBM traffic (a6, the HBM-bound regime: BASELINE
// config 5, 50 016 x 50 016 doubles = 20 GB per agent).
//
//   Z = V P,   P = (Q + lambda I)^-1 symmetric,  V: R x N (row-major copy VT[a][q])
//
// P is stored in full (dense_inverse.cu mirrors it) but only its LOWER triangle is read.  It is cut into tiles of
// kPanel columns x kRows rows; a tile strictly below the diagonal serves BOTH products its entries take part in:
//   (a)  Z[:, c] += sum_q V[:, q] P[q, c]      for the tile's columns c      (reduction over the rows: thread-strided
//                                                                            accumulation + one block reduction per
//                                                                            12-column pass, as in phases.cuh dense_pass)
//   (b)  Z[:, q] += sum_c V[:, c] P[q, c]      for the tile's rows q         (= P[c, q] by symmetry; thread-local: the
//                                                                            thread that loaded row q owns its sum, kept
//                                                                            in shared memory across the passes)
// The first tile of a column panel contains the diagonal block; there rows q < c1 (the square block, read in full) only
// take part in (a).  Every tile writes its two partial results to its own slots; k_sym_reduce adds them per column of
// Z in tile order -- no atomics, bitwise reproducible whichever CTA processed which tile (tiles are handed out by an
// atomic counter for balance).  Partials: ~1.6 % of the bytes of P.
#include <vector>

#include "kernels.h"

namespace dpgo {

void count_launch();

namespace {

constexpr int kPassCols = 12;                 // columns per register pass (R x 12 accumulators)
constexpr int kPanel = 28 * kPassCols;        // 336 columns per tile
constexpr int kQC = 2;                        // rows per thread in flight (kQC x 12 loads of P before the first multiply)
// rows per tile: R x rows doubles of product (b) sums live in shared memory (164 KB at r = 5)
__host__ __device__ constexpr int rows_of(int r) { return r <= 6 ? 4096 : 2048; }

__host__ __device__ inline int panels_of(int n4) { return (n4 + kPanel - 1) / kPanel; }
// tiles of panel p: rows from its first column down, in chunks of rows_of(r)
__host__ __device__ inline int tiles_of_panel(int n4, int p, int r) {
  const int c0 = p * kPanel;
  return (n4 - c0 + rows_of(r) - 1) / rows_of(r);
}

}  // namespace

int sym_precond_tiles(int n4, int r, std::vector<int> &first_tile_of_panel) {
  const int np = panels_of(n4);
  first_tile_of_panel.assign(np + 1, 0);
  for (int p = 0; p < np; ++p) first_tile_of_panel[p + 1] = first_tile_of_panel[p] + tiles_of_panel(n4, p, r);
  return first_tile_of_panel[np];
}
size_t sym_precond_partial_doubles(int ntiles, int r) { return (size_t)ntiles * 8 * (kPanel + rows_of(r)); }

template <int R>
__global__ void __launch_bounds__(256, 1) k_sym_precond(const __grid_constant__ SymPrecondArgs a) {
  constexpr int kRows = rows_of(R);
  extern __shared__ __align__(16) double smem[];
  double *zb = smem;                              // [R][kRows]   product (b) sums of this tile's rows
  double *vc = zb + R * kRows;                    // [R][kPassCols] V of the pass's columns
  double *red = vc + 8 * kPassCols;               // [8 warps][16][8] cross-warp reduction of product (a)
  __shared__ int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = a.n4;
  const size_t ld = a.ld;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(a.tile_counter, 1);
    __syncthreads();
    const int t = s_tile;
    if (t >= a.ntiles) break;
    // tile geometry from the panel table
    int p = 0;
    {
      int lo = 0, hi = a.npanels;   // largest p with first_tile[p] <= t
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.first_tile[mid] <= t) lo = mid; else hi = mid;
      }
      p = lo;
    }
    const int c0 = p * kPanel, c1 = min(n4, c0 + kPanel);
    const int r0 = c0 + (t - a.first_tile[p]) * kRows, r1 = min(n4, r0 + kRows);
    double *pa = a.partials + (size_t)t * 8 * (kPanel + kRows);   // [8][kPanel]
    double *pb = pa + 8 * kPanel;                                 // [8][kRows]
    for (int i = threadIdx.x; i < R * kRows; i += 256) zb[i] = 0.0;
    for (int sub = c0; sub < c1; sub += kPassCols) {
      const int ncols = min(kPassCols, c1 - sub);
      __syncthreads();
      if (threadIdx.x < R * kPassCols) {
        const int aa = threadIdx.x / kPassCols, j = threadIdx.x % kPassCols;
        vc[aa * kPassCols + j] = (j < ncols) ? a.VT[(size_t)aa * n4 + sub + j] : 0.0;
      }
      __syncthreads();
      double acc[R][kPassCols];
#pragma unroll
      for (int aa = 0; aa < R; ++aa)
#pragma unroll
        for (int j = 0; j < kPassCols; ++j) acc[aa][j] = 0.0;
      const double *cols = a.P + (size_t)sub * ld;
      for (int q0 = r0 + threadIdx.x; q0 < r1; q0 += 256 * kQC) {
        double vr[kQC][R], pv[kQC][kPassCols];
#pragma unroll
        for (int i = 0; i < kQC; ++i) {
          const int q = q0 + 256 * i;
#pragma unroll
          for (int aa = 0; aa < R; ++aa) vr[i][aa] = (q < r1) ? a.VT[(size_t)aa * n4 + q] : 0.0;
#pragma unroll
          for (int j = 0; j < kPassCols; ++j) pv[i][j] = (q < r1 && j < ncols) ? cols[(size_t)j * ld + q] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < kQC; ++i) {
          const int q = q0 + 256 * i;
          double zrow[R];
#pragma unroll
          for (int aa = 0; aa < R; ++aa) zrow[aa] = 0.0;
#pragma unroll
          for (int j = 0; j < kPassCols; ++j) {
#pragma unroll
            for (int aa = 0; aa < R; ++aa) {
              acc[aa][j] = fma(vr[i][aa], pv[i][j], acc[aa][j]);
              zrow[aa] = fma(vc[aa * kPassCols + j], pv[i][j], zrow[aa]);
            }
          }
          if (q < r1 && q >= c1) {   // strictly below the diagonal block: the entry also stands for P[c, q]
#pragma unroll
            for (int aa = 0; aa < R; ++aa) zb[aa * kRows + (q - r0)] += zrow[aa];
          }
        }
      }
      // product (a): sum the 256 threads' accumulators in a fixed order -- reduce-scatter over 16 (12 + 4 zero)
      // columns inside the warp (xor 1, 2, 4, 8, then 16), warps in order through shared memory
      double h1[R];
      {
        double h8[R][8], h4[R][4], h2[R][2];
        {
          const bool hi = lane & 1;
#pragma unroll
          for (int aa = 0; aa < R; ++aa)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const double up = (c + 8 < kPassCols) ? acc[aa][c + 8 < kPassCols ? c + 8 : 0] : 0.0;
              const double keep = hi ? up : acc[aa][c];
              const double send = hi ? acc[aa][c] : up;
              h8[aa][c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        }
        {
          const bool hi = lane & 2;
#pragma unroll
          for (int aa = 0; aa < R; ++aa)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const double keep = hi ? h8[aa][c + 4] : h8[aa][c];
              const double send = hi ? h8[aa][c] : h8[aa][c + 4];
              h4[aa][c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
        }
        {
          const bool hi = lane & 4;
#pragma unroll
          for (int aa = 0; aa < R; ++aa)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const double keep = hi ? h4[aa][c + 2] : h4[aa][c];
              const double send = hi ? h4[aa][c] : h4[aa][c + 2];
              h2[aa][c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
        }
        {
          const bool hi = lane & 8;
#pragma unroll
          for (int aa = 0; aa < R; ++aa) {
            const double keep = hi ? h2[aa][1] : h2[aa][0];
            const double send = hi ? h2[aa][0] : h2[aa][1];
            double v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            h1[aa] = v;
          }
        }
      }
      __syncthreads();   // red reuse
      if (lane < 16) {   // lane l holds column ((l&1)<<3 | (l&2)<<1 | (l&4)>>1 | (l&8)>>3)
        const int col = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);
#pragma unroll
        for (int aa = 0; aa < R; ++aa) red[(warp * 16 + col) * 8 + aa] = h1[aa];
      }
      __syncthreads();
      if (threadIdx.x < 128) {
        const int j = threadIdx.x >> 3, aa = threadIdx.x & 7;
        if (aa < R && j < ncols) {
          double s = 0;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += red[(w * 16 + j) * 8 + aa];
          pa[(size_t)aa * kPanel + (sub - c0) + j] = s;
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * kRows; i += 256) {
      const int aa = i / kRows, rr = i % kRows;
      pb[(size_t)aa * kRows + rr] = zb[i];
    }
  }
}

// Zt[a][c] = sum of the partials that cover column c: product (a) of the tiles of c's own panel (tile order), then
// product (b) of the tile of every earlier panel whose rows contain c (panel order)
template <int R>
__global__ void k_sym_reduce(const __grid_constant__ SymPrecondArgs a) {
  constexpr int kRows = rows_of(R);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n4) return;
  const int p = c / kPanel;
  double s[R];
#pragma unroll
  for (int aa = 0; aa < R; ++aa) s[aa] = 0.0;
  for (int t = a.first_tile[p]; t < a.first_tile[p + 1]; ++t) {
    const double *pa = a.partials + (size_t)t * 8 * (kPanel + kRows);
#pragma unroll
    for (int aa = 0; aa < R; ++aa) s[aa] += pa[(size_t)aa * kPanel + (c - p * kPanel)];
  }
  for (int pp = 0; pp < p; ++pp) {
    const int c0 = pp * kPanel;
    const int k = (c - c0) / kRows;   // the tile of panel pp whose rows contain c (c >= c1 of that panel: pp < p)
    const int t = a.first_tile[pp] + k;
    const double *pb = a.partials + (size_t)t * 8 * (kPanel + kRows) + 8 * kPanel;
    const int rr = c - (c0 + k * kRows);
#pragma unroll
    for (int aa = 0; aa < R; ++aa) s[aa] += pb[(size_t)aa * kRows + rr];
  }
#pragma unroll
  for (int aa = 0; aa < R; ++aa) a.Zt[(size_t)aa * a.n4 + c] = s[aa];
}

template <int R>
static cudaError_t launch_sym_t(const SymPrecondArgs &a, int grid, cudaStream_t s) {
  const size_t smem = (size_t)(R * rows_of(R) + 8 * kPassCols + 8 * 16 * 8) * sizeof(double);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(k_sym_precond<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  cudaError_t e = cudaMemsetAsync(a.tile_counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  count_launch();
  k_sym_precond<R><<<grid, 256, smem, s>>>(a);
  count_launch();
  k_sym_reduce<R><<<(a.n4 + 127) / 128, 128, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_sym_precond(const SymPrecondArgs &a, int r, int grid, cudaStream_t s) {
  switch (r) {
    case 3: return launch_sym_t<3>(a, grid, s);
    case 4: return launch_sym_t<4>(a, grid, s);
    case 5: return launch_sym_t<5>(a, grid, s);
    case 6: return launch_sym_t<6>(a, grid, s);
    case 7: return launch_sym_t<7>(a, grid, s);
    case 8: return launch_sym_t<8>(a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace dpgo
