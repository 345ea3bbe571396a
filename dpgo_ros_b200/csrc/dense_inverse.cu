// Dense SPD inverse on the device: builds (Q + lambda I)^-1 for the
// preconditioner (a6).  The reference factors Q + lambda I with CHOLMOD and
// runs two sparse triangular solves per application -- sequential work that a
// GPU cannot hide.  Here the factorisation is paid once per Q change and each
// application becomes one streaming (r x 4n)(4n x 4n) product (phases.cuh).
//
// Round 2: recursive blocked form in which everything but the 64 x 64 diagonal
// leaves is one FP64 GEMM kernel (k_gemm):
//
//   F(A, W, n):   A = [A11 . ; A21 A22] (lower)  ->  W = chol(A)^-1 (lower)
//     F(A11, W11)
//     L21 = A21 W11^T            (into the W21 block, free at that point)
//     A22 -= L21 L21^T           (lower tiles only)
//     F(A22, W22)
//     T   = L21 W11              (into the A21 block, dead by then)
//     W21 = -W22 T
//   P = W^T W                    (lower tiles, mirrored into the upper triangle)
//
// The triangular operands are skipped at tile granularity through per-tile k
// ranges.  N^3 / 2 multiply-adds in total; N is a multiple of 32.
// Round 1's version (32-wide right-looking tiles, one serial single-warp kernel
// per diagonal block and per row of the triangular inverse) spent 86 % of its
// time in those serial kernels (profiles/launches_r2_summary.csv).
#include <atomic>
#include <cstdlib>

#include "kernels.h"

namespace dpgo {

// ---------------------------------------------------------------------------
// k_scatter_blocks: block-CSR (Q) -> dense column-major P, + lambda on the diagonal
// ---------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------
// k_leaf: W = chol(A)^-1 for one diagonal block of n <= 64 rows (one CTA).
// Both 64 x 64 triangles live in registers: thread (ty, tx) of 16 x 16 owns the
// 4 x 4 blocks S(4 ty + i, 4 tx + c) of the matrix being eliminated and
// X(4 ty + i, 4 tx + c) of the inverse being built.  Square-root-free
// elimination, ONE barrier per column: the owners of column j of S and of row j
// of X post them to a double-buffered shared-memory strip, then every thread
// applies
//     S(t, c) -= v_t v_c / d_j      (c > j)         v = column j, d_j = S(j, j)
//     X(i, c) -= (v_i / d_j) X(j, c)  (i > j)
// to its blocks (6 LDS.128 + 32 FMA per step; the owner of d_{j+1} posts its
// reciprocal with the column).  Entries above the diagonal carry garbage that never
// reaches a valid entry; the 1 / sqrt(d) scalings are applied on the way out.  (A first version kept both triangles in shared memory and
// updated them with read-modify-write loops: 55 us per leaf, a third of the
// whole inverse at N = 5000.)
// ---------------------------------------------------------------------------
constexpr int kLeaf = 64;

__global__ void __launch_bounds__(256) k_leaf(const double *A, size_t lda, double *W, size_t ldw, int n, int off,
                                              int *info) {
  __shared__ __align__(16) double colv[2][kLeaf];  // column j of S
  __shared__ __align__(16) double rowx[2][kLeaf];  // row j of X
  __shared__ double diag[kLeaf];
  __shared__ double rdv[2];  // 1 / d_j
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  double S[4][4], X[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * ty + i, cc = 4 * tx + c;
      const bool in = r < n && cc < n;
      // outside the block: identity, so that the elimination runs over all 64 columns without a special case
      S[i][c] = in ? (r >= cc ? A[(size_t)cc * lda + r] : 0.0) : (r == cc ? 1.0 : 0.0);
      X[i][c] = (r == cc) ? 1.0 : 0.0;
    }
  // post column 0 / row 0 / 1 / d_0
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) colv[0][4 * ty + i] = S[i][0];
  }
  if (ty == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) rowx[0][4 * tx + c] = X[0][c];
  }
  if (tid == 0) {
    const double d0 = S[0][0];
    if (!(d0 > 0.0)) atomicExch(info, off + 1);
    diag[0] = d0;
    rdv[0] = 1.0 / d0;
  }
  __syncthreads();
  for (int jb = 0; jb < kLeaf / 4; ++jb) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {  // (unrolled: which register of a block holds column / row j + 1 is then static)
      const int b = jj & 1;           // j = 4 jb + jj
      const double rd = rdv[b];
      const double2 vr0 = *reinterpret_cast<const double2 *>(&colv[b][4 * ty]), vr1 = *reinterpret_cast<const double2 *>(&colv[b][4 * ty + 2]);
      const double2 vc0 = *reinterpret_cast<const double2 *>(&colv[b][4 * tx]), vc1 = *reinterpret_cast<const double2 *>(&colv[b][4 * tx + 2]);
      const double2 xr0 = *reinterpret_cast<const double2 *>(&rowx[b][4 * tx]), xr1 = *reinterpret_cast<const double2 *>(&rowx[b][4 * tx + 2]);
      const double vr[4] = {vr0.x * rd, vr0.y * rd, vr1.x * rd, vr1.y * rd};  // v_t / d_j
      const double vc[4] = {vc0.x, vc0.y, vc1.x, vc1.y};
      const double xr[4] = {xr0.x, xr0.y, xr1.x, xr1.y};
      // S: unconditional.  What it overwrites besides the live entries (t >= c > j) are columns <= j, rows <= j and the
      // strict upper triangle -- all dead or garbage by construction, and none of them feeds a live entry.
      // X: rows <= j are final, so their multiplier is zeroed (X(j, c) is finite).
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool below = ty > jb || (ty == jb && i > jj);  // 4 ty + i > j
        const double vx = below ? vr[i] : 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          S[i][c] = fma(-vr[i], vc[c], S[i][c]);
          X[i][c] = fma(-vx, xr[c], X[i][c]);
        }
      }
      // post column j + 1 of S, row j + 1 of X (final after this step) and 1 / d_{j+1} to the other strip
      const int ob = (jj == 3) ? jb + 1 : jb;  // block that owns j + 1
      const int oi = (jj + 1) & 3;             // its position inside the block (static)
      if (ob < kLeaf / 4) {
        if (tx == ob) {
#pragma unroll
          for (int i = 0; i < 4; ++i) colv[b ^ 1][4 * ty + i] = S[i][oi];
          if (ty == ob) {
            const double dn = S[oi][oi];
            const int jn = 4 * ob + oi;
            if (jn < n && !(dn > 0.0)) atomicExch(info, off + jn + 1);
            diag[jn] = dn;
            rdv[b ^ 1] = 1.0 / dn;
          }
        }
        if (ty == ob) {
#pragma unroll
          for (int c = 0; c < 4; ++c) rowx[b ^ 1][4 * tx + c] = X[oi][c];
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * ty + i, cc = 4 * tx + c;
      if (r < n && cc < n) W[(size_t)cc * ldw + r] = (r >= cc) ? X[i][c] * rsqrt(diag[r]) : 0.0;
    }
}

// ---------------------------------------------------------------------------
// k_gemm: C (M x N, column-major) = alpha * A * B + beta * C on sub-blocks of
// column-major storage.
//   A(i, k): AK ? Ap[k + i * lda] : Ap[i + k * lda]
//   B(k, j): BK ? Bp[k + j * ldb] : Bp[j + k * ldb]
// CTA tile (64 HM) x (64 HN), 8 warps as 2 x 4, the products on the FP64 tensor
// path (mma.sync.m8n8k4.f64, DMMA).  B200 runs DMMA at the DFMA rate (64-73 FMA /
// clk / SM, profiles/microbench_r1.txt), so this is not about a higher peak: a
// register-tiled DFMA loop needs one LDS.128 (4 shared-memory wavefronts, broadcast
// or not) per 8 FMA per thread and measured 51 % FP64-pipe utilisation with
// mio_throttle on top of the stall list (ncu, round 2); the fragment form needs 4 x
// fewer wavefronts per FMA.  K in chunks of 32 (16 for the large tile), two stages, global -> register
// prefetch of the next chunk during the product.  M, N multiples of 32 (guards), K
// ranges multiples of 32.
// ---------------------------------------------------------------------------
constexpr int kPad = 4;  // stage row stride 64 H + 4 doubles = 4 mod 16: conflict-free fragment loads, 16-byte aligned rows

enum KRange { K_FULL = 0, K_LO_NTILE = 1, K_HI_NTILE = 2, K_HI_MTILE = 3, K_LO_MTILE = 4 };

struct GemmArgs {
  const double *A, *B;
  double *C;
  size_t lda, ldb, ldc;
  int M, N, K;
  double alpha, beta;
  int krange;      // KRange
  int lower_only;  // skip tiles strictly above the diagonal (C square, same origin on the diagonal)
  int mirror;      // also write C(j, i)
};

template <int H, bool KC, int D>
__device__ __forceinline__ void gemm_fetch(const double *P, size_t ld, int x0, int X, int k0, int tid,
                                           double2 (&reg)[H * D / 8]) {
  // tile of (64 H) x D doubles = 32 H D double2 slots, H D / 8 per thread
#pragma unroll
  for (int q = 0; q < H * D / 8; ++q) {
    const int s = tid + 256 * q;
    double2 v = make_double2(0.0, 0.0);
    if (KC) {
      const int k2 = s % (D / 2), x = s / (D / 2);  // pairs along k
      if (x0 + x < X) v = *reinterpret_cast<const double2 *>(P + (size_t)(x0 + x) * ld + k0 + 2 * k2);
    } else {
      const int x2 = s % (32 * H), k = s / (32 * H);  // pairs along the tile dimension
      if (x0 + 2 * x2 < X) v = *reinterpret_cast<const double2 *>(P + (size_t)(k0 + k) * ld + x0 + 2 * x2);
    }
    reg[q] = v;
  }
}

template <int H, bool KC, int D>
__device__ __forceinline__ void gemm_stash(double *stage, int tid, const double2 (&reg)[H * D / 8]) {
  constexpr int LDS_ = 64 * H + kPad;
#pragma unroll
  for (int q = 0; q < H * D / 8; ++q) {
    const int s = tid + 256 * q;
    if (KC) {
      const int k2 = s % (D / 2), x = s / (D / 2);
      stage[(2 * k2) * LDS_ + x] = reg[q].x;
      stage[(2 * k2 + 1) * LDS_ + x] = reg[q].y;
    } else {
      const int x2 = s % (32 * H), k = s / (32 * H);
      *reinterpret_cast<double2 *>(stage + k * LDS_ + 2 * x2) = reg[q];
    }
  }
}

// K chunk: 32 for the 64 x 64 tiles (their products are short, so the chain of chunks -- fetch latency + stash + barrier
// each -- is what a small product costs), 16 for the 128 x 128 tiles (register budget)
template <int HM, int HN>
struct GemmChunk {
  static constexpr int value = (HM * HN >= 4) ? 16 : 32;
};

template <int HM, int HN, bool AK, bool BK>
__global__ void __launch_bounds__(256, (HM * HN >= 4) ? 1 : 2) k_gemm(GemmArgs g) {
  constexpr int BM = 64 * HM, BN = 64 * HN, kBK = GemmChunk<HM, HN>::value;
  constexpr int LDA_ = BM + kPad, LDB_ = BN + kPad;
  extern __shared__ __align__(16) double smem[];
  double *As = smem;                     // [2][kBK][LDA_]
  double *Bs = smem + 2 * kBK * LDA_;    // [2][kBK][LDB_]
  const int tid = threadIdx.x;
  // Tiles are handed out longest first (the hardware assigns CTAs in linear order): with triangular operands the k
  // range, hence the work, varies by the tile's row or column -- in natural order the heaviest tiles of a K_HI range
  // would all start last and set the kernel's tail.
  const int ntm = gridDim.x, ntn = gridDim.y;
  const int lin = blockIdx.x + ntm * blockIdx.y;
  int ti, tj;
  switch (g.krange) {
    case K_HI_NTILE: tj = ntn - 1 - lin / ntm; ti = lin % ntm; break;
    case K_HI_MTILE: ti = ntm - 1 - lin / ntn; tj = lin % ntn; break;
    case K_LO_MTILE: ti = lin / ntn; tj = lin % ntn; break;
    default: tj = lin / ntm; ti = lin % ntm; break;  // K_FULL, K_LO_NTILE
  }
  const int m0 = ti * BM, n0 = tj * BN;
  if (g.lower_only && m0 + BM - 1 < n0) return;
  int klo = 0, khi = g.K;
  switch (g.krange) {
    case K_LO_NTILE: klo = n0; break;
    case K_HI_NTILE: khi = min(g.K, n0 + BN); break;
    case K_HI_MTILE: khi = min(g.K, m0 + BM); break;
    case K_LO_MTILE: klo = m0; break;
    default: break;
  }
  // 8 warps as 2 (m) x 4 (n); a warp owns MT x NT tiles of 8 x 8 (mma.m8n8k4.f64: lane = 4 g + t holds A(g, t),
  // B(t, g) and C(g, 2 t), C(g, 2 t + 1))
  constexpr int MT = 4 * HM, NT = 2 * HN;
  const int lane = tid & 31, w = tid >> 5;
  const int fg = lane >> 2, ft = lane & 3;
  const int wm0 = (w & 1) * (BM / 2), wn0 = (w >> 1) * (BN / 4);
  double c[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) c[i][j][0] = c[i][j][1] = 0.0;

  const int nchunk = (khi - klo) / kBK;
  double2 ra[HM * kBK / 8], rb[HN * kBK / 8];
  if (nchunk > 0) {
    gemm_fetch<HM, AK, kBK>(g.A, g.lda, m0, g.M, klo, tid, ra);
    gemm_fetch<HN, BK, kBK>(g.B, g.ldb, n0, g.N, klo, tid, rb);
    gemm_stash<HM, AK, kBK>(As, tid, ra);
    gemm_stash<HN, BK, kBK>(Bs, tid, rb);
  }
  __syncthreads();
  for (int ch = 0; ch < nchunk; ++ch) {
    const int st = ch & 1;
    const bool more = ch + 1 < nchunk;
    if (more) {
      gemm_fetch<HM, AK, kBK>(g.A, g.lda, m0, g.M, klo + (ch + 1) * kBK, tid, ra);
      gemm_fetch<HN, BK, kBK>(g.B, g.ldb, n0, g.N, klo + (ch + 1) * kBK, tid, rb);
    }
    // fragment loads: row (4 kk + t) of the stage, element wm0 + 8 i + g -- with a row stride of 4 mod 16 doubles the
    // 16 lanes of a half-warp (t = 0..3 x g = 0..3 or 4..7) hit 16 different 8-byte bank pairs
    const double *as = As + st * kBK * LDA_ + ft * LDA_ + wm0 + fg;
    const double *bs = Bs + st * kBK * LDB_ + ft * LDB_ + wn0 + fg;
#pragma unroll
    for (int kk = 0; kk < kBK / 4; ++kk) {
      double a[MT], b[NT];
#pragma unroll
      for (int i = 0; i < MT; ++i) a[i] = as[kk * 4 * LDA_ + 8 * i];
#pragma unroll
      for (int j = 0; j < NT; ++j) b[j] = bs[kk * 4 * LDB_ + 8 * j];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
    }
    if (more) {
      gemm_stash<HM, AK, kBK>(As + (st ^ 1) * kBK * LDA_, tid, ra);
      gemm_stash<HN, BK, kBK>(Bs + (st ^ 1) * kBK * LDB_, tid, rb);
    }
    __syncthreads();
  }
  // epilogue: C(row, col .. col + 1) per 8 x 8 tile; the 8 lanes of a column pair cover 8 consecutive rows
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int row = m0 + wm0 + 8 * i + fg, col = n0 + wn0 + 8 * j + 2 * ft;
      if (row >= g.M || col >= g.N) continue;  // M, N are multiples of 32: a pair of columns is inside or outside
      double v0 = g.alpha * c[i][j][0], v1 = g.alpha * c[i][j][1];
      double *d0 = g.C + (size_t)col * g.ldc + row, *d1 = d0 + g.ldc;
      if (g.beta != 0.0) {
        v0 += g.beta * *d0;
        v1 += g.beta * *d1;
      }
      *d0 = v0;
      *d1 = v1;
      if (g.mirror) *reinterpret_cast<double2 *>(g.C + (size_t)row * g.ldc + col) = make_double2(v0, v1);
    }
}

static std::atomic<long long> g_inv_launches{0};  // (robots on several host threads build at the same time)
long long dense_inverse_launch_count() { return g_inv_launches.load(); }

template <int HM, int HN, bool AK, bool BK>
static cudaError_t gemm_launch_t(const GemmArgs &g, cudaStream_t s) {
  constexpr size_t smem = sizeof(double) * 2 * GemmChunk<HM, HN>::value * ((64 * HM + kPad) + (64 * HN + kPad));
  static std::atomic<bool> attr_done[64];  // (the attribute is per device)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !attr_done[dev].load()) {
    e = cudaFuncSetAttribute(k_gemm<HM, HN, AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev].store(true);
  }
  dim3 grid((g.M + 64 * HM - 1) / (64 * HM), (g.N + 64 * HN - 1) / (64 * HN));
  ++g_inv_launches;
  k_gemm<HM, HN, AK, BK><<<grid, 256, smem, s>>>(g);
  return cudaGetLastError();
}

template <bool AK, bool BK>
static cudaError_t gemm_launch(const GemmArgs &g, cudaStream_t s) {
  // 128 x 128 tiles halve the L2 -> SM operand traffic (64 x 64 tiles would need ~4.6 TB/s at the FP64 peak), but with
  // triangular operands the heaviest tile carries twice the average work: they are used once a GPU-full of them is a
  // small part of the product (the longest tile then stays below the per-SM share)
  const long tiles128 = (long)((g.M + 127) / 128) * ((g.N + 127) / 128) / (g.lower_only ? 2 : 1);
  static const long big_from = getenv("DPGO_B200_GEMM_BIG_TILES") ? atol(getenv("DPGO_B200_GEMM_BIG_TILES")) : 4 * 148;  // (diagnostics)
  if (tiles128 >= big_from) return gemm_launch_t<2, 2, AK, BK>(g, s);
  return gemm_launch_t<1, 1, AK, BK>(g, s);
}

cudaError_t launch_scatter_blocks(double *P, size_t ld, const int *rowptr, const int *col, const double *val,
                                  int n, double lambda, int npad, cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(P, 0, sizeof(double) * ld * ld, s);
  if (err != cudaSuccess) return err;
  const int extra = (npad - 4 * n + 63) / 64;
  ++g_inv_launches;
  k_scatter_blocks<<<n + extra, 64, 0, s>>>(P, ld, rowptr, col, val, n, lambda, npad);
  return cudaGetLastError();
}

namespace {

struct InvCtx {
  double *A, *W;
  size_t ld;
  int *info;
  cudaStream_t s;
  cudaError_t err = cudaSuccess;
};

inline double *blk(double *base, size_t ld, int r, int c) { return base + (size_t)c * ld + r; }

void chol_inv_rec(InvCtx &X, int o, int n) {
  if (X.err != cudaSuccess) return;
  if (n <= kLeaf) {
    ++g_inv_launches;
    k_leaf<<<1, 256, 0, X.s>>>(blk(X.A, X.ld, o, o), X.ld, blk(X.W, X.ld, o, o), X.ld, n, o, X.info);
    X.err = cudaGetLastError();
    return;
  }
  const int n1 = ((n / 2 + 63) / 64) * 64, n2 = n - n1;
  chol_inv_rec(X, o, n1);
  if (X.err != cudaSuccess) return;
  double *A21 = blk(X.A, X.ld, o + n1, o), *A22 = blk(X.A, X.ld, o + n1, o + n1);
  double *W11 = blk(X.W, X.ld, o, o), *W21 = blk(X.W, X.ld, o + n1, o), *W22 = blk(X.W, X.ld, o + n1, o + n1);
  GemmArgs g{};
  g.lda = g.ldb = g.ldc = X.ld;
  // L21 = A21 W11^T  ->  W21 block.  B(k, j) = W11(j, k): n-contiguous, nonzero for k <= j
  g.A = A21; g.B = W11; g.C = W21; g.M = n2; g.N = n1; g.K = n1; g.alpha = 1.0; g.beta = 0.0;
  g.krange = K_HI_NTILE; g.lower_only = 0; g.mirror = 0;
  if ((X.err = gemm_launch<false, false>(g, X.s)) != cudaSuccess) return;
  // A22 -= L21 L21^T (lower tiles)
  g.A = W21; g.B = W21; g.C = A22; g.M = n2; g.N = n2; g.K = n1; g.alpha = -1.0; g.beta = 1.0;
  g.krange = K_FULL; g.lower_only = 1;
  if ((X.err = gemm_launch<false, false>(g, X.s)) != cudaSuccess) return;
  chol_inv_rec(X, o + n1, n2);
  if (X.err != cudaSuccess) return;
  // T = L21 W11  ->  A21 block.  B(k, j) = W11(k, j): k-contiguous, nonzero for k >= j
  g.A = W21; g.B = W11; g.C = A21; g.M = n2; g.N = n1; g.K = n1; g.alpha = 1.0; g.beta = 0.0;
  g.krange = K_LO_NTILE; g.lower_only = 0;
  if ((X.err = gemm_launch<false, true>(g, X.s)) != cudaSuccess) return;
  // W21 = -W22 T.  A(i, k) = W22(i, k): nonzero for k <= i
  g.A = W22; g.B = A21; g.C = W21; g.M = n2; g.N = n1; g.K = n2; g.alpha = -1.0; g.beta = 0.0;
  g.krange = K_HI_MTILE;
  X.err = gemm_launch<false, true>(g, X.s);
}

}  // namespace

// A (N x N, column-major, ld = N, lower triangle read) <- A^-1 (both triangles); work: N x N doubles
cudaError_t spd_inverse(double *A, double *work, int N, int *d_info, cudaStream_t s) {
  if (N <= 0 || N % 32 != 0) return cudaErrorInvalidValue;
  const size_t ld = N;
  cudaError_t err = cudaMemsetAsync(d_info, 0, sizeof(int), s);
  if (err != cudaSuccess) return err;
  // the strictly upper triangle of W is read as zeros by the triangular products
  err = cudaMemsetAsync(work, 0, sizeof(double) * ld * ld, s);
  if (err != cudaSuccess) return err;
  InvCtx X{A, work, ld, d_info, s};
  chol_inv_rec(X, 0, N);
  if (X.err != cudaSuccess) return X.err;
  // P = W^T W: A(i, k) = W(k, i), B(k, j) = W(k, j), both k-contiguous; k >= i on the lower tiles
  GemmArgs g{};
  g.lda = g.ldb = g.ldc = ld;
  g.A = work; g.B = work; g.C = A; g.M = N; g.N = N; g.K = N; g.alpha = 1.0; g.beta = 0.0;
  g.krange = K_LO_MTILE; g.lower_only = 1; g.mirror = 1;
  return gemm_launch<true, true>(g, s);
}

}  // namespace dpgo
