// Phases of one RBCD iteration as grid-cooperative device functions.  Each
// phase is called by every CTA of the grid; phases are separated by
// grid_barrier / grid_reduce in the persistent kernel (kernels.cu) or by kernel
// boundaries in the single-op kernels that back the parity hooks.
//
// Everything here is latency-bound (a 312-pose agent is ~200 KB of work spread
// over 148 SMs), so the code is organised around the number of DEPENDENT memory
// round trips per phase (~600 cycles each to L2), not around bytes:
//   * sparse phases read an ELL(8) copy of Q: slot columns and 4x4 blocks of a
//     pose sit at fixed addresses, so a pose costs two trips (indices+blocks,
//     then the gathered poses) instead of a CSR pointer chase per block;
//   * the dense preconditioner slab of a CTA is fetched by ONE TMA bulk copy
//     (cp.async.bulk -> shared memory, mbarrier completion) that is issued an
//     iteration ahead, so its HBM/L2 latency is off the critical path.
//
// Reference call sites (relative to the reference repo): the arithmetic lives
// in the un-vendored mit-acl/dpgo; what is cited is the wrapper line that
// triggers it.  iterate(): src/PGOAgentROS.cpp:160,1185.
#pragma once
#include "device.cuh"

namespace dpgo {

// debug timeline (diagnostics): clock64() marks of CTA 0 / thread 0, enabled by launch_set_dbg()
static __device__ long long *g_dbg = nullptr;
#define DBG(k) do { if (g_dbg && threadIdx.x == 0 && blockIdx.x == 0) g_dbg[(k)] = clock64(); } while (0)

// ---- pose -> 8-lane-group mapping --------------------------------------------
// item (pose) = blockIdx.x + gridDim.x * (local_group + 32 k): consecutive poses
// land on different SMs so a 300-pose agent is spread over the whole chip.
struct PoseIter {
  int a;       // row handled by this lane
  int lg;      // local group id 0..31 (absolute: also selects the staging tile)
  int k;       // loop counter
  int lgp;     // group id within the participating warp subset (< 0: this warp sits the phase out)
  int gpc;     // participating groups per CTA
  __device__ __forceinline__ PoseIter() : a(threadIdx.x & 7), lg(threadIdx.x >> 3), k(0), lgp(threadIdx.x >> 3),
                                          gpc(kGroupsPerCta) {}
  // restrict the phase to warps [warp_lo, warp_lo + warp_cnt) so two independent phases can run side by side
  __device__ __forceinline__ PoseIter(int warp_lo, int warp_cnt)
      : a(threadIdx.x & 7), lg(threadIdx.x >> 3), k(0), gpc(4 * warp_cnt) {
    const int w = (int)(threadIdx.x >> 5) - warp_lo;
    lgp = (w >= 0 && w < warp_cnt) ? lg - 4 * warp_lo : -1;
  }
  // warp-uniform "any lane of my warp still has work" + my item
  __device__ __forceinline__ bool next(int total, int &item) {
    if (lgp < 0) return false;
    const int g0 = (lgp & ~3) + gpc * k;  // first group of my warp at this step
    if ((int)blockIdx.x + (int)gridDim.x * g0 >= total) return false;
    item = (int)blockIdx.x + (int)gridDim.x * (lgp + gpc * k);
    ++k;
    return true;
  }
};

__device__ __forceinline__ void publish(const int *rowptr, double *const *dst, int j, int r, int a, bool act,
                                        const double (&x)[4]) {
  const int e0 = rowptr[j], e1 = rowptr[j + 1];
  for (int e = e0; e < e1; ++e) st4(dst[e], r, a, act, x);
}
__device__ __forceinline__ void publish_range(int e0, int e1, double *const *dst, int r, int a, bool act,
                                              const double (&x)[4]) {
  for (int e = e0; e < e1; ++e) st4(dst[e], r, a, act, x);
}

// The first two destinations of a pose's publication entries, requested as soon as the entry range is known so that
// the pointer loads overlap the pose's arithmetic instead of forming a dependent L2 trip right before the stores
// (a pose is public towards at most two neighbours in the chain / ring splits; further entries take the loop).
struct PubPtrs {
  double *d0, *d1;
  int e0, e1;
};
__device__ __forceinline__ PubPtrs pub_prefetch(double *const *dst, int e0, int e1) {
  PubPtrs p;
  p.e0 = e0;
  p.e1 = e1;
  p.d0 = (e0 < e1) ? dst[e0] : nullptr;
  p.d1 = (e0 + 1 < e1) ? dst[e0 + 1] : nullptr;
  return p;
}
__device__ __forceinline__ void publish_pre(const PubPtrs &p, double *const *dst, int r, int a, bool act,
                                            const double (&x)[4]) {
  if (p.d0) st4(p.d0, r, a, act, x);
  if (p.d1) st4(p.d1, r, a, act, x);
  for (int e = p.e0 + 2; e < p.e1; ++e) st4(dst[e], r, a, act, x);
}

// out_row(1x4) += x_row(1x4) * B(4x4 col-major)
__device__ __forceinline__ void row_times_block(const double (&x)[4], const double *B, double (&acc)[4]) {
#pragma unroll
  for (int cp = 0; cp < 4; ++cp) {
    const double b0 = B[cp * 4 + 0], b1 = B[cp * 4 + 1], b2 = B[cp * 4 + 2], b3 = B[cp * 4 + 3];
    acc[cp] = fma(x[0], b0, fma(x[1], b1, fma(x[2], b2, fma(x[3], b3, acc[cp]))));
  }
}

// shuffle within the 8-lane group
__device__ __forceinline__ int gshfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

// ---------------------------------------------------------------------------
// Phase A -- Nesterov bookkeeping of iterate() for every local agent (a7):
//   Y = proj((1-alpha) X + alpha V); non-selected agents: X = Y (V = proj(V) = V);
//   restart iterations: non-selected agents V = Y = X.
// Publishes Y (aux) and, for non-selected agents, X (reg) into the neighbours'
// inboxes (a9: getAuxSharedPoseDictWithNeighbor :666 / updateAuxNeighborPoses :1278).
// ---------------------------------------------------------------------------
// Speculated iterate(false) steps of a stand-alone agent that the host has consumed without a launch (see
// phase_lookahead): the agent's state is then X = Y = Xsrc, V = Vsrc instead of what A.X / A.V / A.Y hold, and
// the first kernel that runs afterwards commits it while it loads the poses anyway.
struct LaCommit {
  const double *X, *V;  // null: nothing to commit
};

// per-pose body (group-collective)
template <int RC>
__device__ __forceinline__ void nesterov_pose(const TeamDev &T, int ai, int j, bool valid, int a, int sel_local,
                                              bool restart, double alpha, LaCommit lc = LaCommit{nullptr, nullptr},
                                              bool commit_only = false) {
  const AgentDev &A = T.ag[ai];
  const int r = rdim<RC>(A);
  const bool act = valid && a < r;
  const size_t off = (size_t)j * 4 * r;
  double x[4], v[4], m[4];
  int pe0 = 0, pe1 = 0;
  if (valid) {
    pe0 = A.pub_rowptr[j];
    pe1 = A.pub_rowptr[j + 1];
  }
  ld4((lc.X ? lc.X : A.X) + off, r, a, act, x);
  if (lc.X) {
    ld4((lc.V ? lc.V : A.V) + off, r, a, act, v);
    if (valid) {
      st4(A.X + off, r, a, act, x);
      st4(A.Y + off, r, a, act, x);
      if (lc.V) st4(A.V + off, r, a, act, v);
    }
    if (commit_only) return;
  }
  if (restart) {
    if (ai != sel_local && valid) {
      st4(A.V + off, r, a, act, x);
      st4(A.Y + off, r, a, act, x);
      publish_range(pe0, pe1, A.pub_dst_aux, r, a, act, x);
      publish_range(pe0, pe1, A.pub_dst_reg, r, a, act, x);
    }
    return;
  }
  if (!lc.X) ld4(A.V + off, r, a, act, v);
  const PubPtrs paux = pub_prefetch(A.pub_dst_aux, pe0, pe1);
  const PubPtrs preg = (ai != sel_local) ? pub_prefetch(A.pub_dst_reg, pe0, pe1) : PubPtrs{nullptr, nullptr, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) m[c] = (1.0 - alpha) * x[c] + alpha * v[c];
  if (!valid) {  // keep idle groups on the fast path of sym3_invsqrt
    m[0] = (a == 0);
    m[1] = (a == 1);
    m[2] = (a == 2);
  }
  stiefel_project_row(m);
  if (valid) {
    st4(A.Y + off, r, a, act, m);
    publish_pre(paux, A.pub_dst_aux, r, a, act, m);
    if (ai != sel_local) {
      st4(A.X + off, r, a, act, m);
      publish_pre(preg, A.pub_dst_reg, r, a, act, m);
    }
  }
}

// pose -> group by PoseIter (stand-alone kernels)
template <int RC>
__device__ __forceinline__ void phase_nesterov(const TeamDev &T, int sel_local, bool restart, double alpha,
                                               LaCommit lc = LaCommit{nullptr, nullptr}, bool commit_only = false) {
  PoseIter it;
  const int total = T.pose_prefix[T.num_local];
  int item;
  while (it.next(total, item)) {
    const bool valid = item < total;
    int ai = 0;
    if (valid) {
      while (item >= T.pose_prefix[ai + 1]) ++ai;
    }
    nesterov_pose<RC>(T, ai, valid ? item - T.pose_prefix[ai] : 0, valid, it.a, sel_local, restart, alpha, lc,
                      commit_only);
  }
}

// ---------------------------------------------------------------------------
// Lookahead for the stand-alone (per-robot API) path.  iterate(false) of an accelerated agent
// (src/PGOAgentROS.cpp:1185) depends on nothing but the agent's own X and V:
//     Y = proj((1-alpha) X + alpha V), X = Y     (restart iterations: V = Y = X)
// and in the synchronous schedule an agent answers N-1 of them between two solves.  A launch costs ~10 us
// before its result is visible to the host, the arithmetic ~1 us, so the kernel that ends a solve (or an
// iterate(false) that had to launch) also SPECULATES the next `depth` iterate(false) steps, pose by pose with
// no grid sync: the states go to A.LX[j], the public poses to lookahead outbox j in mapped host memory.
// The host then serves iterate(false) + getSharedPoseDict for those steps without touching the GPU and the
// next launch commits the consumed state (LaCommit).
// ---------------------------------------------------------------------------
template <int RC, class Iter>
__device__ __forceinline__ void lookahead_pose(const AgentDev &A, int j, bool valid, int a, int depth,
                                               const double2 *tab) {
  const int r = rdim<RC>(A);
  const bool act = valid && a < r;
  const size_t off = (size_t)j * 4 * r, vec = (size_t)4 * r * A.n;
  double x[4], v[4];
  int pe0 = 0, pe1 = 0;
  if (valid) {
    pe0 = A.pub_rowptr[j];
    pe1 = A.pub_rowptr[j + 1];
  }
  ld4(A.X + off, r, a, act, x);
  ld4(A.V + off, r, a, act, v);
  for (int s = 0; s < depth; ++s) {
    const double alpha = tab[s].x;
    if (tab[s].y != 0.0) {  // restart iteration: V = Y = X
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = x[c];
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) x[c] = (1.0 - alpha) * x[c] + alpha * v[c];
      if (!valid) {
        x[0] = (a == 0); x[1] = (a == 1); x[2] = (a == 2);
      }
      stiefel_project_row(x);
    }
    if (valid) {
      st4(A.LX + s * vec + off, r, a, act, x);
      for (int e = pe0; e < pe1; ++e)
        st4(A.la_out + (size_t)s * A.la_stride + (A.pub_dst_reg[e] - A.outbox_base), r, a, act, x);
    }
  }
}
template <int RC>
__device__ __forceinline__ void phase_lookahead(const AgentDev &A, int depth, const double2 *tab) {
  PoseIter it;
  int j;
  while (it.next(A.n, j)) lookahead_pose<RC, PoseIter>(A, j < A.n ? j : 0, j < A.n, it.a, depth, tab);
}

// Per-CTA pose ownership used by the persistent kernel: CTA b owns the balanced chunk
// [p0, p0+np) of EVERY local agent, for the Nesterov phase and for the dense step alike, so a
// pose's X / V / Y are always written and then re-read by the same CTA (no grid sync needed
// between the end of one iteration and the Nesterov phase of the next).
struct ChunkTable {
  int p0[kMaxLocal], np[kMaxLocal], prefix[kMaxLocal + 1];
};
template <int RC>
__device__ __forceinline__ void phase_nesterov_chunk(const TeamDev &T, const ChunkTable &ct, int sel_local,
                                                     bool restart, double alpha,
                                                     LaCommit lc = LaCommit{nullptr, nullptr}) {
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  const int total = ct.prefix[T.num_local];
  for (int base = 0; base < total; base += kGroupsPerCta) {
    const int item = base + lg;
    const bool valid = item < total;
    int ai = 0;
    if (valid) {
      while (item >= ct.prefix[ai + 1]) ++ai;
    }
    nesterov_pose<RC>(T, ai, valid ? ct.p0[ai] + item - ct.prefix[ai] : 0, valid, a, sel_local, restart, alpha, lc);
  }
}

// ---------------------------------------------------------------------------
// Sparse gather for one pose: acc += sum over the pose's ELL slots (and CSR
// overflow) of  V_{col} * block.  Two dependent round trips: (slot columns +
// blocks) then (gathered rows).  The 4x4 blocks are staged through a per-group
// shared-memory tile because every lane needs all 16 entries of every block.
//   W      : ELL width (8 for Q, 4 for the neighbour term)
//   stage  : this group's staging tile, W*16 doubles
// ---------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void ell_gather(const int *ell_col, const double *ell_val, const int *ovf_rowptr,
                                           const int *ovf_col, const double *ovf_val, const double *Vsrc, int j,
                                           bool valid, int r, int a, bool act, double *stage, double (&acc)[4]) {
  // trip 1: my slot's column, my share of the blocks, the overflow range
  const int mycol = (valid && a < W) ? ell_col[(size_t)j * W + a] : -1;
  int o0 = 0, o1 = 0;
  if (valid) {
    o0 = ovf_rowptr[j];
    o1 = ovf_rowptr[j + 1];
  }
  double2 bq[W];
  const double2 *bsrc = reinterpret_cast<const double2 *>(ell_val + (size_t)(valid ? j : 0) * W * 16);
#pragma unroll
  for (int k = 0; k < W; ++k) bq[k] = valid ? bsrc[k * 8 + a] : make_double2(0.0, 0.0);
  double2 *st2 = reinterpret_cast<double2 *>(stage);
#pragma unroll
  for (int k = 0; k < W; ++k) st2[k * 8 + a] = bq[k];
  DBG(W == 8 ? 1 : 5);
  // trip 2: the gathered rows
  double xi[W][4];
  int cols[W];
#pragma unroll
  for (int k = 0; k < W; ++k) {
    cols[k] = gshfl(mycol, k);
    ld4(Vsrc + (size_t)(cols[k] >= 0 ? cols[k] : 0) * 4 * r, r, a, act && cols[k] >= 0, xi[k]);
  }
  __syncwarp();
  if (xi[0][0] == 123.456) DBG(15);
  DBG(W == 8 ? 2 : 6);
#pragma unroll
  for (int k = 0; k < W; ++k)
    if (cols[k] >= 0) row_times_block(xi[k], stage + k * 16, acc);
  __syncwarp();
  if (acc[0] == 123.456) DBG(15);
  DBG(W == 8 ? 3 : 7);
  // overflow (poses with more than W blocks): plain CSR walk
  for (int e = o0; e < o1; ++e) {
    double xo[4];
    ld4(Vsrc + (size_t)ovf_col[e] * 4 * r, r, a, act, xo);
    row_times_block(xo, ovf_val + (size_t)e * 16, acc);
  }
}

constexpr int kStageStride = 8 * 16 + 4;  // doubles per group staging tile (+4: de-conflict the 4 groups of a warp)

// ---------------------------------------------------------------------------
// Cost / gradient (a3, a4): egrad = Xin Q + G, rgrad = Proj_Xin(egrad),
// f = 0.5 <Xin Q, Xin> + <G, Xin>.  G is (re)assembled from the inbox when
// build_g (a4: "G rebuilt every iteration", updateNeighborPoses :1276).
// Writes G (if build_g), S = sym(Y^T egrad_Y) per pose, Rg and its row-major
// copy RgT.  Accumulates partial f and |rgrad|^2.
// ---------------------------------------------------------------------------
template <int RC>
__device__ __forceinline__ void phase_grad(const AgentDev &A, const double *Xin, const double *inbox, bool build_g,
                                           double *Sout, double *Rgout, double *RgTout, double *egrad_out,
                                           double *stage_all, double &pf, double &pg2, int warp_lo = 0,
                                           int warp_cnt = kThreads / 32) {
  PoseIter it(warp_lo, warp_cnt);
  const int n = A.n, r = rdim<RC>(A);
  const size_t n4 = (size_t)4 * n;
  double *stage = stage_all + (size_t)it.lg * kStageStride;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], accq[4] = {0, 0, 0, 0}, accg[4] = {0, 0, 0, 0};
    DBG(0);
    ld4(Xin + off, r, it.a, act, x);
    if (!build_g) ld4(A.G + off, r, it.a, act, accg);
    ell_gather<8>(A.qe_col, A.qe_val, A.qo_rowptr, A.qo_col, A.qo_val, Xin, j, valid, r, it.a, act, stage, accq);
    if (build_g) {
      ell_gather<4>(A.se_slot, A.se_val, A.so_rowptr, A.so_slot, A.so_val, inbox, j, valid, r, it.a, act, stage,
                    accg);
      st4(A.G + off, r, it.a, act, accg);
    }
    double eg[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      eg[c] = accq[c] + accg[c];
      pf += (0.5 * accq[c] + accg[c]) * x[c];
    }
    if (egrad_out) st4(egrad_out + off, r, it.a, act, eg);
    DBG(8);
    const Sym3 S = sym_ytz(x, eg);
    sub_y_sym(x, S, eg);
    if (eg[0] == 123.456) DBG(15);
    DBG(9);
#pragma unroll
    for (int c = 0; c < 4; ++c) pg2 += eg[c] * eg[c];
    if (valid) {
      if (Sout && it.a == 0) {
        double *s = Sout + (size_t)j * 6;
        s[0] = S.a00; s[1] = S.a01; s[2] = S.a02; s[3] = S.a11; s[4] = S.a12; s[5] = S.a22;
      }
      if (Rgout) st4(Rgout + off, r, it.a, act, eg);
      if (RgTout && act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) RgTout[(size_t)it.a * n4 + 4 * j + c] = eg[c];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Riemannian Hessian-vector product at Xbase (a3/a5):
//   H = Proj_Xbase( V Q - V_Y * S ),  S = sym(Y^T egrad_Y) cached per pose.
// pvh accumulates <V, H>.
// ---------------------------------------------------------------------------
template <int RC>
__device__ __forceinline__ void phase_hess(const AgentDev &A, const double *Xbase, const double *S,
                                           const double *Vin, double *Hout, double *stage_all, double &pvh) {
  PoseIter it;
  const int n = A.n, r = rdim<RC>(A);
  double *stage = stage_all + (size_t)it.lg * kStageStride;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], v[4], h[4] = {0, 0, 0, 0};
    ld4(Xbase + off, r, it.a, act, x);
    ld4(Vin + off, r, it.a, act, v);
    Sym3 Sj = {0, 0, 0, 0, 0, 0};
    if (valid) {
      const double *s = S + (size_t)j * 6;
      Sj.a00 = s[0]; Sj.a01 = s[1]; Sj.a02 = s[2]; Sj.a11 = s[3]; Sj.a12 = s[4]; Sj.a22 = s[5];
    }
    ell_gather<8>(A.qe_col, A.qe_val, A.qo_rowptr, A.qo_col, A.qo_val, Vin, j, valid, r, it.a, act, stage, h);
    sub_y_sym(v, Sj, h);
    tangent_project_row(x, h);
#pragma unroll
    for (int c = 0; c < 4; ++c) pvh += v[c] * h[c];
    if (valid) st4(Hout + off, r, it.a, act, h);
  }
}

// ---------------------------------------------------------------------------
// Dense preconditioner (a6):  Z[:, cols] = V * Pinv[:, cols] for the columns
// of this CTA's poses.  Pinv is column-major with leading dimension ldp, so a
// CTA's slab (4 np consecutive columns) is ONE contiguous block -> one TMA bulk
// copy into shared memory.
// ---------------------------------------------------------------------------
struct SlabState {
  int agent;        // local agent whose first sub-chunk is resident / in flight (-1: none)
  unsigned parity;  // mbarrier phase parity of the next completion
  int pending;      // a copy was issued and not yet waited for
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *mbar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// thread 0: arm the mbarrier and issue the bulk copy (bytes: multiple of 16)
__device__ __forceinline__ void slab_issue(uint64_t *mbar, double *dst, const double *src, uint32_t bytes) {
  const uint32_t mb = smem_u32(mbar), d = smem_u32(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(src), "r"(bytes), "r"(mb)
               : "memory");
}
__device__ __forceinline__ void slab_wait(uint64_t *mbar, unsigned parity) {
  const uint32_t mb = smem_u32(mbar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mb), "r"(parity)
        : "memory");
  }
}

// this CTA's balanced share of an agent's poses for the dense phase
__device__ __forceinline__ void cta_pose_chunk(int n, int &p0, int &np) {
  const int base = n / (int)gridDim.x, rem = n % (int)gridDim.x;
  const int b = (int)blockIdx.x;
  np = base + (b < rem ? 1 : 0);
  p0 = b * base + min(b, rem);
}
__device__ __forceinline__ size_t agent_ldp(const AgentDev &A) { return ((size_t)4 * A.n + 31) / 32 * 32; }
// poses per resident sub-chunk for this agent (0: the slab does not fit -> global-load fallback)
__device__ __forceinline__ int slab_poses(const AgentDev &A, int np, size_t slab_cap_bytes) {
  const size_t per = (size_t)4 * agent_ldp(A) * sizeof(double);
  return min(np, (int)(slab_cap_bytes / per));
}

// Prefetch the first sub-chunk of `A`'s slab (called by all threads of the CTA;
// the previous contents of the buffer must no longer be needed).
__device__ __forceinline__ void slab_prefetch(const AgentDev &A, int ai, SlabState &ss, uint64_t *mbar, double *slab,
                                              size_t slab_cap_bytes) {
  if (ss.agent == ai || A.Pinv == nullptr) return;
  int p0, np;
  cta_pose_chunk(A.n, p0, np);
  const int pps = slab_poses(A, np, slab_cap_bytes);
  if (ss.pending) {  // never leave an unobserved completion behind
    slab_wait(mbar, ss.parity);
    ss.parity ^= 1;
    ss.pending = 0;
  }
  ss.agent = -1;
  if (pps <= 0) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t ldp = agent_ldp(A);
    slab_issue(mbar, slab, A.Pinv + (size_t)4 * p0 * ldp, (uint32_t)((size_t)4 * pps * ldp * sizeof(double)));
  }
  ss.agent = ai;
  ss.pending = 1;
}

// columns (poses x 4) one thread accumulates per pass: 12 while the R x 12 accumulator fits the
// register file next to the prefetched V^T values, 8 for the larger ranks
template <int R>
struct DensePassCols {
  static constexpr int value = (R <= 6) ? 12 : 8;
};

// acc over one pass of <= DensePassCols/4 poses whose columns start at `cols`
// (shared memory or global, column stride ld); result to zs[pose][c][8].
// STREAM: `cols` is global memory read once (agents whose slab does not fit shared memory)
// AXPY: the left operand is VT + alpha * HT, formed while it is loaded (tCG: r+ = r + alpha H[delta] without a pass
// of its own and the grid barrier that would have to follow it)
template <int R, bool STREAM = false, bool AXPY = false>
__device__ __forceinline__ void dense_pass(const double *cols, size_t ld, const double *VT, int r, int n4, int npass,
                                           double *zs, double *red /* [8 warps][16 cols][8] */,
                                           const double *HT = nullptr, double alpha = 0.0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ncols = npass * 4;
  constexpr int NC = DensePassCols<R>::value;  // columns held per thread (register budget)
  double acc[R][NC];
#pragma unroll
  for (int a = 0; a < R; ++a)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[a][c] = 0.0;
  // q-strided accumulation; the V^T values of a whole chunk of q-iterations are requested up front
  // (one L2 round trip per chunk instead of one per iteration)
  constexpr int QC = (R <= 5) ? 3 : 2;
  for (int q0 = threadIdx.x; q0 < n4; q0 += kThreads * QC) {
    double vr[QC][R];
#pragma unroll
    for (int i = 0; i < QC; ++i) {
      const int q = q0 + kThreads * i;
#pragma unroll
      for (int a = 0; a < R; ++a) vr[i][a] = (a < r && q < n4) ? VT[(size_t)a * n4 + q] : 0.0;
      if constexpr (AXPY) {
        double hr[R];
#pragma unroll
        for (int a = 0; a < R; ++a) hr[a] = (a < r && q < n4) ? HT[(size_t)a * n4 + q] : 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) vr[i][a] = fma(alpha, hr[a], vr[i][a]);
      }
    }
    if constexpr (STREAM) {
      // columns streamed from HBM: request the whole chunk (QC x NC independent loads per thread) before the first
      // multiply -- the bytes in flight per SM are what bounds this pass
      double pv[QC][NC];
#pragma unroll
      for (int i = 0; i < QC; ++i) {
        const int q = q0 + kThreads * i;
#pragma unroll
        for (int c = 0; c < NC; ++c) pv[i][c] = (q < n4 && c < ncols) ? cols[(size_t)c * ld + q] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < QC; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int a = 0; a < R; ++a) acc[a][c] = fma(vr[i][a], pv[i][c], acc[a][c]);
    } else {
#pragma unroll
      for (int i = 0; i < QC; ++i) {
        const int q = q0 + kThreads * i;
        if (q < n4) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const double pv = (c < ncols) ? cols[(size_t)c * ld + q] : 0.0;
#pragma unroll
            for (int a = 0; a < R; ++a) acc[a][c] = fma(vr[i][a], pv, acc[a][c]);
          }
        }
      }
    }
  }
  if (acc[0][0] == 123.456) DBG(31);
  DBG(17);
  // reduce-scatter over 16 (12 + 4 zero) columns: xor 1, 2, 4, 8 then butterfly 16
  double h8[R][8], h4[R][4], h2[R][2], h1[R];
  {
    const bool hi = lane & 1;
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double up = (c + 8 < NC) ? acc[a][c + 8 < NC ? c + 8 : 0] : 0.0;
        const double keep = hi ? up : acc[a][c];
        const double send = hi ? acc[a][c] : up;
        h8[a][c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
  }
  {
    const bool hi = lane & 2;
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double keep = hi ? h8[a][c + 4] : h8[a][c];
        const double send = hi ? h8[a][c] : h8[a][c + 4];
        h4[a][c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const double keep = hi ? h4[a][c + 2] : h4[a][c];
        const double send = hi ? h4[a][c] : h4[a][c + 2];
        h2[a][c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int a = 0; a < R; ++a) {
      const double keep = hi ? h2[a][1] : h2[a][0];
      const double send = hi ? h2[a][0] : h2[a][1];
      double v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      h1[a] = v;
    }
  }
  if (h1[0] == 123.456) DBG(31);
  DBG(18);
  // lane l < 16 holds column ((l&1)<<3 | (l&2)<<1 | (l&4)>>1 | (l&8)>>3)
  __syncthreads();  // red reuse
  if (lane < 16) {
    const int col = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);
#pragma unroll
    for (int a = 0; a < R; ++a) red[(warp * 16 + col) * 8 + a] = h1[a];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int col = threadIdx.x >> 3, a = threadIdx.x & 7;
    if (a < r && col < ncols) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) s += red[(w * 16 + col) * 8 + a];
      zs[((col >> 2) * 4 + (col & 3)) * 8 + a] = s;
    }
  }
}

// Z slab of this CTA's chunk [p0, p0+np) into zs[pose_local][c][8].
// BIG: instantiated for teams with an agent whose slab does not fit shared memory; only those kernels carry the
// register-hungry streaming pass (it costs the small-agent kernels 4 us per iteration in spills otherwise)
template <int R, bool BIG = false, bool AXPY = false>
__device__ __forceinline__ void dense_slab(const AgentDev &A, int ai, const double *VT, int p0, int np,
                                           SlabState &ss, uint64_t *mbar, double *slab, size_t slab_cap_bytes,
                                           double *zs, double *red, const double *HT = nullptr, double alpha = 0.0) {
  const int r = rdim<R>(A), n4 = 4 * A.n;
  const size_t ldp = agent_ldp(A);
  const int pps = slab_poses(A, np, slab_cap_bytes);
  // BIG kernels stream whenever the CTA's whole chunk does not fit shared memory: re-filling the buffer pose by pose
  // is load-then-compute (1.3 TB/s at n = 1250), the streaming pass keeps QC x NC loads per thread in flight
  if (BIG ? pps < np : pps <= 0) {
    // slab larger than shared memory (BASELINE config 5: n = 12 500 poses, 136 MB of columns per CTA): stream the
    // columns from global memory.  This is the HBM-bound regime -- Pinv is read exactly once per application.
    // With all 36 loads of a chunk requested before the first multiply (dense_pass<R, true>) the pass runs at
    // 4.6 TB/s; left to the compiler's schedule it reached 3.3 - 3.7 TB/s (2.0 in the parallel-schedule kernel).
    // (Two shared-memory ring variants, fed by per-column TMA bulk copies and by 16-byte cp.async, measured 1.6
    // and 2.3 TB/s: the ring's per-tile barriers cost more than the registers it frees.)
    constexpr int NPP = DensePassCols<R>::value / 4;
    for (int sub = 0; sub < np; sub += NPP)
      dense_pass<R, BIG, AXPY>(A.Pinv + (size_t)4 * (p0 + sub) * ldp, ldp, VT, r, n4, min(NPP, np - sub), zs + sub * 32,
                               red, HT, alpha);
    __syncthreads();
    return;
  }
  for (int s0 = 0; s0 < np; s0 += pps) {
    const int cnt = min(pps, np - s0);
    if (s0 == 0 && ss.agent == ai) {
      if (ss.pending) {
        slab_wait(mbar, ss.parity);
        ss.parity ^= 1;
        ss.pending = 0;
      }
    } else {
      if (ss.pending) {
        slab_wait(mbar, ss.parity);
        ss.parity ^= 1;
        ss.pending = 0;
      }
      __syncthreads();  // everyone is done with the previous contents
      if (threadIdx.x == 0)
        slab_issue(mbar, slab, A.Pinv + (size_t)4 * (p0 + s0) * ldp, (uint32_t)((size_t)4 * cnt * ldp * sizeof(double)));
      slab_wait(mbar, ss.parity);
      ss.parity ^= 1;
      ss.agent = (s0 == 0) ? ai : -1;
    }
    DBG(16);
    constexpr int NPP = DensePassCols<R>::value / 4;
    for (int sub = 0; sub < cnt; sub += NPP)
      dense_pass<R, false, AXPY>(slab + (size_t)4 * sub * ldp, ldp, VT, r, n4, min(NPP, cnt - sub),
                                 zs + (s0 + sub) * 32, red, HT, alpha);
  }
  if (np > pps) ss.agent = -1;  // the buffer no longer holds sub-chunk 0
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Commit a new iterate for pose j (group-collective): relative-change partial,
// X <- xnew (and an optional copy for the deferred statistics pass), publish,
// and the Nesterov V update (a7):
//   V = proj(V + gamma (X+ - Y))    or, on restart iterations, V = Y = X+.
// ---------------------------------------------------------------------------
template <int RC>
__device__ __forceinline__ void finish_pose(const AgentDev &A, int j, bool valid, int a, const double (&xnew)[4],
                                            bool accel, bool restart, double gamma, double *xcopy, double &prel) {
  const int r = rdim<RC>(A);
  const bool act = valid && a < r;
  const size_t off = (size_t)(valid ? j : 0) * 4 * r;
  double xold[4], y[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
  int pe0 = 0, pe1 = 0;
  if (valid) {
    pe0 = A.pub_rowptr[j];
    pe1 = A.pub_rowptr[j + 1];
  }
  ld4(A.X + off, r, a, act, xold);
  if (accel && !restart) {
    ld4(A.Y + off, r, a, act, y);
    ld4(A.V + off, r, a, act, v);
  }
  const PubPtrs preg = pub_prefetch(A.pub_dst_reg, pe0, pe1);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const double d = xnew[c] - xold[c];
    prel += act ? d * d : 0.0;
  }
  if (valid) {
    st4(A.X + off, r, a, act, xnew);
    if (xcopy) st4(xcopy + off, r, a, act, xnew);
    publish_pre(preg, A.pub_dst_reg, r, a, act, xnew);
  }
  if (accel) {
    if (restart) {
      if (valid) {
        st4(A.V + off, r, a, act, xnew);
        st4(A.Y + off, r, a, act, xnew);
        publish_range(pe0, pe1, A.pub_dst_aux, r, a, act, xnew);
      }
    } else {
      double m[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) m[c] = v[c] + gamma * (xnew[c] - y[c]);
      if (!valid) {
        m[0] = (a == 0);
        m[1] = (a == 1);
        m[2] = (a == 2);
      }
      stiefel_project_row(m);
      if (valid) st4(A.V + off, r, a, act, m);
    }
  }
}

// ---------------------------------------------------------------------------
// RGD step (a2, src/PGOAgentROSNode.cpp:96-97):
//   X+ = Retr_Xs( -eta * Proj_Xs( P^-1 rgrad ) ), fused with finish_pose.
// With the preconditioner the CTA first computes its dense slab.
// ---------------------------------------------------------------------------
// ext_zt != nullptr: Z^T = (Rg Pinv)^T was computed in front of this launch (sym_precond.cu), [r][4n] row-major
template <int R, bool BIG = false>
__device__ __forceinline__ void phase_rgd_step(const AgentDev &A, int ai, const SolverParams &P, const double *Xs,
                                               bool accel, bool restart, double gamma, SlabState &ss,
                                               uint64_t *mbar, double *slab, size_t slab_cap, double *zs,
                                               double *red, double *xcopy, double &prel,
                                               const double *ext_zt = nullptr) {
  const int n = A.n, r = rdim<R>(A);
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  if (P.rgd_use_precond) {
    int p0, np;
    cta_pose_chunk(n, p0, np);
    // warm L1 with what the epilogue of this phase reads (state of my poses, publication ranges)
    if (lg < np && a < 2) {
      const size_t off = (size_t)(p0 + lg) * 4 * r + a * 16;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.X + off));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.V + off));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.Y + off));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.pub_rowptr + p0 + lg));
    }
    DBG(20);
    // poses of this CTA's chunk are processed by group k (same ownership as phase_nesterov_chunk)
    if (!ext_zt) dense_slab<R, BIG>(A, ai, A.RgT, p0, np, ss, mbar, slab, slab_cap, zs, red);
    DBG(19);
    for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
      const int k = k0 + lg;
      const bool valid = k < np;
      const int j = p0 + (valid ? k : 0);
      const bool act = valid && a < r;
      double y[4], z[4];
      ld4(Xs + (size_t)j * 4 * r, r, a, act, y);
      if (ext_zt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) z[c] = act ? ext_zt[(size_t)a * 4 * n + 4 * j + c] : 0.0;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) z[c] = act ? zs[((valid ? k : 0) * 4 + c) * 8 + a] : 0.0;
      }
      tangent_project_row(y, z);
      if (z[0] == 123.456) DBG(31);
      DBG(21);
      double xn[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) xn[c] = y[c] - P.rgd_stepsize * z[c];
      if (!valid) {
        xn[0] = (a == 0); xn[1] = (a == 1); xn[2] = (a == 2);
      }
      qf_row(xn);
      if (xn[0] == 123.456) DBG(31);
      DBG(22);
      finish_pose<R>(A, j, valid, a, xn, accel, restart, gamma, xcopy, prel);
      DBG(23);
    }
  } else {
    PoseIter it;
    int j;
    while (it.next(n, j)) {
      const bool valid = j < n;
      const bool act = valid && it.a < r;
      const size_t off = (size_t)(valid ? j : 0) * 4 * r;
      double y[4], z[4], xn[4];
      ld4(Xs + off, r, it.a, act, y);
      ld4(A.Rg + off, r, it.a, act, z);
#pragma unroll
      for (int c = 0; c < 4; ++c) xn[c] = y[c] - P.rgd_stepsize * z[c];
      if (!valid) {
        xn[0] = (it.a == 0); xn[1] = (it.a == 1); xn[2] = (it.a == 2);
      }
      qf_row(xn);
      finish_pose<R>(A, valid ? j : 0, valid, it.a, xn, accel, restart, gamma, xcopy, prel);
    }
  }
}

// Z = Proj_Xbase( V Pinv ) for the whole agent (tCG preconditioner, a6).  Writes
// Z and optionally dlt = -Z; pzr accumulates <Z, Rin>.
template <int R, bool BIG = false>
__device__ __forceinline__ void phase_precond(const AgentDev &A, int ai, const double *Xbase, const double *Rin,
                                              const double *RinT, double *Zout, double *neg_out, SlabState &ss,
                                              uint64_t *mbar, double *slab, size_t slab_cap, double *zs,
                                              double *red, double &pzr) {
  const int n = A.n, r = rdim<R>(A);
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  int p0, np;
  cta_pose_chunk(n, p0, np);
  dense_slab<R, BIG>(A, ai, RinT, p0, np, ss, mbar, slab, slab_cap, zs, red);
  for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
    const int k = k0 + lg;
    const bool valid = k < np;
    const int j = p0 + (valid ? k : 0);
    const bool act = valid && a < r;
    const size_t off = (size_t)j * 4 * r;
    double y[4], z[4], rr[4];
    ld4(Xbase + off, r, a, act, y);
    ld4(Rin + off, r, a, act, rr);
#pragma unroll
    for (int c = 0; c < 4; ++c) z[c] = act ? zs[((valid ? k : 0) * 4 + c) * 8 + a] : 0.0;
    tangent_project_row(y, z);
#pragma unroll
    for (int c = 0; c < 4; ++c) pzr += z[c] * rr[c];
    if (valid) {
      st4(Zout + off, r, a, act, z);
      if (neg_out) {
        double nz[4] = {-z[0], -z[1], -z[2], -z[3]};
        st4(neg_out + off, r, a, act, nz);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// tCG with TWO grid-wide synchronisations per inner iteration (team_run.cuh, rtr_solve).
//
// The textbook loop needs four: <delta, H delta> after the Hessian-vector product, |r+|^2 after the residual update,
// <z+, r+> after the preconditioner, and a barrier after the new direction.  Here
//   * the Hessian-vector phase forms delta+ = -z + beta delta ON THE FLY for the pose it owns and for every pose it
//     gathers (z and the previous delta are both complete by then), so the direction never needs a pass and a
//     barrier of its own;
//   * the preconditioner phase reads r+ = r + alpha H[delta] on the fly (dense_pass<.., AXPY>), and its epilogue
//     -- chunk-owned poses -- stores r+, advances eta and accumulates |r+|^2 and <z+, r+> for ONE joint reduction.
// The arithmetic of every vector is the one of the four-barrier loop, operation for operation (the preconditioner
// application of the iteration that turns out to be the last is speculative and discarded).
// ---------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void ell_gather_dir(const int *ell_col, const double *ell_val, const int *ovf_rowptr,
                                               const int *ovf_col, const double *ovf_val, const double *Zsrc,
                                               const double *Dprev, double beta, int j, bool valid, int r, int a,
                                               bool act, double *stage, double (&acc)[4]) {
  const int mycol = (valid && a < W) ? ell_col[(size_t)j * W + a] : -1;
  int o0 = 0, o1 = 0;
  if (valid) {
    o0 = ovf_rowptr[j];
    o1 = ovf_rowptr[j + 1];
  }
  double2 bq[W];
  const double2 *bsrc = reinterpret_cast<const double2 *>(ell_val + (size_t)(valid ? j : 0) * W * 16);
#pragma unroll
  for (int k = 0; k < W; ++k) bq[k] = valid ? bsrc[k * 8 + a] : make_double2(0.0, 0.0);
  double2 *st2 = reinterpret_cast<double2 *>(stage);
#pragma unroll
  for (int k = 0; k < W; ++k) st2[k * 8 + a] = bq[k];
  int cols[W];
#pragma unroll
  for (int k = 0; k < W; ++k) cols[k] = gshfl(mycol, k);
  __syncwarp();
  // two half-width batches: 2 x 4 x 4 loads in flight per lane (z and the previous direction of four poses)
#pragma unroll
  for (int h = 0; h < W; h += 4) {
    double zi[4][4], di[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool on = act && cols[h + k] >= 0;
      const size_t off = (size_t)(cols[h + k] >= 0 ? cols[h + k] : 0) * 4 * r;
      ld4(Zsrc + off, r, a, on, zi[k]);
      ld4(Dprev + off, r, a, on, di[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (cols[h + k] >= 0) {
        double xi[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xi[c] = -zi[k][c] + beta * di[k][c];
        row_times_block(xi, stage + (h + k) * 16, acc);
      }
  }
  __syncwarp();
  for (int e = o0; e < o1; ++e) {
    double zo[4], dp[4], xo[4];
    const size_t off = (size_t)ovf_col[e] * 4 * r;
    ld4(Zsrc + off, r, a, act, zo);
    ld4(Dprev + off, r, a, act, dp);
#pragma unroll
    for (int c = 0; c < 4; ++c) xo[c] = -zo[c] + beta * dp[c];
    row_times_block(xo, ovf_val + (size_t)e * 16, acc);
  }
}

// H[delta] with delta given (first inner iteration: Dcur already holds -z0) or formed on the fly as
// -Z + beta Dprev (and then stored to Dcur for the pose this group owns).  Writes Hd and its row-major copy HdT;
// pvh accumulates <delta, H delta>.
template <int RC>
__device__ __forceinline__ void phase_hess_dir(const AgentDev &A, const double *Xbase, const double *S, bool first,
                                               const double *Zsrc, const double *Dprev, double beta, double *Dcur,
                                               double *Hout, double *HoutT, double *stage_all, double &pvh) {
  PoseIter it;
  const int n = A.n, r = rdim<RC>(A);
  const size_t n4 = (size_t)4 * n;
  double *stage = stage_all + (size_t)it.lg * kStageStride;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], v[4], h[4] = {0, 0, 0, 0};
    ld4(Xbase + off, r, it.a, act, x);
    Sym3 Sj = {0, 0, 0, 0, 0, 0};
    if (valid) {
      const double *s = S + (size_t)j * 6;
      Sj.a00 = s[0]; Sj.a01 = s[1]; Sj.a02 = s[2]; Sj.a11 = s[3]; Sj.a12 = s[4]; Sj.a22 = s[5];
    }
    if (first) {
      ld4(Dcur + off, r, it.a, act, v);
      ell_gather<8>(A.qe_col, A.qe_val, A.qo_rowptr, A.qo_col, A.qo_val, Dcur, j, valid, r, it.a, act, stage, h);
    } else {
      double z[4], d[4];
      ld4(Zsrc + off, r, it.a, act, z);
      ld4(Dprev + off, r, it.a, act, d);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = -z[c] + beta * d[c];
      if (valid) st4(Dcur + off, r, it.a, act, v);
      ell_gather_dir<8>(A.qe_col, A.qe_val, A.qo_rowptr, A.qo_col, A.qo_val, Zsrc, Dprev, beta, j, valid, r, it.a,
                        act, stage, h);
    }
    sub_y_sym(v, Sj, h);
    tangent_project_row(x, h);
#pragma unroll
    for (int c = 0; c < 4; ++c) pvh += v[c] * h[c];
    if (valid) {
      st4(Hout + off, r, it.a, act, h);
      if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) HoutT[(size_t)it.a * n4 + 4 * j + c] = h[c];
      }
    }
  }
}

// z+ = Proj_Xbase( (r + alpha Hd) Pinv ) for the whole agent, and for the poses of this CTA's chunk:
//   r+ = r + alpha Hd  (stored column- and row-major),  eta (+)= alpha delta,
//   prr += |r+|^2,  pzr += <z+, r+>,
//   cand = Retr_Xbase(eta)  -- the candidate of the outer iteration, should tCG stop after this iteration (it does so
//   on the residual test the reduction behind this phase feeds): a retraction per pose and inner iteration buys the
//   barrier + retraction phase + barrier the outer iteration would otherwise spend after tCG.
template <int R, bool BIG = false>
__device__ __forceinline__ void phase_precond_cg(const AgentDev &A, int ai, const double *Xbase, const double *Rin,
                                                 const double *RinT, const double *Hd, const double *HdT,
                                                 double alpha, const double *Dcur, bool eta_zero, double *eta,
                                                 double *cand, double *Rout, double *RoutT, double *Zout, SlabState &ss,
                                                 uint64_t *mbar, double *slab, size_t slab_cap, double *zs,
                                                 double *red, double &prr, double &pzr) {
  const int n = A.n, r = rdim<R>(A);
  const size_t n4 = (size_t)4 * n;
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  int p0, np;
  cta_pose_chunk(n, p0, np);
  dense_slab<R, BIG, true>(A, ai, RinT, p0, np, ss, mbar, slab, slab_cap, zs, red, HdT, alpha);
  for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
    const int k = k0 + lg;
    const bool valid = k < np;
    const int j = p0 + (valid ? k : 0);
    const bool act = valid && a < r;
    const size_t off = (size_t)j * 4 * r;
    double y[4], z[4], rr[4], hd[4], d[4], e[4];
    ld4(Xbase + off, r, a, act, y);
    ld4(Rin + off, r, a, act, rr);
    ld4(Hd + off, r, a, act, hd);
    ld4(Dcur + off, r, a, act, d);
    if (eta_zero) {
      e[0] = e[1] = e[2] = e[3] = 0.0;
    } else {
      ld4(eta + off, r, a, act, e);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) z[c] = act ? zs[((valid ? k : 0) * 4 + c) * 8 + a] : 0.0;
    tangent_project_row(y, z);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      rr[c] = fma(alpha, hd[c], rr[c]);   // the same operation dense_pass<.., AXPY> applied to its operand
      e[c] += alpha * d[c];
      prr += rr[c] * rr[c];
      pzr += z[c] * rr[c];
    }
    if (valid) {
      st4(Zout + off, r, a, act, z);
      st4(Rout + off, r, a, act, rr);
      st4(eta + off, r, a, act, e);
      if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c) RoutT[(size_t)a * n4 + 4 * j + c] = rr[c];
      }
    }
    {  // the same operations, in the same order, as phase_retract (team_run.cuh)
      double xc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) xc[c] = y[c] + e[c];
      if (!valid) {
        xc[0] = (a == 0); xc[1] = (a == 1); xc[2] = (a == 2);
      }
      qf_row(xc);
      if (valid) st4(cand + off, r, a, act, xc);
    }
  }
}

}  // namespace dpgo
