// The persistent cooperative kernel k_team_run<R> and its launcher; instantiated once per
// relaxation rank in team_run_r<R>.cu so the instantiations compile in parallel.
#pragma once
#include <algorithm>
#include <atomic>

#include "kernels.h"
#include "phases.cuh"

namespace dpgo {

void count_launch();

// dynamic shared memory layout of the persistent kernel:
//   [ slab (slab_cap bytes) | staging tiles (32 groups) | zs (chunk poses x 32) ]
struct SmemLayout {
  double *slab;
  size_t slab_cap;
  double *stage;
  double *zs;
};

// ---------------------------------------------------------------------------
// pose-local vector phases used by tCG
// ---------------------------------------------------------------------------
// eta (+)= tau * dlt  (trust-region boundary / negative curvature exit)
template <int RC>
__device__ __forceinline__ void phase_axpy_eta(const AgentDev &A, double tau, bool eta_zero, const double *dlt,
                                               double *eta) {
  PoseIter it;
  const int n = A.n, r = rdim<RC>(A);
  int j;
  while (it.next(n, j)) {
    const bool act = j < n && it.a < r;
    if (!act) continue;
    const size_t off = (size_t)j * 4 * r;
    double d[4], e[4];
    ld4(dlt + off, r, it.a, act, d);
    if (eta_zero) {
      e[0] = e[1] = e[2] = e[3] = 0.0;
    } else {
      ld4(eta + off, r, it.a, act, e);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) e[c] += tau * d[c];
    st4(eta + off, r, it.a, act, e);
  }
}

// eta (+)= tau * dlt and cand = Retr_x1(eta) for the poses of this CTA's chunk -- the owner of eta in phase_precond_cg, so
// no barrier separates the two (dlt comes from the Hessian-vector phase, behind a grid reduction)
template <int RC>
__device__ __forceinline__ void phase_axpy_eta_retract(const AgentDev &A, double tau, bool eta_zero, const double *dlt,
                                                       double *eta, const double *X1, double *cand) {
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  const int r = rdim<RC>(A);
  int p0, np;
  cta_pose_chunk(A.n, p0, np);
  for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
    const int k = k0 + lg;
    const bool valid = k < np;
    const int j = p0 + (valid ? k : 0);
    const bool act = valid && a < r;
    const size_t off = (size_t)j * 4 * r;
    double d[4], e[4], x[4];
    ld4(dlt + off, r, a, act, d);
    ld4(X1 + off, r, a, act, x);
    if (eta_zero) {
      e[0] = e[1] = e[2] = e[3] = 0.0;
    } else {
      ld4(eta + off, r, a, act, e);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) e[c] += tau * d[c];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] += e[c];
    if (!valid) {
      x[0] = (a == 0); x[1] = (a == 1); x[2] = (a == 2);
    }
    qf_row(x);
    if (valid) {
      st4(eta + off, r, a, act, e);
      st4(cand + off, r, a, act, x);
    }
  }
}

// out = Retr_x(eta)  (group-collective)
template <int RC>
__device__ __forceinline__ void phase_retract(const AgentDev &A, const double *X1, const double *eta, double *out) {
  PoseIter it;
  const int n = A.n, r = rdim<RC>(A);
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], e[4];
    ld4(X1 + off, r, it.a, act, x);
    ld4(eta + off, r, it.a, act, e);
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] += e[c];
    if (!valid) {
      x[0] = (it.a == 0); x[1] = (it.a == 1); x[2] = (it.a == 2);
    }
    qf_row(x);
    if (valid) st4(out + off, r, it.a, act, x);
  }
}

template <int RC>
__device__ __forceinline__ void phase_dot(const AgentDev &A, const double *U, const double *W, double &p) {
  PoseIter it;
  const int n = A.n, r = rdim<RC>(A);
  int j;
  while (it.next(n, j)) {
    const bool act = j < n && it.a < r;
    if (!act) continue;
    const size_t off = (size_t)j * 4 * r;
    double u[4], w[4];
    ld4(U + off, r, it.a, act, u);
    ld4(W + off, r, it.a, act, w);
#pragma unroll
    for (int c = 0; c < 4; ++c) p += u[c] * w[c];
  }
}

// commit x1 as the agent's new X (RTR epilogue); chunk ownership like every writer of X / V / Y
template <int RC>
__device__ __forceinline__ void phase_commit(const AgentDev &A, const double *X1, bool accel, bool restart,
                                             double gamma, double &prel) {
  const int a = threadIdx.x & 7, lg = threadIdx.x >> 3;
  int p0, np;
  cta_pose_chunk(A.n, p0, np);
  for (int k0 = 0; k0 < np; k0 += kGroupsPerCta) {
    const int k = k0 + lg;
    const bool valid = k < np;
    const int j = p0 + (valid ? k : 0);
    const int r = rdim<RC>(A);
    const bool act = valid && a < r;
    double xn[4];
    ld4(X1 + (size_t)j * 4 * r, r, a, act, xn);
    finish_pose<RC>(A, j, valid, a, xn, accel, restart, gamma, nullptr, prel);
  }
}

// ---------------------------------------------------------------------------
// RTR-tCG local solve (a2), ROPTLIB RTRNewton semantics as restated in
// oracle/dpgo_oracle.cpp (rtrRun): every scalar decision is taken redundantly
// by all threads from bit-identical reduced values.  The preconditioner slab
// of this CTA stays resident in shared memory for the whole solve.
// ---------------------------------------------------------------------------
struct RtrOut {
  const double *x;  // final iterate
  double f_init, gn_init, f_opt, gn_opt;
  int outer, tcg, rej;
};

template <int R, bool BIG = false>
__device__ __forceinline__ RtrOut rtr_solve(const AgentDev &A, int ai, const SolverParams &P, const GridSync &gs,
                                            BarState &bs, const double *Xs, const double *inbox, SlabState &ss,
                                            uint64_t *mbar, const SmemLayout &L, double *red, double *sm) {
  RtrOut out;
  out.outer = out.tcg = out.rej = 0;
  int dseq = 0;  // diagnostics: clock64 of CTA 0 after every phase / reduction of the first solve of a profiled launch
#define RMARK() do { if (g_dbg && threadIdx.x == 0 && blockIdx.x == 0 && dseq < 160) g_dbg[64 + dseq] = clock64(); ++dseq; } while (0)
  RMARK();
  const double *x1 = Xs;
  double *cand = A.X2;
  double *Rg1 = A.Rg, *Rg1T = A.RgT, *S1 = A.S;
  double *Rg2 = A.Rg2, *Rg2T = A.Rg2T, *S2 = A.S2;
  double v[4];
  // gradient at the starting point (also assembles G)
  v[0] = v[1] = v[2] = v[3] = 0;
  phase_grad<R>(A, x1, inbox, true, S1, Rg1, Rg1T, nullptr, L.stage, v[0], v[1]);
  RMARK();
  grid_reduce<2>(gs, bs, reinterpret_cast<double(&)[2]>(v), sm);
  RMARK();
  double f1 = v[0], ngf = sqrt(v[1]);
  out.f_init = f1;
  out.gn_init = ngf;
  const bool single = (P.rtr_iterations == 1);
  double Delta = P.rtr_initial_radius;
  double maxDelta = single ? Delta : 5.0 * P.rtr_initial_radius;
  int iter = 0, shrink = 0;
  // ROPTLIB evaluates the stopping criterion before the first iteration: a start whose gradient norm already meets
  // the tolerance is returned untouched (same rule in oracle rtrRun)
  bool stop = ngf < P.gradnorm_tol;
  const double theta = 1.0, kappa = 0.1;
  while (true) {
    if (stop && iter == 0) break;
    if (!single && (stop || iter >= P.rtr_iterations)) break;
    // ---------------- tCG
    const double *rsrc = Rg1, *rsrcT = Rg1T;
    const double norm_r0 = ngf;
    v[0] = 0;
    phase_precond<R, BIG>(A, ai, x1, rsrc, rsrcT, A.Z, A.dlt0, ss, mbar, L.slab, L.slab_cap, L.zs, red, v[0]);
    RMARK();
    grid_reduce<1>(gs, bs, reinterpret_cast<double(&)[1]>(v), sm);
    RMARK();
    double z_r = v[0], d_Pd = z_r, e_Pe = 0.0, e_Pd = 0.0;
    bool eta_zero = true;
    int status = 4;  // 0 negcurv, 1 exceeded, 2 lcon, 3 scon, 4 maxiter
    int j = 0;
    // two grid-wide synchronisations per inner iteration (phases.cuh: phase_hess_dir / phase_precond_cg)
    double beta = 0.0;
    double *dcur = A.dlt0, *dprev = A.dlt1;   // delta_j (materialised by the Hessian-vector phase) / delta_{j-1}
    double *rnext = A.rv, *rnextT = A.rvT;    // ping-pong target of r+ (never the buffer other CTAs still read)
    bool cand_pending = false;
    for (j = 0; j < P.rtr_tcg_iterations; ++j) {
      v[0] = 0;
      phase_hess_dir<R>(A, x1, S1, j == 0, A.Z, dprev, beta, dcur, A.Hd, A.HdT, L.stage, v[0]);
      RMARK();
      grid_reduce<1>(gs, bs, reinterpret_cast<double(&)[1]>(v), sm);
      RMARK();
      const double d_Hd = v[0];
      const double alpha = z_r / d_Hd;
      const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
      if (d_Hd <= 0 || e_Pe_new >= Delta * Delta) {
        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
        phase_axpy_eta_retract<R>(A, tau, eta_zero, dcur, A.eta, x1, cand);
        eta_zero = false;
        status = (d_Hd <= 0) ? 0 : 1;
        cand_pending = true;   // written after the last grid-wide synchronisation
        break;
      }
      e_Pe = e_Pe_new;
      v[0] = v[1] = 0;
      phase_precond_cg<R, BIG>(A, ai, x1, rsrc, rsrcT, A.Hd, A.HdT, alpha, dcur, eta_zero, A.eta, cand, rnext, rnextT, A.Z,
                               ss, mbar, L.slab, L.slab_cap, L.zs, red, v[0], v[1]);
      eta_zero = false;
      RMARK();
      grid_reduce<2>(gs, bs, reinterpret_cast<double(&)[2]>(v), sm);
      RMARK();
      rsrc = rnext;
      rsrcT = rnextT;
      rnext = (rnext == A.rv) ? A.rw : A.rv;
      rnextT = (rnextT == A.rvT) ? A.rwT : A.rvT;
      const double norm_r = sqrt(v[0]);
      const double tempnum = pow(norm_r0, theta);
      if (norm_r <= norm_r0 * fmin(tempnum, kappa)) {
        status = (kappa < tempnum) ? 2 : 3;
        break;
      }
      const double zold_rold = z_r;
      z_r = v[1];
      beta = z_r / zold_rold;
      {  // the next Hessian-vector phase forms delta_{j+1} = -z + beta delta_j while it gathers
        double *t = dcur;
        dcur = dprev;
        dprev = t;
      }
      e_Pd = beta * (e_Pd + alpha * d_Pd);
      d_Pd = z_r + beta * beta * d_Pd;
    }
    out.tcg += min(j + 1, P.rtr_tcg_iterations);
    // ---------------- candidate, model decrease, ratio
    // eta and cand = Retr_x1(eta) are already there: phase_precond_cg keeps the candidate of "tCG stops here" up to date
    // (complete behind its grid reduction), the boundary / negative-curvature exit writes both in one chunk-owned pass.
    if (eta_zero) {  // maxInner == 0: eta = 0
      phase_axpy_eta<R>(A, 0.0, true, A.dlt0, A.eta);
      grid_barrier(gs, bs);
      phase_retract<R>(A, x1, A.eta, cand);
      cand_pending = true;
    }
    RMARK();
    if (cand_pending) grid_barrier(gs, bs);
    RMARK();
    v[0] = v[1] = v[2] = v[3] = 0;
    phase_grad<R>(A, cand, inbox, false, S2, Rg2, Rg2T, nullptr, L.stage, v[0], v[1]);
    RMARK();
    phase_hess<R>(A, x1, S1, A.eta, A.zeta, L.stage, v[2]);
    RMARK();
    phase_dot<R>(A, A.eta, Rg1, v[3]);
    RMARK();
    grid_reduce<4>(gs, bs, v, sm);
    RMARK();
    const double f2 = v[0];
    const double rho = (f1 - f2) / (-(v[3] + 0.5 * v[2]));
    if (rho > 0.75) {
      if (status == 0 || status == 1) Delta = fmin(2.0 * Delta, maxDelta);
    } else if (rho < 0.25) {
      Delta = 0.25 * Delta;
    }
    const bool accept =
        (rho > 0.1) || (fabs(f1 - f2) / (fabs(f1) + 1.0) < 1.4901161193847656e-08 && f2 < f1);
    ++iter;
    out.outer++;
    if (accept) {
      x1 = cand;
      cand = (cand == A.X2) ? A.X3 : A.X2;
      f1 = f2;
      ngf = sqrt(v[1]);
      double *t;
      t = Rg1; Rg1 = Rg2; Rg2 = t;
      t = Rg1T; Rg1T = Rg2T; Rg2T = t;
      t = S1; S1 = S2; S2 = t;
    } else {
      out.rej++;
    }
    stop = ngf < P.gradnorm_tol;
    if (single) {
      // single-step mode: shrink the radius until the step is accepted
      if (accept) break;
      if (shrink > 10) break;  // give up: x1 is still the starting point
      Delta = maxDelta = maxDelta / 4.0;
      ++shrink;
    }
  }
#undef RMARK
  out.x = x1;
  out.f_opt = f1;
  out.gn_opt = ngf;
  return out;
}

// ---------------------------------------------------------------------------
// deferred reporting sums.  Values nobody needs on the critical path (fInit,
// gradNormInit, fOpt, gradNormOpt of mLocalOptResult, and the relative change
// between leader turns) are NOT grid-reduced when they are produced: every warp
// parks its partial in defer[agent][cta][warp][q] and the totals are formed
// (fixed order) when somebody can observe them -- at the leader's turn
// (shouldTerminate, src/PGOAgentROS.cpp:208) and at kernel exit.
//   q: 0 f_init, 1 |grad_init|^2, 2 f_opt, 3 |grad_opt|^2, 4 |X+ - X|^2
// ---------------------------------------------------------------------------
constexpr int kDeferQ = 8;
__device__ __forceinline__ double *defer_slot(const TeamDev &T, int ai) {
  return T.defer + (((size_t)ai * gridDim.x + blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5)) * kDeferQ;
}
__device__ __forceinline__ void defer_store(const TeamDev &T, int ai, int q, double partial) {
  partial = wsum32(partial);
  if ((threadIdx.x & 31) == 0) defer_slot(T, ai)[q] = partial;
}
// total of quantity q of agent ai over the whole grid (warp-collective, same order everywhere).  The loads of
// eight entries are issued together: the loop is a chain of L2 round trips otherwise (~37 per lane at 148 CTAs)
__device__ __forceinline__ double defer_total(const TeamDev &T, int ai, int q) {
  const int lane = threadIdx.x & 31;
  const int entries = (int)gridDim.x * (kThreads / 32);
  const double *base = T.defer + (size_t)ai * entries * kDeferQ + q;
  double s = 0;
  for (int e0 = lane; e0 < entries; e0 += 32 * 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + 32 * u;
      v[u] = (e < entries) ? __ldcg(base + (size_t)e * kDeferQ) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) s += v[u];
  }
  return wsum32(s);
}
// all five reporting sums of agent ai at once (warp-collective): an entry is one 64-byte row
__device__ __forceinline__ void defer_total5(const TeamDev &T, int ai, double (&t)[5]) {
  const int lane = threadIdx.x & 31;
  const int entries = (int)gridDim.x * (kThreads / 32);
  const double *base = T.defer + (size_t)ai * entries * kDeferQ;
#pragma unroll
  for (int q = 0; q < 5; ++q) t[q] = 0;
  for (int e0 = lane; e0 < entries; e0 += 32 * 4) {
    double2 a[4], b[4];
    double c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + 32 * u;
      const bool ok = e < entries;
      const double *row = base + (size_t)(ok ? e : 0) * kDeferQ;
      a[u] = ok ? __ldcg(reinterpret_cast<const double2 *>(row)) : make_double2(0, 0);
      b[u] = ok ? __ldcg(reinterpret_cast<const double2 *>(row) + 1) : make_double2(0, 0);
      c[u] = ok ? __ldcg(row + 4) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      t[0] += a[u].x; t[1] += a[u].y; t[2] += b[u].x; t[3] += b[u].y; t[4] += c[u];
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) t[q] = wsum32(t[q]);
}

// f and |rgrad|^2 of an agent whose gradient came from k_edge_grad (edge_grad.cu): warp 0 of CTA 0 sums that kernel's
// per-CTA partials (fixed order) into the same parked slots a gradient phase of this kernel would have filled
__device__ __forceinline__ void ext_grad_stats(const TeamDev &T, const RunArgs &args, int ai) {
  double pf = 0, pg2 = 0;
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const double *p = args.ext_partials[ai];
    for (int b = threadIdx.x; b < args.ext_grid[ai]; b += 32) {
      pf += __ldcg(p + 2 * b);
      pg2 += __ldcg(p + 2 * b + 1);
    }
  }
  defer_store(T, ai, 0, pf);
  defer_store(T, ai, 1, pg2);
}

// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
// M: local solver / schedule, fixed at compile time (0 RTR, 1 RGD, 2 RGD under the parallel schedule) so that the RGD kernel -- the bench workload and the
// stand-alone iterate() path -- does not carry the RTR-tCG code (instruction-cache footprint on a cold launch,
// register pressure)
// BIG: some agent's preconditioner slab does not fit shared memory (streaming dense pass)
template <int R, int M, bool BIG = false>
__global__ void __launch_bounds__(kThreads, 1)
    k_team_run(const __grid_constant__ TeamDev T, const __grid_constant__ RunArgs args) {
  extern __shared__ __align__(128) unsigned char dyn_smem_raw[];
  __shared__ double sm_red[64];
  __shared__ double sm_slab[8 * 16 * 8];
  __shared__ double sm_rel[kMaxLocal];
  __shared__ unsigned long long sm_mask;
  __shared__ ChunkTable chunks;
  __shared__ __align__(8) uint64_t mbar[1];
  SmemLayout L;
  L.slab = reinterpret_cast<double *>(dyn_smem_raw);
  L.slab_cap = args.slab_cap;
  L.stage = reinterpret_cast<double *>(dyn_smem_raw + args.slab_cap);
  L.zs = L.stage + kGroupsPerCta * kStageStride;
  const SolverParams &P = T.p;
  const GridSync &gs = T.gs;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) g_dbg = T.prof ? T.prof + 4096 : nullptr;
    mbar_init(&mbar[0]);
    chunks.prefix[0] = 0;
    for (int i = 0; i < T.num_local; ++i) {
      cta_pose_chunk(T.ag[i].n, chunks.p0[i], chunks.np[i]);
      chunks.prefix[i + 1] = chunks.prefix[i] + chunks.np[i];
    }
  }
  __syncthreads();
  SlabState ss{-1, 0u, 0};
  TeamCtl c = args.ctl_in;  // control state: identical in every thread
  BarState bs;
  bar_init(gs, bs);
  const int N = T.num_robots;
  const bool accel = P.acceleration != 0;
  const bool use_slab = (M == 0) || P.rgd_use_precond;
  if (M == 2 && !args.parallel) return;  // (never launched that way)
  const bool schedule = args.force_selected < -1;
  if (args.pull_mask && !args.armed) {
    // updateNeighborPoses staged the neighbours' poses in pinned host memory: fetch them with 16-byte loads
    // spread over the whole grid (one PCIe round trip, overlapped with the Nesterov phase that follows)
    for (int ai = 0; ai < T.num_local; ++ai)
      if (args.pull_mask & (1u << ai)) {
        const double2 *src = reinterpret_cast<const double2 *>(T.ag[ai].inbox_src);
        double2 *dst = reinterpret_cast<double2 *>(T.ag[ai].inbox_reg);
        const int cnt = T.ag[ai].inbox_doubles / 2;
        for (int i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads) dst[i] = src[i];
      }
    if (!accel || args.mode == 2) grid_barrier(gs, bs);  // otherwise the barrier after the Nesterov phase orders it
  }
  // multi-GPU: every rank runs this same loop on its own robots; see struct Fabric (device.cuh)
  const Fabric &F = T.fab;
  const bool fab = args.fabric && F.world > 1;
  FabState fs{F.seq0};
#define FAB_POST(value) fabric_post(F, F.nbr_ranks, (value))
  bool dead = false;
  int done = 0;
  int stop_reason = 0;
  int pend_ai = -1;        // RGD: agent whose post-step statistics (fOpt, gradNormOpt) are still due
  unsigned touched = 0;    // local agents with deferred sums newer than their AgentStat
  unsigned rel_due = 0;    // local agents whose relative change has not been totalled yet
#define PROF(k)                                                                                       \
  if (T.prof && step < T.prof_iters && threadIdx.x == 0 && (int)blockIdx.x == T.prof_cta)             \
    T.prof[step * 16 + (k)] = clock64();
  unsigned pend_all = 0;  // parallel schedule: agents whose fOpt / gradNormOpt are evaluated at exit
  for (int step = 0; step < args.max_iters; ++step) {
    const int iter = (args.mode == 2) ? c.iter : c.iter + 1;
    PROF(0)
    if constexpr (M == 2) {
      {
        // ---- asynchronous mode as its equal-rate / unit-delay schedule (oracle: Team::runParallel; wrapper:
        // runOnceAsynchronous, src/PGOAgentROS.cpp:119-127): every robot takes an RGD step against the neighbour
        // poses of the previous tick.  One agent per GPU makes the ticks of all robots truly concurrent.
        for (int ai = 0; ai < T.num_local; ++ai) {
          const AgentDev &A = T.ag[ai];
          if ((args.ext_grad_mask >> ai) & 1u) {   // G, Rg, RgT are there already (k_edge_grad in front of this launch)
            ext_grad_stats(T, args, ai);
            continue;
          }
          double pf = 0, pg2 = 0;
          phase_grad<R>(A, A.X, A.inbox_reg, true, nullptr, A.Rg, A.RgT, nullptr, L.stage, pf, pg2);
          defer_store(T, ai, 0, pf);
          defer_store(T, ai, 1, pg2);
        }
        grid_barrier(gs, bs);
        if (fab) {  // every rank has assembled its G: the inboxes may be overwritten
          fabric_arrive(F, fs, 0);
          if (!fabric_wait(F, gs, bs, fs.seq)) { dead = true; break; }
        }
        for (int ai = 0; ai < T.num_local; ++ai) {
          const AgentDev &A = T.ag[ai];
          if (use_slab) slab_prefetch(A, ai, ss, mbar, L.slab, L.slab_cap);
          double prel = 0;
          phase_rgd_step<R, BIG>(A, ai, P, A.X, false, false, 0.0, ss, mbar, L.slab, L.slab_cap, L.zs, sm_slab, A.X2,
                                 prel, ((args.ext_grad_mask >> ai) & 1u) ? args.ext_zt[ai] : nullptr);
          defer_store(T, ai, 4, prel);
          __syncthreads();  // the slab buffer and zs are reused by the next agent
        }
        if (fab) __threadfence_system();
        grid_barrier(gs, bs);
        if (fab) {  // every rank's X+ has reached its neighbours' inboxes
          fabric_arrive(F, fs, 0);
          if (!fabric_wait(F, gs, bs, fs.seq)) { dead = true; break; }
        }
        const unsigned everyone = (T.num_local >= 32) ? ~0u : ((1u << T.num_local) - 1u);
        touched |= everyone;
        rel_due |= everyone;
        pend_all = everyone;
        c.iter = iter;
        ++done;
        continue;
      }
    }
    int sel_robot, sel_local;
    if (!schedule) {
      sel_local = args.force_selected;
      sel_robot = sel_local >= 0 ? T.ag[sel_local].id : -1;
    } else {
      sel_robot = c.selected;
      sel_local = T.local_of_robot[sel_robot];
    }
    const bool restart = accel && ((iter + 1) % P.restart_interval == 0);
    // multi-GPU: this is global step `kstep` of the fabric's point-to-point protocol (struct Fabric, device.cuh)
    const unsigned long long kstep = F.step0 + (unsigned long long)step + 1ull;
    // Nesterov sequences come from the host (same arithmetic on every path): gamma_t, alpha_t
    double gamma = 0, alpha = 0;
    if (accel) {
      const double2 ga = args.gamma_tab ? args.gamma_tab[step] : make_double2(args.gamma0, args.alpha0);
      gamma = ga.x;
      alpha = ga.y;
      if (args.mode != 2) {
        __syncthreads();  // X / V / Y of my chunk were last written by other threads of this CTA
        // my neighbours have consumed what I stored into their inboxes in the previous step
        if (fab && kstep > 1 && !fabric_wait_prog(F, gs, bs, F.nbr_ranks, 2ull * (kstep - 1) + 1ull, /*acquire=*/false)) { dead = true; break; }
        PROF(7)
        LaCommit lc{nullptr, nullptr};
        if (step == 0 && args.la_commit > 0) {
          const size_t vec = (size_t)4 * R * T.ag[0].n;
          lc.X = T.ag[0].LX + (size_t)(args.la_commit - 1) * vec;
          lc.V = args.la_vsrc >= 0 ? T.ag[0].LX + (size_t)args.la_vsrc * vec : nullptr;
        }
        if (args.armed && step == 0 && use_slab && sel_local >= 0)   // its HBM / L2 latency hides behind the wait below
          slab_prefetch(T.ag[sel_local], sel_local, ss, mbar, L.slab, L.slab_cap);
        phase_nesterov_chunk<R>(T, chunks, sel_local, restart, alpha, lc);
        if (args.armed && step == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
          // everything that does not need the neighbours is done: wait for the host's iterate(true)
          volatile unsigned long long *db = reinterpret_cast<volatile unsigned long long *>(
              reinterpret_cast<unsigned char *>(T.ctl) + kCtlDoorbellOff);
          const unsigned long long go = args.seq * 2ull + 1ull, ab = args.seq * 2ull;
          const unsigned long long t0 = globaltimer_ns();
          int dec = 3;
          for (unsigned spins = 1;; ++spins) {
            const unsigned long long v = *db;
            if (v == go) { dec = 1; break; }
            if (v == ab) { dec = 2; break; }
            if ((spins & 31u) == 0 && globaltimer_ns() - t0 > args.arm_timeout_ns) {
              dec = (*db == go) ? 1 : 3;   // one last look: the host takes "expired" as final
              break;
            }
          }
          *reinterpret_cast<volatile int *>(args.arm_decision) = dec;
        }
        PROF(1)
        // peer stores of any CTA -> grid barrier (release / acquire at gpu scope) -> the poster's system fence + release
        // store: the progress word is ordered after them by cumulativity (the NCCL / NVSHMEM "barrier, then one thread
        // fences and flags" pattern).  A system fence in EVERY thread here costs ~6 us per phase (2-GPU profile,
        // profiles/fabric_profile_r2.txt); fab_variant bit 1 brings it back for comparison.
        if (fab && (args.fab_variant & 2)) __threadfence_system();
        grid_barrier(gs, bs);
        // "my Y (and X) of this step have landed".  A rank without the selected robot has no inbox to read in this
        // step and says so in the same word; the rank WITH it says nothing yet: nobody reads its Y before a later
        // step, and its "inbox consumed" word, a gradient phase later, covers these stores as well (one system
        // fence less on the critical path of the selected robot)
        if (fab && (sel_local < 0 || (args.fab_variant & 1))) FAB_POST(sel_local >= 0 ? 2ull * kstep : 2ull * kstep + 1ull);
        if (args.armed && step == 0) {
          const int dec = *reinterpret_cast<volatile int *>(args.arm_decision);   // written before CTA 0 arrived
          if (dec != 1) {
            // not solving after all (another call came first, or nobody rang in time): the agent's state is the
            // committed one with Y = X again, re-published, as if this launch had only materialised the lookahead
            const AgentDev &A0 = T.ag[0];
            for (int k0 = 0; k0 < chunks.np[0]; k0 += kGroupsPerCta) {
              const int k = k0 + (int)(threadIdx.x >> 3), a = threadIdx.x & 7;
              if (k < chunks.np[0]) {
                const int j = chunks.p0[0] + k;
                const bool act = a < R;
                double x[4];
                ld4(A0.X + (size_t)j * 4 * R, R, a, act, x);
                st4(A0.Y + (size_t)j * 4 * R, R, a, act, x);
                publish(A0.pub_rowptr, A0.pub_dst_aux, j, R, a, act, x);
              }
            }
            if (ss.pending) slab_wait(mbar, ss.parity);
            grid_barrier(gs, bs);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
              __threadfence_system();
              *reinterpret_cast<volatile unsigned long long *>(reinterpret_cast<unsigned char *>(T.ctl) + kCtlArmStateOff) =
                  args.seq * 2ull;
            }
            return;
          }
          // go: the neighbours' poses are staged in pinned host memory -- fetch them (16-byte loads, whole grid)
          const double2 *src = reinterpret_cast<const double2 *>(T.ag[0].inbox_src);
          double2 *dst = reinterpret_cast<double2 *>(T.ag[0].inbox_reg);
          const int cnt = T.ag[0].inbox_doubles / 2;
          for (int i = blockIdx.x * kThreads + threadIdx.x; i < cnt; i += gridDim.x * kThreads) dst[i] = __ldcv(src + i);
          grid_barrier(gs, bs);
        }
        PROF(2)
      }
    }
    if (fab && sel_local >= 0) {
      // the gate of src/PGOAgentROS.cpp:136-149, between GPUs: the selected robot waits for ITS neighbours only --
      // accelerated: their Y of this step; plain RBCD: the end of their previous step (their latest X+ has landed)
      const unsigned long long need = accel ? 2ull * kstep : 2ull * (kstep - 1) + 1ull;
      if (need > 1 && !fabric_wait_prog(F, gs, bs, F.agent_nbr_ranks[sel_local], need)) { dead = true; break; }
      PROF(8)
    }
    if (sel_local >= 0) {
      const AgentDev &A = T.ag[sel_local];
      const bool use_aux = accel && !restart;
      const double *Xs = use_aux ? A.Y : A.X;
      const double *inbox = use_aux ? A.inbox_aux : A.inbox_reg;
      if (use_slab) slab_prefetch(A, sel_local, ss, mbar, L.slab, L.slab_cap);  // no-op when already in flight
      if constexpr (M == 2) {
        // parallel schedule: handled at the top of the loop
      } else if constexpr (M == 1) {
        // ---- RGD (a2): gradient (+ the previous step's deferred statistics), preconditioned step
        // the two gradient passes are independent: warps 0-3 take the step's gradient, warps 4-7 the
        // deferred statistics of the previous step (both fit: <= 4 groups of 16 per CTA are busy)
        const bool split = pend_ai >= 0 && A.n <= 16 * (int)gridDim.x && T.ag[pend_ai].n <= 16 * (int)gridDim.x;
        if ((args.ext_grad_mask >> sel_local) & 1u) {   // computed by k_edge_grad in front of this launch
          ext_grad_stats(T, args, sel_local);
        } else {
          double pf = 0, pg2 = 0;
          phase_grad<R>(A, Xs, inbox, true, nullptr, A.Rg, A.RgT, nullptr, L.stage, pf, pg2, 0, split ? 4 : 8);
          defer_store(T, sel_local, 0, pf);
          defer_store(T, sel_local, 1, pg2);
        }
        if (pend_ai >= 0) {
          const AgentDev &B = T.ag[pend_ai];
          double qf = 0, qg2 = 0;
          phase_grad<R>(B, B.X2, nullptr, false, nullptr, nullptr, nullptr, nullptr, L.stage, qf, qg2,
                        split ? 4 : 0, split ? 4 : 8);
          defer_store(T, pend_ai, 2, qf);
          defer_store(T, pend_ai, 3, qg2);
          pend_ai = -1;
        }
        PROF(3)
        grid_barrier(gs, bs);
        // G is assembled: my neighbours may overwrite my inbox (their next Nesterov phase)
        if (fab && accel) FAB_POST(2ull * kstep + 1ull);
        PROF(4)
        double prel = 0;
        phase_rgd_step<R, BIG>(A, sel_local, P, Xs, accel, restart, gamma, ss, mbar, L.slab, L.slab_cap, L.zs,
                               sm_slab, A.X2, prel, ((args.ext_grad_mask >> sel_local) & 1u) ? args.ext_zt[sel_local] : nullptr);
        defer_store(T, sel_local, 4, prel);
        // the next agent's slab is fetched while the following phases run
        if (use_slab && schedule) {
          const int nxt = T.local_of_robot[(sel_robot + 1) % N];
          if (nxt >= 0) slab_prefetch(T.ag[nxt], nxt, ss, mbar, L.slab, L.slab_cap);
        }
        PROF(5)
        pend_ai = sel_local;
        touched |= 1u << sel_local;
        rel_due |= 1u << sel_local;
        if (fab && (args.fab_variant & 2)) __threadfence_system();  // (see the Nesterov phase: the next poster's fence covers X+)
        if (!accel) {
          grid_barrier(gs, bs);  // plain RBCD has no Nesterov phase (and its sync) before the next gradient
          if (fab) FAB_POST(2ull * kstep + 1ull);   // my step, X+ included, is over
        }
        PROF(6)
      } else {
        // ---- RTR (a2)
        const RtrOut ro = rtr_solve<R, BIG>(A, sel_local, P, gs, bs, Xs, inbox, ss, mbar, L, sm_slab, sm_red);
        double v[1] = {0};
        phase_commit<R>(A, ro.x, accel, restart, gamma, v[0]);
        if (fab) __threadfence_system();
        if (schedule) {
          const int nxt = T.local_of_robot[(sel_robot + 1) % N];
          if (nxt >= 0) slab_prefetch(T.ag[nxt], nxt, ss, mbar, L.slab, L.slab_cap);
        }
        grid_reduce<1>(gs, bs, v, sm_red);
        if (fab) FAB_POST(2ull * kstep + 1ull);   // inbox consumed, X+ published
        const double relchange = sqrt(v[0] / A.n);
        const bool ready = !(relchange > P.rel_change_tol) && A.conv_ok;
        if (ready)
          c.ready_mask |= (1ull << sel_robot);
        else
          c.ready_mask &= ~(1ull << sel_robot);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
          AgentStat *st = A.stat;
          st->f_init = ro.f_init; st->f_opt = ro.f_opt; st->gn_init = ro.gn_init; st->gn_opt = ro.gn_opt;
          st->tcg_iters = ro.tcg; st->rtr_outer = ro.outer; st->rtr_rej = ro.rej;
          st->relchange = relchange;
          st->ready = ready;
          st->optimized = 1;
        }
      }
    }
    if (fab && !accel && sel_local < 0) FAB_POST(2ull * kstep + 1ull);   // nothing to do in this step
    c.iter = iter;
    if (P.robust && args.mode != 2) c.robust_inner_iter++;
    ++done;
    if (schedule) {
      c.selected = (sel_robot + 1) % N;  // RoundRobin, src/PGOAgentROS.cpp:464-472
      if (sel_robot == args.leader) {    // leader decides, :207-217
        if (rel_due) {
          // total the parked relative changes of every agent that stepped since the last turn
          grid_barrier(gs, bs);
          {
            const int ai = threadIdx.x >> 5;  // one warp per local agent (kMaxLocal == warps per CTA)
            if (ai < T.num_local && (rel_due & (1u << ai))) {
              const double t = defer_total(T, ai, 4);
              if ((threadIdx.x & 31) == 0) sm_rel[ai] = t;
            }
          }
          __syncthreads();
          for (int ai = 0; ai < T.num_local; ++ai)
            if (rel_due & (1u << ai)) {
              const double relchange = sqrt(sm_rel[ai] / T.ag[ai].n);
              const int rid = T.ag[ai].id;
              if (!(relchange > P.rel_change_tol) && T.ag[ai].conv_ok)
                c.ready_mask |= (1ull << rid);
              else
                c.ready_mask &= ~(1ull << rid);
            }
          rel_due = 0;
        }
        if (fab) {
          // every rank is authoritative for the ready bits of its own robots: union over the ranks
          const unsigned long long mine = c.ready_mask & F.local_mask;
          fabric_arrive(F, fs, mine);
          if (!fabric_wait(F, gs, bs, fs.seq)) { dead = true; break; }
          if (threadIdx.x == 0) sm_mask = fabric_or_payload(F, fs.seq, mine);
          __syncthreads();
          c.ready_mask = sm_mask;
        }
        const unsigned long long all = (N >= 64) ? ~0ull : ((1ull << N) - 1ull);
        bool terminate;
        if (iter > P.max_num_iters)
          terminate = true;
        else if (P.robust && c.weight_update_count < P.robust_num_weight_updates)
          terminate = false;
        else
          terminate = (c.ready_mask & all) == all;
        if (terminate) {
          stop_reason = 1;
          if (args.stop_on_terminate) break;
        } else if (P.robust && c.weight_update_count < P.robust_num_weight_updates &&
                   (c.robust_inner_iter >= P.robust_inner_iters || (c.ready_mask & all) == all)) {
          stop_reason = 2;
          break;
        }
      }
    }
  }
  // ---- epilogue: everything that was deferred becomes observable now
  grid_barrier(gs, bs);
  if (args.skip_stats) {  // fOpt / gradNormOpt are evaluated on demand by the host (finish_opt_stats) or by a later launch
    pend_ai = -1;
    pend_all = 0;
  }
  if (pend_all) {
    for (int ai = 0; ai < T.num_local; ++ai) {
      const AgentDev &B = T.ag[ai];
      double qf = 0, qg2 = 0;
      phase_grad<R>(B, B.X2, nullptr, false, nullptr, nullptr, nullptr, nullptr, L.stage, qf, qg2);
      defer_store(T, ai, 2, qf);
      defer_store(T, ai, 3, qg2);
    }
    grid_barrier(gs, bs);
  }
  if (pend_ai >= 0) {
    // statistics of the last RGD step (mLocalOptResult.fOpt / gradNormOpt, src/PGOAgentROS.cpp:169-172)
    const AgentDev &B = T.ag[pend_ai];
    double qf = 0, qg2 = 0;
    phase_grad<R>(B, B.X2, nullptr, false, nullptr, nullptr, nullptr, nullptr, L.stage, qf, qg2);
    defer_store(T, pend_ai, 2, qf);
    defer_store(T, pend_ai, 3, qg2);
    grid_barrier(gs, bs);
  }
  if (touched && blockIdx.x == 0) {
    const int ai = threadIdx.x >> 5;  // one warp per local agent
    if (ai < T.num_local && (touched & (1u << ai))) {
        double t[5];
        defer_total5(T, ai, t);
        if ((threadIdx.x & 31) == 0) {
          AgentStat *st = T.ag[ai].stat;
          const double relchange = sqrt(t[4] / T.ag[ai].n);
          st->f_init = t[0]; st->gn_init = sqrt(t[1]); st->f_opt = t[2]; st->gn_opt = sqrt(t[3]);
          st->relchange = relchange;
          st->ready = !(relchange > P.rel_change_tol) && T.ag[ai].conv_ok;
          st->optimized = 1;
          st->tcg_iters = 0; st->rtr_outer = 0; st->rtr_rej = 0;
        }
      }
  }
  // the ready bits of agents that stepped after the last leader turn (uniform: every thread needs c)
  if (rel_due) {
    {
      const int ai = threadIdx.x >> 5;
      if (ai < T.num_local && (rel_due & (1u << ai))) {
        const double t = defer_total(T, ai, 4);
        if ((threadIdx.x & 31) == 0) sm_rel[ai] = t;
      }
    }
    __syncthreads();
    for (int ai = 0; ai < T.num_local; ++ai)
      if (rel_due & (1u << ai)) {
        const double relchange = sqrt(sm_rel[ai] / T.ag[ai].n);
        const int rid = T.ag[ai].id;
        if (!(relchange > P.rel_change_tol) && T.ag[ai].conv_ok)
          c.ready_mask |= (1ull << rid);
        else
          c.ready_mask &= ~(1ull << rid);
      }
  }
  if (ss.pending) slab_wait(mbar, ss.parity);  // do not exit with a bulk copy in flight
  if (fab) __threadfence_system();
  grid_barrier(gs, bs);  // every CTA's result-block writes (outboxes, stats) are ordered before the flag
  if (fab && !dead) {
    // leave together: on return every publication of every rank has landed in its destination inbox
    fabric_arrive(F, fs, 0);
    if (!fabric_wait(F, gs, bs, fs.seq)) dead = true;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c.fab_seq = fs.seq;
    c.fab_step = F.step0 + (M == 2 ? 0ull : (unsigned long long)done);
    if (dead) stop_reason = -1;
    c.stop_reason = stop_reason;
    c.iters_done = done;
    c.seq = 0;
    *T.ctl = c;
    __threadfence_system();
    reinterpret_cast<volatile TeamCtl *>(T.ctl)->seq = args.seq;
  }
  if (args.la_depth > 0 && !dead) {
    // the host already has this launch's result; speculate the agent's next iterate(false) steps behind it
    phase_lookahead<R>(T.ag[0], args.la_depth, args.la_tab);
    grid_barrier(gs, bs);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      __threadfence_system();
      reinterpret_cast<volatile TeamCtl *>(T.ctl)->la_seq = args.seq;
    }
  }
}

constexpr size_t kMaxDynSmem = 227 * 1024 - 11 * 1024;  // leave room for the static arrays

// slab capacity + total dynamic bytes for a team / agent
static void smem_plan(int max_n, int grid, bool want_slab, size_t &slab_cap, size_t &total) {
  const size_t chunk = (size_t)std::max(1, (max_n + grid - 1) / grid);
  const size_t fixed = (size_t)kGroupsPerCta * kStageStride * sizeof(double) + chunk * 32 * sizeof(double);
  slab_cap = 0;
  if (want_slab && fixed + 16 * 1024 < kMaxDynSmem) slab_cap = ((kMaxDynSmem - fixed) / 128) * 128;
  total = slab_cap + fixed;
}

template <int R, int M, bool BIG>
cudaError_t launch_run_t(const TeamDev &T, RunArgs args, int grid, cudaStream_t stream) {
  int max_n = 1;
  for (int i = 0; i < T.num_local; ++i) max_n = std::max(max_n, T.ag[i].n);
  const bool want_slab = (T.p.method == 0) || T.p.rgd_use_precond;
  size_t slab_cap, smem;
  smem_plan(max_n, grid, want_slab, slab_cap, smem);
  args.slab_cap = slab_cap;
  static std::atomic<size_t> configured[64];  // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > configured[dev].load()) {
    cudaError_t err = cudaFuncSetAttribute(k_team_run<R, M, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured[dev].store(smem);
  }
  void *params[] = {(void *)&T, (void *)&args};
  count_launch();
  return cudaLaunchCooperativeKernel((void *)k_team_run<R, M, BIG>, dim3(grid), dim3(kThreads), params, smem, stream);
}

}  // namespace dpgo
