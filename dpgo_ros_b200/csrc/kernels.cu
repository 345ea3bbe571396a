// sm_100a kernels of the RBCD hot path.
//
//  * k_team_run<R>  -- the persistent cooperative kernel: runs whole global RBCD
//    iterations (Nesterov bookkeeping, G assembly, gradient, preconditioned RGD
//    step or RTR-tCG solve, retraction, public-pose publication, termination
//    test) for all co-located agents with grid barriers between phases and no
//    host round trip.  This is PGOAgent::iterate() for every robot plus the
//    wrapper's synchronous schedule (src/PGOAgentROS.cpp:160,1185,464-472,207-217).
//  * single-op kernels backing the parity hooks (eval / hess / precond /
//    manifold ops) and the GNC-TLS residual+weight kernel (a8).
#include <algorithm>
#include <atomic>

#include "kernels.h"
#include "phases.cuh"

namespace dpgo {

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
// eta (+)= alpha * dlt ;  r = rsrc + alpha * Hd (also row-major copy) ; partial |r|^2
__device__ __forceinline__ void phase_tcg_update(const AgentDev &A, double alpha, bool eta_zero, const double *dlt,
                                                 const double *Hd, const double *rsrc, double *eta, double *rv,
                                                 double *rvT, double &prr) {
  PoseIter it;
  const int n = A.n, r = A.r;
  const size_t n4 = (size_t)4 * n;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    if (!act) continue;
    const size_t off = (size_t)j * 4 * r;
    double d[4], h[4], rs[4], e[4];
    ld4(dlt + off, r, it.a, act, d);
    ld4(Hd + off, r, it.a, act, h);
    ld4(rsrc + off, r, it.a, act, rs);
    if (eta_zero) {
      e[0] = e[1] = e[2] = e[3] = 0.0;
    } else {
      ld4(eta + off, r, it.a, act, e);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      e[c] += alpha * d[c];
      rs[c] += alpha * h[c];
      prr += rs[c] * rs[c];
      rvT[(size_t)it.a * n4 + 4 * j + c] = rs[c];
    }
    st4(eta + off, r, it.a, act, e);
    st4(rv + off, r, it.a, act, rs);
  }
}

// eta (+)= tau * dlt  (trust-region boundary / negative curvature exit)
__device__ __forceinline__ void phase_axpy_eta(const AgentDev &A, double tau, bool eta_zero, const double *dlt,
                                               double *eta) {
  PoseIter it;
  const int n = A.n, r = A.r;
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

// dlt = -z + beta * dlt
__device__ __forceinline__ void phase_direction(const AgentDev &A, double beta, const double *Z, double *dlt) {
  PoseIter it;
  const int n = A.n, r = A.r;
  int j;
  while (it.next(n, j)) {
    const bool act = j < n && it.a < r;
    if (!act) continue;
    const size_t off = (size_t)j * 4 * r;
    double z[4], d[4];
    ld4(Z + off, r, it.a, act, z);
    ld4(dlt + off, r, it.a, act, d);
#pragma unroll
    for (int c = 0; c < 4; ++c) d[c] = -z[c] + beta * d[c];
    st4(dlt + off, r, it.a, act, d);
  }
}

// out = Retr_x(eta)  (group-collective)
__device__ __forceinline__ void phase_retract(const AgentDev &A, const double *X1, const double *eta, double *out) {
  PoseIter it;
  const int n = A.n, r = A.r;
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

__device__ __forceinline__ void phase_dot(const AgentDev &A, const double *U, const double *W, double &p) {
  PoseIter it;
  const int n = A.n, r = A.r;
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

// commit x1 as the agent's new X (RTR epilogue)
__device__ __forceinline__ void phase_commit(const AgentDev &A, const double *X1, bool accel, bool restart,
                                             double gamma, double &prel) {
  PoseIter it;
  const int n = A.n, r = A.r;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    double xn[4];
    ld4(X1 + (size_t)(valid ? j : 0) * 4 * r, r, it.a, act, xn);
    finish_pose(A, valid ? j : 0, valid, it.a, xn, accel, restart, gamma, nullptr, prel);
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

template <int R>
__device__ __forceinline__ RtrOut rtr_solve(const AgentDev &A, int ai, const SolverParams &P, const GridSync &gs,
                                            BarState &bs, int &parity, const double *Xs, const double *inbox,
                                            SlabState &ss, uint64_t *mbar, const SmemLayout &L, double *red,
                                            double *sm) {
  RtrOut out;
  out.outer = out.tcg = out.rej = 0;
  const double *x1 = Xs;
  double *cand = A.X2;
  double *Rg1 = A.Rg, *Rg1T = A.RgT, *S1 = A.S;
  double *Rg2 = A.Rg2, *Rg2T = A.Rg2T, *S2 = A.S2;
  double v[4];
  // gradient at the starting point (also assembles G)
  v[0] = v[1] = v[2] = v[3] = 0;
  phase_grad(A, x1, inbox, true, S1, Rg1, Rg1T, nullptr, L.stage, v[0], v[1]);
  grid_reduce<2>(gs, bs, parity, reinterpret_cast<double(&)[2]>(v), sm);
  double f1 = v[0], ngf = sqrt(v[1]);
  out.f_init = f1;
  out.gn_init = ngf;
  const bool single = (P.rtr_iterations == 1);
  double Delta = P.rtr_initial_radius;
  double maxDelta = single ? Delta : 5.0 * P.rtr_initial_radius;
  int iter = 0, shrink = 0;
  bool stop = false;
  const double theta = 1.0, kappa = 0.1;
  while (true) {
    if (!single && (stop || iter >= P.rtr_iterations)) break;
    // ---------------- tCG
    const double *rsrc = Rg1, *rsrcT = Rg1T;
    const double norm_r0 = ngf;
    v[0] = 0;
    phase_precond<R>(A, ai, x1, rsrc, rsrcT, A.Z, A.dlt0, ss, mbar, L.slab, L.slab_cap, L.zs, red, v[0]);
    grid_reduce<1>(gs, bs, parity, reinterpret_cast<double(&)[1]>(v), sm);
    double z_r = v[0], d_Pd = z_r, e_Pe = 0.0, e_Pd = 0.0;
    bool eta_zero = true;
    int status = 4;  // 0 negcurv, 1 exceeded, 2 lcon, 3 scon, 4 maxiter
    int j = 0;
    for (j = 0; j < P.rtr_tcg_iterations; ++j) {
      v[0] = 0;
      phase_hess(A, x1, S1, A.dlt0, A.Hd, L.stage, v[0]);
      grid_reduce<1>(gs, bs, parity, reinterpret_cast<double(&)[1]>(v), sm);
      const double d_Hd = v[0];
      const double alpha = z_r / d_Hd;
      const double e_Pe_new = e_Pe + 2.0 * alpha * e_Pd + alpha * alpha * d_Pd;
      if (d_Hd <= 0 || e_Pe_new >= Delta * Delta) {
        const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta * Delta - e_Pe))) / d_Pd;
        phase_axpy_eta(A, tau, eta_zero, A.dlt0, A.eta);
        eta_zero = false;
        status = (d_Hd <= 0) ? 0 : 1;
        break;
      }
      e_Pe = e_Pe_new;
      v[0] = 0;
      phase_tcg_update(A, alpha, eta_zero, A.dlt0, A.Hd, rsrc, A.eta, A.rv, A.rvT, v[0]);
      eta_zero = false;
      grid_reduce<1>(gs, bs, parity, reinterpret_cast<double(&)[1]>(v), sm);
      rsrc = A.rv;
      rsrcT = A.rvT;
      const double norm_r = sqrt(v[0]);
      const double tempnum = pow(norm_r0, theta);
      if (norm_r <= norm_r0 * fmin(tempnum, kappa)) {
        status = (kappa < tempnum) ? 2 : 3;
        break;
      }
      v[0] = 0;
      phase_precond<R>(A, ai, x1, rsrc, rsrcT, A.Z, nullptr, ss, mbar, L.slab, L.slab_cap, L.zs, red, v[0]);
      grid_reduce<1>(gs, bs, parity, reinterpret_cast<double(&)[1]>(v), sm);
      const double zold_rold = z_r;
      z_r = v[0];
      const double beta = z_r / zold_rold;
      phase_direction(A, beta, A.Z, A.dlt0);
      grid_barrier(gs, bs);
      e_Pd = beta * (e_Pd + alpha * d_Pd);
      d_Pd = z_r + beta * beta * d_Pd;
    }
    out.tcg += min(j + 1, P.rtr_tcg_iterations);
    if (eta_zero) {  // maxInner == 0: eta = 0
      phase_axpy_eta(A, 0.0, true, A.dlt0, A.eta);
    }
    grid_barrier(gs, bs);
    // ---------------- candidate, model decrease, ratio
    phase_retract(A, x1, A.eta, cand);
    grid_barrier(gs, bs);
    v[0] = v[1] = v[2] = v[3] = 0;
    phase_grad(A, cand, inbox, false, S2, Rg2, Rg2T, nullptr, L.stage, v[0], v[1]);
    phase_hess(A, x1, S1, A.eta, A.zeta, L.stage, v[2]);
    phase_dot(A, A.eta, Rg1, v[3]);
    grid_reduce<4>(gs, bs, parity, v, sm);
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
  out.x = x1;
  out.f_opt = f1;
  out.gn_opt = ngf;
  return out;
}

// ---------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 1)
    k_team_run(const __grid_constant__ TeamDev T, const __grid_constant__ RunArgs args) {
  extern __shared__ __align__(128) unsigned char dyn_smem_raw[];
  __shared__ double sm_red[64];
  __shared__ double sm_slab[8 * 16 * 8];
  __shared__ __align__(8) uint64_t mbar;
  SmemLayout L;
  L.slab = reinterpret_cast<double *>(dyn_smem_raw);
  L.slab_cap = args.slab_cap;
  L.stage = reinterpret_cast<double *>(dyn_smem_raw + args.slab_cap);
  L.zs = L.stage + kGroupsPerCta * kStageStride;
  const SolverParams &P = T.p;
  const GridSync &gs = T.gs;
  if (threadIdx.x == 0) mbar_init(&mbar);
  __syncthreads();
  SlabState ss{-1, 0u, 0};
  BarState bs;
  bar_init(gs, bs);
  // control state: identical in every thread
  TeamCtl c = args.ctl_in;
  int parity = 0;
  const int N = T.num_robots;
  const bool accel = P.acceleration != 0;
  const bool use_slab = (P.method == 0) || P.rgd_use_precond;
  int done = 0;
  int stop_reason = 0;
  int pend_ai = -1;  // RGD: agent whose post-step statistics (fOpt, gradNormOpt) are still due
#define PROF(k)                                                                                       \
  if (T.prof && step < T.prof_iters && threadIdx.x == 0 && (int)blockIdx.x == T.prof_cta)             \
    T.prof[step * 16 + (k)] = clock64();
  for (int step = 0; step < args.max_iters; ++step) {
    const int iter = c.iter + 1;
    PROF(0)
    int sel_robot, sel_local;
    if (args.force_selected >= -1) {
      sel_local = args.force_selected;
      sel_robot = sel_local >= 0 ? T.ag[sel_local].id : -1;
    } else {
      sel_robot = c.selected;
      sel_local = T.local_of_robot[sel_robot];
    }
    const bool restart = accel && ((iter + 1) % P.restart_interval == 0);
    double gamma = c.gamma, alpha = c.alpha;
    if (accel) {
      gamma = (1.0 + sqrt(1.0 + 4.0 * (double)N * N * gamma * gamma)) / (2.0 * N);
      alpha = 1.0 / (gamma * N);
      phase_nesterov(T, sel_local, restart, alpha);
      PROF(1)
      grid_barrier(gs, bs);
      PROF(2)
    }
    if (sel_local >= 0) {
      const AgentDev &A = T.ag[sel_local];
      const bool use_aux = accel && !restart;
      const double *Xs = use_aux ? A.Y : A.X;
      const double *inbox = use_aux ? A.inbox_aux : A.inbox_reg;
      if (use_slab) slab_prefetch(A, sel_local, ss, &mbar, L.slab, L.slab_cap);  // no-op when already in flight
      double v[4] = {0, 0, 0, 0};
      double rel2;
      if (P.method == 1) {
        // ---- RGD (a2): gradient (+ the previous step's deferred statistics), preconditioned step
        phase_grad(A, Xs, inbox, true, nullptr, A.Rg, A.RgT, nullptr, L.stage, v[0], v[1]);
        if (pend_ai >= 0) {
          const AgentDev &B = T.ag[pend_ai];
          phase_grad(B, B.X2, nullptr, false, nullptr, nullptr, nullptr, nullptr, L.stage, v[2], v[3]);
        }
        PROF(3)
        grid_reduce<4>(gs, bs, parity, v, sm_red);
        PROF(4)
        if (blockIdx.x == 0 && threadIdx.x == 0) {
          A.stat->f_init = v[0];
          A.stat->gn_init = sqrt(v[1]);
          if (pend_ai >= 0) {
            T.ag[pend_ai].stat->f_opt = v[2];
            T.ag[pend_ai].stat->gn_opt = sqrt(v[3]);
          }
        }
        v[0] = 0;
        phase_rgd_step<R>(A, sel_local, P, Xs, accel, restart, gamma, ss, &mbar, L.slab, L.slab_cap, L.zs, sm_slab,
                          A.X2, v[0]);
        // the next agent's slab is fetched while the remaining phases run
        if (use_slab && args.force_selected < -1) {
          const int nxt = T.local_of_robot[(sel_robot + 1) % N];
          if (nxt >= 0) slab_prefetch(T.ag[nxt], nxt, ss, &mbar, L.slab, L.slab_cap);
        }
        PROF(5)
        grid_reduce<1>(gs, bs, parity, reinterpret_cast<double(&)[1]>(v), sm_red);
        PROF(6)
        rel2 = v[0];
        pend_ai = sel_local;
      } else {
        // ---- RTR (a2)
        const RtrOut ro = rtr_solve<R>(A, sel_local, P, gs, bs, parity, Xs, inbox, ss, &mbar, L, sm_slab, sm_red);
        v[0] = 0;
        phase_commit(A, ro.x, accel, restart, gamma, v[0]);
        if (args.force_selected < -1) {
          const int nxt = T.local_of_robot[(sel_robot + 1) % N];
          if (nxt >= 0) slab_prefetch(T.ag[nxt], nxt, ss, &mbar, L.slab, L.slab_cap);
        }
        grid_reduce<1>(gs, bs, parity, reinterpret_cast<double(&)[1]>(v), sm_red);
        rel2 = v[0];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
          AgentStat *st = A.stat;
          st->f_init = ro.f_init; st->f_opt = ro.f_opt; st->gn_init = ro.gn_init; st->gn_opt = ro.gn_opt;
          st->tcg_iters = ro.tcg; st->rtr_outer = ro.outer; st->rtr_rej = ro.rej;
        }
      }
      const double relchange = sqrt(rel2 / A.n);
      const bool ready = !(relchange > P.rel_change_tol);
      if (ready)
        c.ready_mask |= (1ull << sel_robot);
      else
        c.ready_mask &= ~(1ull << sel_robot);
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        AgentStat *st = A.stat;
        st->relchange = relchange;
        st->ready = ready;
        st->optimized = 1;
      }
    }
    if (restart) {
      gamma = 0;
      alpha = 0;
    }
    c.gamma = gamma;
    c.alpha = alpha;
    c.iter = iter;
    if (P.robust) c.robust_inner_iter++;
    ++done;
    if (args.force_selected < -1) {
      c.selected = (sel_robot + 1) % N;  // RoundRobin, src/PGOAgentROS.cpp:464-472
      if (sel_robot == args.leader) {    // leader decides, :207-217
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
  if (pend_ai >= 0) {
    // statistics of the last RGD step (mLocalOptResult.fOpt / gradNormOpt, src/PGOAgentROS.cpp:169-172)
    const AgentDev &B = T.ag[pend_ai];
    double v[2] = {0, 0};
    phase_grad(B, B.X2, nullptr, false, nullptr, nullptr, nullptr, nullptr, L.stage, v[0], v[1]);
    grid_reduce<2>(gs, bs, parity, v, sm_red);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      B.stat->f_opt = v[0];
      B.stat->gn_opt = sqrt(v[1]);
    }
  }
  if (ss.pending) slab_wait(&mbar, ss.parity);  // do not exit with a bulk copy in flight
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c.stop_reason = stop_reason;
    c.iters_done = done;
    *T.ctl = c;
  }
}

// iterate(false) of accelerated agents when nobody on this device optimises: a single pose-local
// phase, so no grid barrier and no cooperative launch (src/PGOAgentROS.cpp:1185)
__global__ void __launch_bounds__(kThreads) k_nesterov_only(const __grid_constant__ TeamDev T,
                                                            const __grid_constant__ RunArgs args) {
  TeamCtl c = args.ctl_in;
  const int N = T.num_robots;
  const int iter = c.iter + 1;
  const bool accel = T.p.acceleration != 0;
  const bool restart = accel && ((iter + 1) % T.p.restart_interval == 0);
  double gamma = c.gamma, alpha = c.alpha;
  if (accel) {
    gamma = (1.0 + sqrt(1.0 + 4.0 * (double)N * N * gamma * gamma)) / (2.0 * N);
    alpha = 1.0 / (gamma * N);
    phase_nesterov(T, -1, restart, alpha);
  }
  if (restart) gamma = alpha = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c.gamma = gamma;
    c.alpha = alpha;
    c.iter = iter;
    if (T.p.robust) c.robust_inner_iter++;
    c.stop_reason = 0;
    c.iters_done = 1;
    *T.ctl = c;
  }
}

// barrier / reduction micro-benchmark (diagnostics)
__global__ void __launch_bounds__(kThreads, 1) k_barrier_bench(GridSync gs, int iters, int mode, double *out) {
  __shared__ double sm[64];
  int parity = 0;
  BarState bs;
  bar_init(gs, bs);
  double v[2] = {1.0, 2.0};
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      grid_barrier(gs, bs);
    } else {
      v[0] = 1.0;
      v[1] = 2.0;
      grid_reduce<2>(gs, bs, parity, v, sm);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = v[0];
}

// ---------------------------------------------------------------------------
// single-op kernels (parity hooks; also used by the team cost evaluation)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_eval(const __grid_constant__ AgentDev A, const double *X,
                                                   const double *inbox, double *egrad, double *rgrad,
                                                   double *partials /* [grid][2] */) {
  __shared__ double stage[kGroupsPerCta * kStageStride];
  double pf = 0, pg2 = 0;
  phase_grad(A, X, inbox, true, A.S, rgrad, A.RgT, egrad, stage, pf, pg2);
  __shared__ double sm[2 * (kThreads / 32)];
  pf = wsum32(pf);
  pg2 = wsum32(pg2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sm[warp * 2] = pf;
    sm[warp * 2 + 1] = pg2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      a += sm[w * 2];
      b += sm[w * 2 + 1];
    }
    partials[blockIdx.x * 2] = a;
    partials[blockIdx.x * 2 + 1] = b;
  }
}

__global__ void __launch_bounds__(kThreads) k_hess(const __grid_constant__ AgentDev A, const double *X,
                                                   const double *V, double *out) {
  __shared__ double stage[kGroupsPerCta * kStageStride];
  double p = 0;
  phase_hess(A, X, A.S, V, out, stage, p);
}

// rows of V (r x 4n col-major) -> VT ([r][4n])
__global__ void k_transpose_rows(const double *V, double *VT, int r, int n4) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < r * n4) {
    const int q = idx / r, a = idx % r;
    VT[(size_t)a * n4 + q] = V[idx];
  }
}

template <int R>
__global__ void __launch_bounds__(kThreads) k_precond(const __grid_constant__ AgentDev A, const double *X,
                                                      const double *V, const double *VT, double *out,
                                                      size_t slab_cap) {
  extern __shared__ __align__(128) unsigned char dyn_smem_raw[];
  __shared__ double sm_slab[8 * 16 * 8];
  __shared__ __align__(8) uint64_t mbar;
  double *slab = reinterpret_cast<double *>(dyn_smem_raw);
  double *zs = reinterpret_cast<double *>(dyn_smem_raw + slab_cap);
  if (threadIdx.x == 0) mbar_init(&mbar);
  __syncthreads();
  SlabState ss{-1, 0u, 0};
  double p = 0;
  phase_precond<R>(A, 0, X, V, VT, out, nullptr, ss, &mbar, slab, slab_cap, zs, sm_slab, p);
}

__global__ void __launch_bounds__(kThreads) k_manifold_op(int op, int r, int n, const double *Ain, const double *Bin,
                                                          double *out) {
  // op 0: Stiefel projection of A; 1: tangent projection of B at A; 2: retraction of B at A
  PoseIter it;
  int j;
  while (it.next(n, j)) {
    const bool valid = j < n;
    const bool act = valid && it.a < r;
    const size_t off = (size_t)(valid ? j : 0) * 4 * r;
    double x[4], z[4] = {0, 0, 0, 0};
    ld4(Ain + off, r, it.a, act, x);
    if (op != 0) ld4(Bin + off, r, it.a, act, z);
    if (op == 0) {
      if (!valid) {
        x[0] = (it.a == 0); x[1] = (it.a == 1); x[2] = (it.a == 2);
      }
      stiefel_project_row(x);
      if (valid) st4(out + off, r, it.a, act, x);
    } else if (op == 1) {
      tangent_project_row(x, z);
      if (valid) st4(out + off, r, it.a, act, z);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) x[c] += z[c];
      if (!valid) {
        x[0] = (it.a == 0); x[1] = (it.a == 1); x[2] = (it.a == 2);
      }
      qf_row(x);
      if (valid) st4(out + off, r, it.a, act, x);
    }
  }
}

// publishPublicPoses for every local agent (src/PGOAgentROS.cpp:662-690): X -> reg, Y -> aux
__global__ void __launch_bounds__(kThreads) k_publish_all(const __grid_constant__ TeamDev T) {
  PoseIter it;
  const int total = T.pose_prefix[T.num_local];
  int item;
  while (it.next(total, item)) {
    if (item >= total) continue;
    int ai = 0;
    while (item >= T.pose_prefix[ai + 1]) ++ai;
    const AgentDev &A = T.ag[ai];
    const int j = item - T.pose_prefix[ai];
    const int r = A.r;
    const bool act = it.a < r;
    if (A.pub_rowptr[j] == A.pub_rowptr[j + 1]) continue;
    double x[4];
    ld4(A.X + (size_t)j * 4 * r, r, it.a, act, x);
    publish(A.pub_rowptr, A.pub_dst_reg, j, r, it.a, act, x);
    if (T.p.acceleration) {
      ld4(A.Y + (size_t)j * 4 * r, r, it.a, act, x);
      publish(A.pub_rowptr, A.pub_dst_aux, j, r, it.a, act, x);
    }
  }
}

// GNC-TLS (a8): residual and weight of every non-fixed loop closure.
// computeMeasurementResidual (src/PGOAgentROS.cpp:1049) + RobustCost::weight (:1050).
__global__ void k_gnc_weights(LcDev L, int r, const double *X, const double *inbox, double barc_sq, double mu,
                              int cost_type) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= L.count) return;
  const double *Xi = (L.src_remote[e] ? inbox : X) + (size_t)L.src[e] * 4 * r;
  const double *Xj = (L.dst_remote[e] ? inbox : X) + (size_t)L.dst[e] * 4 * r;
  const double *Rm = L.R + (size_t)e * 9;  // column-major
  const double *tm = L.t + (size_t)e * 3;
  double rot = 0, tr = 0;
  for (int a = 0; a < r; ++a) {
    const double y0 = Xi[a], y1 = Xi[r + a], y2 = Xi[2 * r + a];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double s = y0 * Rm[c * 3] + y1 * Rm[c * 3 + 1] + y2 * Rm[c * 3 + 2] - Xj[c * r + a];
      rot += s * s;
    }
    const double s = Xj[3 * r + a] - Xi[3 * r + a] - (y0 * tm[0] + y1 * tm[1] + y2 * tm[2]);
    tr += s * s;
  }
  const double rsq = L.kappa[e] * rot + L.tau[e] * tr;
  L.residual[e] = sqrt(rsq);
  double w = 1.0;
  if (cost_type == 5) {
    const double upper = (mu + 1.0) / mu * barc_sq;
    const double lower = mu / (mu + 1.0) * barc_sq;
    if (rsq >= upper)
      w = 0.0;
    else if (rsq <= lower)
      w = 1.0;
    else
      w = sqrt(barc_sq * mu * (mu + 1.0) / rsq) - mu;
  }
  if (L.update_mask[e]) L.weight[e] = w;
}

// ---------------------------------------------------------------------------
// host-side launch wrappers
// ---------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
long long kernel_launch_count() { return g_launches.load(); }

constexpr size_t kMaxDynSmem = 227 * 1024 - 10 * 1024;  // leave room for the static arrays

// slab capacity + total dynamic bytes for a team / agent
static void smem_plan(int max_n, int grid, bool want_slab, size_t &slab_cap, size_t &total) {
  const size_t chunk = (size_t)std::max(1, (max_n + grid - 1) / grid);
  const size_t fixed = (size_t)kGroupsPerCta * kStageStride * sizeof(double) + chunk * 32 * sizeof(double);
  slab_cap = 0;
  if (want_slab && fixed + 16 * 1024 < kMaxDynSmem) slab_cap = ((kMaxDynSmem - fixed) / 128) * 128;
  total = slab_cap + fixed;
}

template <int R>
static cudaError_t launch_run_t(const TeamDev &T, RunArgs args, int grid, cudaStream_t stream) {
  int max_n = 1;
  for (int i = 0; i < T.num_local; ++i) max_n = std::max(max_n, T.ag[i].n);
  const bool want_slab = (T.p.method == 0) || T.p.rgd_use_precond;
  size_t slab_cap, smem;
  smem_plan(max_n, grid, want_slab, slab_cap, smem);
  args.slab_cap = slab_cap;
  static std::atomic<size_t> configured{0};
  if (smem > configured.load()) {
    cudaError_t err = cudaFuncSetAttribute(k_team_run<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    configured.store(smem);
  }
  void *params[] = {(void *)&T, (void *)&args};
  ++g_launches;
  return cudaLaunchCooperativeKernel((void *)k_team_run<R>, dim3(grid), dim3(kThreads), params, smem, stream);
}

cudaError_t launch_team_run(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream) {
  const int r = T.ag[0].r;
  if (r == 5) return launch_run_t<5>(T, args, grid, stream);
  if (r == 6) return launch_run_t<6>(T, args, grid, stream);
  return launch_run_t<8>(T, args, grid, stream);
}

cudaError_t launch_nesterov_only(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream) {
  ++g_launches;
  k_nesterov_only<<<grid, kThreads, 0, stream>>>(T, args);
  return cudaGetLastError();
}

int max_coop_grid(int device) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return sms;
}

cudaError_t launch_barrier_bench(const GridSync &gs, int iters, int mode, double *out, int grid, cudaStream_t s) {
  void *params[] = {(void *)&gs, (void *)&iters, (void *)&mode, (void *)&out};
  return cudaLaunchCooperativeKernel((void *)k_barrier_bench, dim3(grid), dim3(kThreads), params, 0, s);
}

cudaError_t launch_eval(const AgentDev &A, const double *X, const double *inbox, double *egrad, double *rgrad,
                        double *partials, int grid, cudaStream_t s) {
  ++g_launches;
  k_eval<<<grid, kThreads, 0, s>>>(A, X, inbox, egrad, rgrad, partials);
  return cudaGetLastError();
}
cudaError_t launch_hess(const AgentDev &A, const double *X, const double *V, double *out, int grid, cudaStream_t s) {
  ++g_launches;
  k_hess<<<grid, kThreads, 0, s>>>(A, X, V, out);
  return cudaGetLastError();
}
cudaError_t launch_transpose_rows(const double *V, double *VT, int r, int n4, cudaStream_t s) {
  ++g_launches;
  const int total = r * n4;
  k_transpose_rows<<<(total + 255) / 256, 256, 0, s>>>(V, VT, r, n4);
  return cudaGetLastError();
}

template <int R>
static cudaError_t launch_precond_t(const AgentDev &A, const double *X, const double *V, const double *VT,
                                    double *out, int grid, cudaStream_t s) {
  const size_t chunk = (size_t)std::max(1, (A.n + grid - 1) / grid);
  const size_t zs_bytes = chunk * 32 * sizeof(double);
  size_t slab_cap = ((kMaxDynSmem - zs_bytes) / 128) * 128;
  const size_t smem = slab_cap + zs_bytes;
  cudaError_t err = cudaFuncSetAttribute(k_precond<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  k_precond<R><<<grid, kThreads, smem, s>>>(A, X, V, VT, out, slab_cap);
  return cudaGetLastError();
}
cudaError_t launch_precond(const AgentDev &A, const double *X, const double *V, const double *VT, double *out,
                           int grid, cudaStream_t s) {
  ++g_launches;
  if (A.r == 5) return launch_precond_t<5>(A, X, V, VT, out, grid, s);
  if (A.r == 6) return launch_precond_t<6>(A, X, V, VT, out, grid, s);
  return launch_precond_t<8>(A, X, V, VT, out, grid, s);
}
cudaError_t launch_manifold_op(int op, int r, int n, const double *A, const double *B, double *out, int grid,
                               cudaStream_t s) {
  ++g_launches;
  k_manifold_op<<<grid, kThreads, 0, s>>>(op, r, n, A, B, out);
  return cudaGetLastError();
}
cudaError_t launch_publish_all(const TeamDev &T, int grid, cudaStream_t s) {
  ++g_launches;
  k_publish_all<<<grid, kThreads, 0, s>>>(T);
  return cudaGetLastError();
}
cudaError_t launch_gnc_weights(const LcDev &L, int r, const double *X, const double *inbox, double barc_sq,
                               double mu, int cost_type, cudaStream_t s) {
  if (L.count == 0) return cudaSuccess;
  ++g_launches;
  k_gnc_weights<<<(L.count + 127) / 128, 128, 0, s>>>(L, r, X, inbox, barc_sq, mu, cost_type);
  return cudaGetLastError();
}

}  // namespace dpgo
