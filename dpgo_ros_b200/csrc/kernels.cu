// sm_100a kernels of the RBCD hot path.
//
//  * k_team_run<R>  -- the persistent cooperative kernel: runs whole global RBCD
//    iterations (Nesterov bookkeeping, G assembly, gradient, preconditioned RGD
//    step or RTR-tCG solve, retraction, public-pose publication, termination
//    test) for all co-located agents with grid syncs between phases and no
//    host round trip.  This is PGOAgent::iterate() for every robot plus the
//    wrapper's synchronous schedule (src/PGOAgentROS.cpp:160,1185,464-472,207-217).
//  * single-op kernels backing the parity hooks (eval / hess / precond /
//    manifold ops) and the GNC-TLS residual+weight kernel (a8).
#include <algorithm>
#include <atomic>

#include "kernels.h"
#include "phases.cuh"

namespace dpgo {

void count_launch();

// iterate(false) of accelerated agents when nobody on this device optimises: a single pose-local
// phase, so no grid barrier and no cooperative launch (src/PGOAgentROS.cpp:1185)
__global__ void __launch_bounds__(kThreads) k_nesterov_only(const __grid_constant__ TeamDev T,
                                                            const __grid_constant__ RunArgs args) {
  TeamCtl c = args.ctl_in;
  const int iter = args.commit_only ? c.iter : c.iter + 1;
  const bool accel = T.p.acceleration != 0;
  const bool restart = accel && ((iter + 1) % T.p.restart_interval == 0);
  LaCommit lc{nullptr, nullptr};
  if (args.la_commit > 0) {
    const size_t vec = (size_t)4 * T.ag[0].r * T.ag[0].n;
    lc.X = T.ag[0].LX + (size_t)(args.la_commit - 1) * vec;
    lc.V = args.la_vsrc >= 0 ? T.ag[0].LX + (size_t)args.la_vsrc * vec : nullptr;
  }
  if (accel) phase_nesterov<0>(T, args.force_selected, restart, args.alpha0, lc, args.commit_only != 0);
  // same pose -> thread-group mapping as phase_nesterov, so every pose's new X / V come from this very thread
  if (accel && args.la_depth > 0) phase_lookahead<0>(T.ag[0], args.la_depth, args.la_tab);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long prev = atomicAdd(T.done_counter, 1ull);
    if (prev == gridDim.x - 1) {  // last block: everyone's outbox writes are visible system-wide
      *T.done_counter = 0ull;
      c.iter = iter;
      if (T.p.robust && !args.commit_only) c.robust_inner_iter++;
      c.stop_reason = 0;
      c.iters_done = args.commit_only ? 0 : 1;
      c.la_seq = args.seq;
      c.seq = 0;
      *T.ctl = c;
      __threadfence_system();
      reinterpret_cast<volatile TeamCtl *>(T.ctl)->seq = args.seq;
    }
  }
}

// Ranks enter a fabric launch together: arrive at all-rank barrier `seq`, wait for everybody (or for the time-out;
// the persistent kernel behind it then times out on its own first wait and reports it).
__global__ void k_fabric_rendezvous(const __grid_constant__ Fabric F, unsigned long long seq) {
  const int s = threadIdx.x;
  if (s >= F.world || s == F.rank) return;
  __threadfence_system();
  st_release_sys_u64(F.peer_flags[s] + F.rank, seq);
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_sys_u64(F.flags + s) < seq)
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > F.timeout_ns) break;
}
cudaError_t launch_fabric_rendezvous(const Fabric &F, unsigned long long seq, cudaStream_t stream) {
  count_launch();
  k_fabric_rendezvous<<<1, 32, 0, stream>>>(F, seq);
  return cudaGetLastError();
}

// barrier / reduction micro-benchmark (diagnostics)
__global__ void __launch_bounds__(kThreads, 1) k_barrier_bench(GridSync gs, int iters, int mode, unsigned epoch0,
                                                               double *out) {
  __shared__ double sm[64];
  BarState bs;
  bar_init(gs, bs);
  double v[2] = {1.0, 2.0};
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      grid_barrier(gs, bs);
    } else {
      v[0] = 1.0;
      v[1] = 2.0;
      grid_reduce<2>(gs, bs, v, sm);
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
  phase_grad<0>(A, X, inbox, true, A.S, rgrad, A.RgT, egrad, stage, pf, pg2);
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

// deferred mLocalOptResult.fOpt / gradNormOpt of a stand-alone iterate(true): gradient pass at X+ with the G of
// the solve (build_g = false), nothing written but the partial sums
__global__ void __launch_bounds__(kThreads) k_post_stats(const __grid_constant__ AgentDev A, const double *X,
                                                         double *partials /* [grid][2] */) {
  __shared__ double stage[kGroupsPerCta * kStageStride];
  __shared__ double sm[2 * (kThreads / 32)];
  double pf = 0, pg2 = 0;
  phase_grad<0>(A, X, nullptr, false, nullptr, nullptr, nullptr, nullptr, stage, pf, pg2);
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
  phase_hess<0>(A, X, A.S, V, out, stage, p);
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
  __shared__ __align__(8) uint64_t mbar[1];
  double *slab = reinterpret_cast<double *>(dyn_smem_raw);
  double *zs = reinterpret_cast<double *>(dyn_smem_raw + slab_cap);
  if (threadIdx.x == 0) mbar_init(&mbar[0]);
  __syncthreads();
  SlabState ss{-1, 0u, 0};
  double p = 0;
  phase_precond<R>(A, 0, X, V, VT, out, nullptr, ss, mbar, slab, slab_cap, zs, sm_slab, p);
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

// Z (rows x n4) = B (rows x n4) * P for a symmetric n4 x n4 matrix P (column-major, leading dimension ld): thread c
// owns output column c and walks P's ROW c -- by symmetry its column c -- so neighbouring threads read neighbouring
// addresses.  B and Z are column-major with `rows` (<= 8) rows.  Used by the Chordal initialisation (two solves per
// round, not a hot path).
__global__ void k_rows_times_sym(const double *B, const double *P, size_t ld, int rows, int n4, double *Z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n4) return;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int q = 0; q < n4; ++q) {
    const double p = P[(size_t)q * ld + c];
    for (int a = 0; a < rows; ++a) acc[a] = fma(B[(size_t)q * rows + a], p, acc[a]);
  }
  for (int a = 0; a < rows; ++a) Z[(size_t)c * rows + a] = acc[a];
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

// ---------------------------------------------------------------------------
// host-side launch wrappers
// ---------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
long long kernel_launch_count() { return g_launches.load(); }
void count_launch() { ++g_launches; }

constexpr size_t kMaxDynSmem = 227 * 1024 - 11 * 1024;  // leave room for the static arrays

template <int R, int M, bool BIG>
cudaError_t launch_run_t(const TeamDev &T, RunArgs args, int grid, cudaStream_t stream);
// a team with an agent whose slab (the columns of Pinv of one CTA's poses) exceeds the shared-memory budget takes the
// streaming kernels.  RGD under the synchronous schedule prefetches the selected agent's first sub-chunk during the
// phases in front of the dense pass, so it only streams when not even one pose's columns fit; RTR applies the
// preconditioner back to back (tCG) and streams as soon as the chunk would have to be re-filled in pieces.
static bool needs_streaming(const TeamDev &T, int grid) {
  const bool rtr = T.p.method != 1;
  if (!rtr && !T.p.rgd_use_precond) return false;
  int max_n = 1;
  for (int i = 0; i < T.num_local; ++i) max_n = std::max(max_n, T.ag[i].n);
  // same budget as smem_plan (team_run.cuh): what is left for the slab after the staging tiles and zs
  const size_t chunk = (size_t)std::max(1, (max_n + grid - 1) / grid);
  const size_t fixed = (size_t)kGroupsPerCta * kStageStride * sizeof(double) + chunk * 32 * sizeof(double);
  const size_t slab_cap = fixed + 16 * 1024 < kMaxDynSmem ? ((kMaxDynSmem - fixed) / 128) * 128 : 0;
  for (int i = 0; i < T.num_local; ++i) {
    const size_t ldp = ((size_t)4 * T.ag[i].n + 31) / 32 * 32;
    const size_t poses = rtr ? (size_t)(T.ag[i].n + grid - 1) / grid : 1;
    if (poses * 4 * ldp * sizeof(double) > slab_cap) return true;
  }
  return false;
}
bool team_needs_streaming(const TeamDev &T, int grid) { return needs_streaming(T, grid); }
template <int R>
static cudaError_t launch_run_m(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream) {
  if (T.p.method != 1)
    return needs_streaming(T, grid) ? launch_run_t<R, 0, true>(T, args, grid, stream)
                                    : launch_run_t<R, 0, false>(T, args, grid, stream);
  if (needs_streaming(T, grid))
    return args.parallel ? launch_run_t<R, 2, true>(T, args, grid, stream) : launch_run_t<R, 1, true>(T, args, grid, stream);
  return args.parallel ? launch_run_t<R, 2, false>(T, args, grid, stream) : launch_run_t<R, 1, false>(T, args, grid, stream);
}

cudaError_t launch_team_run(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream) {
  if (grid > kMaxGrid) return cudaErrorInvalidValue;
  switch (T.ag[0].r) {
    case 3: return launch_run_m<3>(T, args, grid, stream);
    case 4: return launch_run_m<4>(T, args, grid, stream);
    case 5: return launch_run_m<5>(T, args, grid, stream);
    case 6: return launch_run_m<6>(T, args, grid, stream);
    case 7: return launch_run_m<7>(T, args, grid, stream);
    case 8: return launch_run_m<8>(T, args, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_nesterov_only(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream) {
  ++g_launches;
  k_nesterov_only<<<grid, kThreads, 0, stream>>>(T, args);
  return cudaGetLastError();
}

int max_coop_grid(int device) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return std::min(sms, kMaxGrid);
}

cudaError_t launch_barrier_bench(const GridSync &gs, int iters, int mode, unsigned epoch0, double *out, int grid,
                                 cudaStream_t s) {
  void *params[] = {(void *)&gs, (void *)&iters, (void *)&mode, (void *)&epoch0, (void *)&out};
  return cudaLaunchCooperativeKernel((void *)k_barrier_bench, dim3(grid), dim3(kThreads), params, 0, s);
}

cudaError_t launch_eval(const AgentDev &A, const double *X, const double *inbox, double *egrad, double *rgrad,
                        double *partials, int grid, cudaStream_t s) {
  ++g_launches;
  k_eval<<<grid, kThreads, 0, s>>>(A, X, inbox, egrad, rgrad, partials);
  return cudaGetLastError();
}
cudaError_t launch_post_stats(const AgentDev &A, const double *X, double *partials, int grid, cudaStream_t s) {
  ++g_launches;
  k_post_stats<<<grid, kThreads, 0, s>>>(A, X, partials);
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
  switch (A.r) {
    case 3: return launch_precond_t<3>(A, X, V, VT, out, grid, s);
    case 4: return launch_precond_t<4>(A, X, V, VT, out, grid, s);
    case 5: return launch_precond_t<5>(A, X, V, VT, out, grid, s);
    case 6: return launch_precond_t<6>(A, X, V, VT, out, grid, s);
    case 7: return launch_precond_t<7>(A, X, V, VT, out, grid, s);
    case 8: return launch_precond_t<8>(A, X, V, VT, out, grid, s);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_manifold_op(int op, int r, int n, const double *A, const double *B, double *out, int grid,
                               cudaStream_t s) {
  ++g_launches;
  k_manifold_op<<<grid, kThreads, 0, s>>>(op, r, n, A, B, out);
  return cudaGetLastError();
}
cudaError_t launch_rows_times_sym(const double *B, const double *P, size_t ld, int rows, int n4, double *Z,
                                  cudaStream_t s) {
  if (rows < 1 || rows > 8) return cudaErrorInvalidValue;
  ++g_launches;
  k_rows_times_sym<<<(n4 + 127) / 128, 128, 0, s>>>(B, P, ld, rows, n4, Z);
  return cudaGetLastError();
}
cudaError_t launch_publish_all(const TeamDev &T, int grid, cudaStream_t s) {
  ++g_launches;
  k_publish_all<<<grid, kThreads, 0, s>>>(T);
  return cudaGetLastError();
}
}  // namespace dpgo
