// Host-visible launch interface of the sm_100a kernels (kernels.cu, dense_inverse.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "device.cuh"

namespace dpgo {

struct RunArgs {
  int max_iters;
  int force_selected;  // -2: follow the RoundRobin schedule; -1: nobody optimises; >= 0: that local agent
  int stop_on_terminate;
  int leader;          // robot id whose turn triggers the termination / weight-update test
  size_t slab_cap;     // bytes of shared memory reserved for the preconditioner slab (set by the launcher)
  TeamCtl ctl_in;      // control state at entry (the kernel writes the exit state to TeamDev::ctl)
  unsigned long long seq;  // completion sequence number the kernel publishes in TeamCtl::seq
  // Nesterov sequences are produced on the host (one arithmetic for every path): per-iteration
  // (gamma_t, alpha_t) table for multi-iteration launches, or the single pair below
  const double2 *gamma_tab;
  double gamma0, alpha0;
  // single-iteration split used when the public poses cross GPUs between the two halves:
  //   0 full iterate, 1 Nesterov phase only (iteration counter advances), 2 local solve only
  int mode;
  int fabric;  // 1: this launch is one rank of a multi-GPU run (TeamDev::fab is live)
  int fab_variant;  // diagnostics (DPGO_B200_FAB_VARIANT): bit 0 the selected rank keeps its redundant "2k" post,
                    // bit 1 every thread fences at system scope after a publishing phase (round 1's belt and braces)
  unsigned pull_mask;  // local agents whose staged host inbox (AgentDev::inbox_src) is copied in by the kernel
  // stand-alone lookahead (phases.cuh, phase_lookahead)
  int la_commit;       // speculated steps the host consumed since the last launch (state = LX[la_commit - 1])
  int la_vsrc;         // consumed step whose X became V (restart iteration), or -1
  int la_depth;        // steps to speculate at the end of this launch
  int commit_only;     // k_nesterov_only: write the consumed state back to X / Y / V and do nothing else
  double2 la_tab[kLaMax];  // (alpha, restart != 0) of the speculated steps
  // LARGE agents (streaming kernels, no acceleration): the gradient of these local agents (bit mask) was computed by
  // k_edge_grad launched in front of this kernel (edge_grad.cu); only its per-CTA sums of f and |rgrad|^2 are read here
  unsigned ext_grad_mask;
  const double *ext_partials[kMaxLocal];
  int ext_grid[kMaxLocal];
  // ... and whose preconditioned gradient Z^T = (Rg Pinv)^T ([r][4n] row-major) was computed by k_sym_precond
  // (sym_precond.cu: one triangle of Pinv streamed) -- the step phase then starts at the tangent projection
  const double *ext_zt[kMaxLocal];
  // ARMED launch of a stand-alone accelerated agent (host.cu, Agent::arm): the kernel is launched BEFORE the host's
  // iterate(true) call, does everything that needs no neighbour pose (lookahead commit, Nesterov phase, slab prefetch),
  // then CTA 0 waits for the doorbell word in the mapped result block; go -> pull the staged inbox and solve,
  // abort / time-out -> put Y back, report and leave
  int armed;
  unsigned long long arm_timeout_ns;
  int *arm_decision;   // device word: CTA 0's reading of the doorbell, published to the grid by the next grid barrier
  int parallel;        // 1: asynchronous mode as the equal-rate / unit-delay schedule (every robot steps every tick)
  int skip_stats;  // 1: leave fOpt / gradNormOpt of the last step to Agent::finish_opt_stats (AgentStat::optimized = 2)
};

// The agent's measurements as a structure of arrays, order [odometry | private loop closures | shared loop closures]
// (assemble.cu).  src / dst: local pose index, or inbox slot when that end belongs to a neighbour (flags bit 0 / 1).
struct MeasDev {
  int count;
  const double *R, *t, *kappa, *tau;   // R column-major 3x3
  double *w;                           // measurement weight (GNC-TLS rewrites it on the device)
  const unsigned char *skip;           // 1: shared edge with a deactivated neighbour -- contributes nothing
  const int *src, *dst;
  const unsigned char *flags;
};
// destination slots of the weight-dependent blocks: Q (block-CSR by output pose, nq slots, contributions
// qc_item = measurement * 4 + role listed per slot) and the linear term (ns blocks, s_item = measurement * 2 +
// incoming).  *_dst >= 0: ELL position (pose * W + k); < 0: -(1 + overflow index).
struct AssembleDev {
  int nq, ns;
  const int *qc_ptr, *qc_item, *q_dst;
  const int *s_item, *s_dst;
  double *q_val, *qe_val, *qo_val;
  double *s_val, *se_val, *so_val;
};
// residuals of `count` measurements (meas == nullptr: all, in order); update_mask != nullptr: entries with 1 also
// get their GNC-TLS weight written into MeasDev::w
struct ResidualJob {
  int count;
  const int *meas;
  const unsigned char *update_mask;
  double *residual;
  double barc_sq, mu;
  int cost_type;
};

// edge_grad.cu: the gradient of a LARGE agent straight from 128-byte edge records (the HBM-bound regime)
struct EdgeGradArgs {
  int n, build_g;
  const double *rec;         // [M][16]: R(9, column-major) | t(3) | w kappa | w tau | pad(2)
  const int *inc_ptr;        // [n + 1] incidence list of every pose ...
  const int2 *inc_item;      // ... (edge * 2 + "this pose is the edge's destination", pose at the other end, or
                             //      -(inbox slot + 1) when that end belongs to a neighbour)
  const double *Xin, *inbox;
  double *G, *Rg, *RgT;      // linear term (written when build_g, read otherwise), Riemannian gradient (+ row-major copy)
  double *partials;          // [grid][2]: f, |rgrad|^2 per CTA
  unsigned long long *tmarks;  // optional: [0] min start, [1] max end (globaltimer ns) over the CTAs
};
int edge_grad_grid(int n);
cudaError_t launch_pack_edge_records(const MeasDev &M, double *rec, cudaStream_t s);
cudaError_t launch_edge_grad(const EdgeGradArgs &a, int r, cudaStream_t s);
// sym_precond.cu: Zt = (V P)^T for a symmetric dense P, reading only its lower triangle (LARGE agents)
struct SymPrecondArgs {
  const double *P;        // n4 x n4 (leading dimension ld), column-major, symmetric
  size_t ld;
  int n4;
  const double *VT;       // [r][n4] row-major operand
  double *Zt;             // [r][n4] row-major result
  double *partials;       // per-tile partial sums (sym_precond_partial_doubles)
  const int *first_tile;  // [npanels + 1]
  int npanels, ntiles;
  int *tile_counter;
};
cudaError_t launch_sym_precond(const SymPrecondArgs &a, int r, int grid, cudaStream_t s);
// tile bookkeeping of the symmetric pass: fills first_tile_of_panel[npanels + 1], returns the number of tiles
int sym_precond_tiles(int n4, int r, std::vector<int> &first_tile_of_panel);
size_t sym_precond_partial_doubles(int ntiles, int r);
// whether a team takes the streaming ("BIG") kernels on a grid of `grid` CTAs
bool team_needs_streaming(const TeamDev &T, int grid);

long long kernel_launch_count();
long long dense_inverse_launch_count();
int max_coop_grid(int device);
cudaError_t launch_team_run(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream);
// all-rank meeting point in front of a fabric launch (one warp; barrier number `seq` in the windows)
cudaError_t launch_fabric_rendezvous(const Fabric &F, unsigned long long seq, cudaStream_t stream);
cudaError_t launch_nesterov_only(const TeamDev &T, const RunArgs &args, int grid, cudaStream_t stream);
cudaError_t launch_eval(const AgentDev &A, const double *X, const double *inbox, double *egrad, double *rgrad,
                        double *partials, int grid, cudaStream_t s);
// f and |rgrad|^2 at X against the cached G: per-CTA partials [grid][2]
cudaError_t launch_post_stats(const AgentDev &A, const double *X, double *partials, int grid, cudaStream_t s);
cudaError_t launch_hess(const AgentDev &A, const double *X, const double *V, double *out, int grid, cudaStream_t s);
cudaError_t launch_transpose_rows(const double *V, double *VT, int r, int n4, cudaStream_t s);
cudaError_t launch_precond(const AgentDev &A, const double *X, const double *V, const double *VT, double *out,
                           int grid, cudaStream_t s);
cudaError_t launch_manifold_op(int op, int r, int n, const double *A, const double *B, double *out, int grid,
                               cudaStream_t s);
cudaError_t launch_barrier_bench(const GridSync &gs, int iters, int mode, unsigned epoch0, double *out, int grid,
                                 cudaStream_t s);
cudaError_t launch_publish_all(const TeamDev &T, int grid, cudaStream_t s);
// Z = B * P, P symmetric n4 x n4 (ld), B / Z column-major with rows <= 8 rows
cudaError_t launch_rows_times_sym(const double *B, const double *P, size_t ld, int rows, int n4, double *Z,
                                  cudaStream_t s);
cudaError_t launch_assemble_values(const MeasDev &M, const AssembleDev &A, cudaStream_t s);
cudaError_t launch_measurement_residuals(const MeasDev &M, const ResidualJob &J, int r, const double *X,
                                         const double *inbox, cudaStream_t s);

// dense_inverse.cu: P <- (blocks scattered) ; P <- P^-1 (SPD), N multiple of 32
cudaError_t launch_scatter_blocks(double *P, size_t ld, const int *rowptr, const int *col, const double *val,
                                  int n, double lambda, int npad, cudaStream_t s);
cudaError_t spd_inverse(double *A, double *work, int N, int *d_info, cudaStream_t s);

}  // namespace dpgo
