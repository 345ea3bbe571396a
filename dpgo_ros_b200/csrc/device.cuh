// Device-side data layout and primitives of the B200 RBCD path (sm_100a).
//
// Layout in HBM (all FP64): every lifted variable (X, Y, V, gradients, tCG
// vectors) is r x 4n column-major -- pose i owns 4r contiguous doubles
// [Y_i(:,0) Y_i(:,1) Y_i(:,2) p_i].  One 8-lane group works on one pose with
// lane a holding row a of the r x 4 block (r <= 8), so a pose is one coalesced
// 4r*8-byte segment and every per-pose 3x3 Gram matrix is three xor-shuffles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dpgo {

constexpr int kMaxGrid = 160;   // grid_sync polls at most 5 x 32 CTAs
constexpr int kMaxLocal = 8;     // co-located agents per team / device
constexpr int kMaxRobots = 64;   // robots in the whole problem
constexpr int kThreads = 256;    // CTA size of every kernel here
constexpr int kGroupsPerCta = kThreads / 8;
constexpr int kRed = 4;          // doubles per grid reduction
constexpr int kLaMax = 8;        // speculated iterate(false) steps per launch (stand-alone path)
constexpr int kMaxRanks = 8;     // GPUs (one process each) that run one problem together

struct AgentStat {
  double relchange, f_init, f_opt, gn_init, gn_opt;
  int ready, optimized, tcg_iters, rtr_outer, rtr_rej, pad;
};

struct AgentDev {
  int id, n, r, n_in;
  // readyToTerminate gate under robust costs: the share of loop closures whose weight has settled at 0 or 1 is at
  // least robustOptMinConvergenceRatio (src/PGOAgentROSNode.cpp:214); recomputed on the host at every weight change
  int conv_ok;
  // state
  double *X, *Y, *V, *Xinit;
  // Q as block-CSR by OUTPUT pose j: out_j += X_{col} * val (4x4 col-major)
  const int *q_rowptr, *q_col;
  const double *q_val;
  // linear term by output pose: G_j += inbox[slot] * val
  const int *s_rowptr, *s_slot;
  const double *s_val;
  // ELL copies read by the hot phases: fixed-width slots per pose (-1 padded) + CSR overflow
  const int *qe_col;      // [n][8]
  const double *qe_val;   // [n][8][16]
  const int *qo_rowptr, *qo_col;
  const double *qo_val;
  const int *se_slot;     // [n][4]
  const double *se_val;   // [n][4][16]
  const int *so_rowptr, *so_slot;
  const double *so_val;
  // neighbour public poses (regular and auxiliary), r x 4 per slot
  double *inbox_reg, *inbox_aux;
  // lookahead of the stand-alone path (phase_lookahead): speculated states [kLaMax][r x 4n], their public poses
  // in mapped host memory [kLaMax][la_stride], and the base of the regular outbox the publication lists point into
  double *LX, *la_out;
  const double *outbox_base;
  int la_stride;
  const double *inbox_src;  // pinned host staging block [reg | aux] filled by updateNeighborPoses (device-readable)
  int inbox_doubles;
  // publication lists, CSR by my pose: destinations of X (reg) and Y (aux)
  const int *pub_rowptr;
  double *const *pub_dst_reg;
  double *const *pub_dst_aux;
  // dense preconditioner (Q + lambda I)^-1, 4n x 4n, symmetric
  const double *Pinv;
  // work vectors (r x 4n) and their row-major "T" copies ([r][4n]) for the dense phase
  double *G, *Rg, *RgT, *Z, *eta, *dlt0, *dlt1, *Hd, *HdT, *rv, *rvT, *rw, *rwT, *X2, *X3, *Rg2, *Rg2T, *zeta;
  double *S, *S2;  // per pose sym(Y^T egrad_Y): 6 doubles
  AgentStat *stat;
};

struct SolverParams {
  int method;  // 0 RTR, 1 RGD
  double rgd_stepsize;
  int rgd_use_precond;
  int rtr_iterations, rtr_tcg_iterations;
  double rtr_initial_radius, gradnorm_tol;
  int acceleration, restart_interval;
  int robust;  // cost type != L2
  int robust_num_weight_updates, robust_inner_iters;
  int max_num_iters;
  double rel_change_tol;
};

// control state carried across launches (uniform across the grid)
struct TeamCtl {
  int iter;       // global iteration number (== every agent's iteration_number)
  int selected;   // robot id that holds the UPDATE token
  double gamma, alpha;
  unsigned long long ready_mask;
  int weight_update_count, robust_inner_iter;
  int stop_reason;  // 0 ran out of max_iters, 1 terminate, 2 weight update requested
  int iters_done;
  unsigned long long seq;  // completion flag: written last, after a system-scope fence (host polls it)
  unsigned epoch;          // grid_sync epoch reached at kernel exit (carried into the next launch)
  unsigned pad;
  unsigned long long fab_seq;  // multi-GPU: sequence number of the last fabric barrier this rank arrived at
  unsigned long long la_seq;   // stand-alone path: the lookahead written by launch `la_seq` is complete
  unsigned long long fab_step; // multi-GPU: global steps this rank has run over the fabric (base of the progress words)
};
constexpr size_t kCtlBytes = 128;   // result block: [TeamCtl (padded) | AgentStat x agents (128 B each) | outboxes]
constexpr size_t kStatBytes = 128;
// two words behind TeamCtl in its slot of the (host-mapped) result block, for ARMED launches of the stand-alone path:
// the host's doorbell (go = 2 seq + 1, abort = 2 seq) and the kernel's answer when it left without solving (2 seq)
constexpr size_t kCtlDoorbellOff = 96, kCtlArmStateOff = 104;
static_assert(sizeof(TeamCtl) <= kCtlDoorbellOff, "TeamCtl must leave room for the doorbell words in its slot");

struct GridSync {
  unsigned long long *counter;  // monotonically increasing arrival counter
  double *slots;                // [2][gridDim.x][kRed] reduction partials
};

// ---------------------------------------------------------------------------
// Multi-GPU fabric: one persistent kernel per GPU (one process each), all running the SAME
// global schedule.  Public poses are published by plain stores into the neighbour's inbox,
// which for a remote neighbour is peer memory of another GPU (CUDA IPC mapping, NVLink), and
// the ranks keep in step with flag words in each other's window -- the device-side replacement
// of the PublicPoses topic and of the iteration gate (src/PGOAgentROS.cpp:662-690, 136-149).
// Every rank's window starts with   flags[kMaxRanks] | payload[2][kMaxRanks] | prog[kMaxRanks]   (u64 each):
// flags[s] = number of the last ALL-RANK barrier rank s arrived at (monotone), payload[q][s] = the word rank s
// attached to its arrival at a barrier of parity q -- used where every rank needs every rank (the leader's
// termination test, entering / leaving a launch, the parallel schedule);
// prog[s] = POINT-TO-POINT progress word of rank s, written only into the windows of the ranks that host a
// neighbour of one of s's robots.  The synchronous schedules run on these alone (the wrapper's gate waits for
// activeNeighborIDs() only, src/PGOAgentROS.cpp:136-149).  In global step k (1-based, monotone across launches):
//     2k      "my Nesterov-phase publications of step k have landed in your inboxes"
//     2k + 1  "my inbox of step k is consumed" -- posted at once by a rank that does not hold the selected robot;
//             plain RBCD: "my step k, X+ publication included, is over"
//   a rank waits for 2(k-1)+1 from its neighbour ranks before its first store of step k into their inboxes, and the
//   selected robot's rank for 2k (plain RBCD: 2(k-1)+1) from that robot's neighbour ranks before it assembles G.
//   Model-checked in tests/test_fabric_protocol_model.py (freshness, no write-after-read, no deadlock).
// ---------------------------------------------------------------------------
struct Fabric {
  int world, rank;  // world <= 1: single-GPU team, no fabric
  unsigned long long *flags;                     // my window
  unsigned long long *payload;                   // my window, [2][kMaxRanks]
  unsigned long long *peer_flags[kMaxRanks];     // peers' windows (peer-mapped); [rank] unused
  unsigned long long *peer_payload[kMaxRanks];
  unsigned long long *prog;                      // my window: progress words of the other ranks
  unsigned long long *peer_prog[kMaxRanks];
  unsigned nbr_ranks;                            // ranks (bit s) that host a neighbour of any local robot
  unsigned agent_nbr_ranks[kMaxLocal];           // same, per local agent
  unsigned long long step0;       // global steps completed before this launch
  unsigned long long seq0;        // barriers completed before this launch
  unsigned long long local_mask;  // robots that live on this rank
  unsigned long long timeout_ns;  // a peer that does not show up for this long aborts the launch
};
constexpr size_t kFabricHeaderBytes = 512;  // 4 * kMaxRanks words + padding (the inboxes stay 256-byte aligned)
static_assert(4 * kMaxRanks * sizeof(unsigned long long) <= kFabricHeaderBytes, "fabric header too small");

struct TeamDev {
  int num_local, num_robots;
  int local_of_robot[kMaxRobots];
  int pose_prefix[kMaxLocal + 1];
  AgentDev ag[kMaxLocal];
  SolverParams p;
  GridSync gs;
  TeamCtl *ctl;
  // optional phase timeline (debug): clock64() of one CTA's thread 0 at phase boundaries
  long long *prof;
  int prof_iters, prof_cta;
  unsigned long long *done_counter;  // last-block-done counter of the non-cooperative kernels
  double *defer;                     // [num_local][grid][warps][8] parked reporting partials
  Fabric fab;
};

// relaxation rank: a compile-time constant in the persistent kernel (RC > 0) so that every pose
// access becomes base + immediate offset; read from the agent only in the generic helper kernels
template <int RC>
__device__ __forceinline__ int rdim(const AgentDev &A) {
  return RC ? RC : A.r;
}

// ---------------------------------------------------------------------------
// 8-lane group primitives (all 32 lanes of the warp must call these)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double gsum8(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ double wsum32(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// load / store row `a` of an r x 4 column-major pose block
__device__ __forceinline__ void ld4(const double *P, int r, int a, bool act, double (&x)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) x[c] = act ? P[c * r + a] : 0.0;
}
__device__ __forceinline__ void st4(double *P, int r, int a, bool act, const double (&x)[4]) {
  if (act) {
#pragma unroll
    for (int c = 0; c < 4; ++c) P[c * r + a] = x[c];
  }
}

// ---------------------------------------------------------------------------
// 3x3 symmetric helpers.  Symmetric storage order: 00 01 02 11 12 22.
// ---------------------------------------------------------------------------
struct Sym3 {
  double a00, a01, a02, a11, a12, a22;
};

__device__ __forceinline__ Sym3 sym3_mul_sym_commuting(const Sym3 &A, const Sym3 &B) {
  // product of two commuting symmetric matrices (result symmetric)
  Sym3 C;
  C.a00 = A.a00 * B.a00 + A.a01 * B.a01 + A.a02 * B.a02;
  C.a01 = A.a00 * B.a01 + A.a01 * B.a11 + A.a02 * B.a12;
  C.a02 = A.a00 * B.a02 + A.a01 * B.a12 + A.a02 * B.a22;
  C.a11 = A.a01 * B.a01 + A.a11 * B.a11 + A.a12 * B.a12;
  C.a12 = A.a01 * B.a02 + A.a11 * B.a12 + A.a12 * B.a22;
  C.a22 = A.a02 * B.a02 + A.a12 * B.a12 + A.a22 * B.a22;
  return C;
}

// Jacobi eigen-decomposition based A^{-1/2} (robust path, any SPD A)
static __device__ __noinline__ Sym3 sym3_invsqrt_jacobi(Sym3 A) {
  double w00 = 1, w01 = 0, w02 = 0, w10 = 0, w11 = 1, w12 = 0, w20 = 0, w21 = 0, w22 = 1;
  for (int sweep = 0; sweep < 12; ++sweep) {
    bool rotated = false;
    // (0,1)
    if (fabs(A.a01) > 1e-17 * sqrt(fabs(A.a00 * A.a11))) {
      rotated = true;
      double th = (A.a11 - A.a00) / (2.0 * A.a01);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      double c = rsqrt(t * t + 1.0), s = t * c;
      double a00 = A.a00 - t * A.a01, a11 = A.a11 + t * A.a01;
      double a02 = c * A.a02 - s * A.a12, a12 = s * A.a02 + c * A.a12;
      A.a00 = a00; A.a11 = a11; A.a01 = 0; A.a02 = a02; A.a12 = a12;
      double t0, t1;
      t0 = c * w00 - s * w01; t1 = s * w00 + c * w01; w00 = t0; w01 = t1;
      t0 = c * w10 - s * w11; t1 = s * w10 + c * w11; w10 = t0; w11 = t1;
      t0 = c * w20 - s * w21; t1 = s * w20 + c * w21; w20 = t0; w21 = t1;
    }
    // (0,2)
    if (fabs(A.a02) > 1e-17 * sqrt(fabs(A.a00 * A.a22))) {
      rotated = true;
      double th = (A.a22 - A.a00) / (2.0 * A.a02);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      double c = rsqrt(t * t + 1.0), s = t * c;
      double a00 = A.a00 - t * A.a02, a22 = A.a22 + t * A.a02;
      double a01 = c * A.a01 - s * A.a12, a12 = s * A.a01 + c * A.a12;
      A.a00 = a00; A.a22 = a22; A.a02 = 0; A.a01 = a01; A.a12 = a12;
      double t0, t1;
      t0 = c * w00 - s * w02; t1 = s * w00 + c * w02; w00 = t0; w02 = t1;
      t0 = c * w10 - s * w12; t1 = s * w10 + c * w12; w10 = t0; w12 = t1;
      t0 = c * w20 - s * w22; t1 = s * w20 + c * w22; w20 = t0; w22 = t1;
    }
    // (1,2)
    if (fabs(A.a12) > 1e-17 * sqrt(fabs(A.a11 * A.a22))) {
      rotated = true;
      double th = (A.a22 - A.a11) / (2.0 * A.a12);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      double c = rsqrt(t * t + 1.0), s = t * c;
      double a11 = A.a11 - t * A.a12, a22 = A.a22 + t * A.a12;
      double a01 = c * A.a01 - s * A.a02, a02 = s * A.a01 + c * A.a02;
      A.a11 = a11; A.a22 = a22; A.a12 = 0; A.a01 = a01; A.a02 = a02;
      double t0, t1;
      t0 = c * w01 - s * w02; t1 = s * w01 + c * w02; w01 = t0; w02 = t1;
      t0 = c * w11 - s * w12; t1 = s * w11 + c * w12; w11 = t0; w12 = t1;
      t0 = c * w21 - s * w22; t1 = s * w21 + c * w22; w21 = t0; w22 = t1;
    }
    if (!rotated) break;
  }
  const double d0 = 1.0 / sqrt(A.a00), d1 = 1.0 / sqrt(A.a11), d2 = 1.0 / sqrt(A.a22);
  Sym3 B;
  B.a00 = w00 * w00 * d0 + w01 * w01 * d1 + w02 * w02 * d2;
  B.a01 = w00 * w10 * d0 + w01 * w11 * d1 + w02 * w12 * d2;
  B.a02 = w00 * w20 * d0 + w01 * w21 * d1 + w02 * w22 * d2;
  B.a11 = w10 * w10 * d0 + w11 * w11 * d1 + w12 * w12 * d2;
  B.a12 = w10 * w20 * d0 + w11 * w21 * d1 + w12 * w22 * d2;
  B.a22 = w20 * w20 * d0 + w21 * w21 * d1 + w22 * w22 * d2;
  return B;
}

// A^{-1/2} for SPD A.  Near the identity (the hot-path case: A = M^T M with M a
// convex combination of nearby Stiefel points) use the coupled Newton-Schulz
// iteration -- multiply-add only, quadratic convergence, all iterates are
// polynomials in A so they commute and stay symmetric.  Otherwise Jacobi.
__device__ __forceinline__ Sym3 sym3_invsqrt(const Sym3 &A) {
  const double e00 = 1.0 - A.a00, e11 = 1.0 - A.a11, e22 = 1.0 - A.a22;
  const double en = e00 * e00 + e11 * e11 + e22 * e22 + 2.0 * (A.a01 * A.a01 + A.a02 * A.a02 + A.a12 * A.a12);
  if (!(en < 0.04)) return sym3_invsqrt_jacobi(A);
  Sym3 Yk = A;
  Sym3 Zk = {1, 0, 0, 1, 0, 1};
  for (int it = 0; it < 8; ++it) {
    Sym3 ZY = sym3_mul_sym_commuting(Zk, Yk);
    Sym3 T = {0.5 * (3.0 - ZY.a00), -0.5 * ZY.a01, -0.5 * ZY.a02, 0.5 * (3.0 - ZY.a11), -0.5 * ZY.a12,
              0.5 * (3.0 - ZY.a22)};
    const double r00 = 1.0 - ZY.a00, r11 = 1.0 - ZY.a11, r22 = 1.0 - ZY.a22;
    const double rn = r00 * r00 + r11 * r11 + r22 * r22 + 2.0 * (ZY.a01 * ZY.a01 + ZY.a02 * ZY.a02 + ZY.a12 * ZY.a12);
    Yk = sym3_mul_sym_commuting(Yk, T);
    Zk = sym3_mul_sym_commuting(T, Zk);
    if (rn < 1e-20) break;  // residual before this step < 1e-10 => after it ~1e-20
  }
  return Zk;
}

// Stiefel projection U V^T = M (M^T M)^{-1/2} of the r x 3 block whose row `a`
// this lane holds in m[0..2].  Group-collective.
__device__ __forceinline__ void stiefel_project_row(double (&m)[4]) {
  Sym3 A;
  A.a00 = gsum8(m[0] * m[0]);
  A.a01 = gsum8(m[0] * m[1]);
  A.a02 = gsum8(m[0] * m[2]);
  A.a11 = gsum8(m[1] * m[1]);
  A.a12 = gsum8(m[1] * m[2]);
  A.a22 = gsum8(m[2] * m[2]);
  const Sym3 B = sym3_invsqrt(A);
  const double o0 = m[0] * B.a00 + m[1] * B.a01 + m[2] * B.a02;
  const double o1 = m[0] * B.a01 + m[1] * B.a11 + m[2] * B.a12;
  const double o2 = m[0] * B.a02 + m[1] * B.a12 + m[2] * B.a22;
  m[0] = o0; m[1] = o1; m[2] = o2;
}

// sym(Y^T Z) of the rotation blocks (row a of each in y[], z[]).  Group-collective.
__device__ __forceinline__ Sym3 sym_ytz(const double (&y)[4], const double (&z)[4]) {
  Sym3 S;
  S.a00 = gsum8(y[0] * z[0]);
  S.a11 = gsum8(y[1] * z[1]);
  S.a22 = gsum8(y[2] * z[2]);
  S.a01 = 0.5 * gsum8(y[0] * z[1] + y[1] * z[0]);
  S.a02 = 0.5 * gsum8(y[0] * z[2] + y[2] * z[0]);
  S.a12 = 0.5 * gsum8(y[1] * z[2] + y[2] * z[1]);
  return S;
}
// z_Y <- z_Y - y * S   (row-wise); translation column untouched
__device__ __forceinline__ void sub_y_sym(const double (&y)[4], const Sym3 &S, double (&z)[4]) {
  const double o0 = y[0] * S.a00 + y[1] * S.a01 + y[2] * S.a02;
  const double o1 = y[0] * S.a01 + y[1] * S.a11 + y[2] * S.a12;
  const double o2 = y[0] * S.a02 + y[1] * S.a12 + y[2] * S.a22;
  z[0] -= o0; z[1] -= o1; z[2] -= o2;
}
// tangent projection at y of z (rotation part); group-collective
__device__ __forceinline__ void tangent_project_row(const double (&y)[4], double (&z)[4]) {
  const Sym3 S = sym_ytz(y, z);
  sub_y_sym(y, S, z);
}

// QF retraction of the rotation block: Q factor of (y + xi) with diag(R) > 0,
// computed as Cholesky-QR (A^T A = R^T R, Q = A R^{-1}); A^T A = I + xi^T xi is
// well conditioned on the tangent space so this matches Householder QR to
// rounding.  Translation: p + xi_p.  In/out: x = y + xi on entry (row a).
__device__ __forceinline__ void qf_row(double (&x)[4]) {
  const double g00 = gsum8(x[0] * x[0]);
  const double g01 = gsum8(x[0] * x[1]);
  const double g02 = gsum8(x[0] * x[2]);
  const double g11 = gsum8(x[1] * x[1]);
  const double g12 = gsum8(x[1] * x[2]);
  const double g22 = gsum8(x[2] * x[2]);
  // upper-triangular R with R^T R = G
  const double i00 = rsqrt(g00);          // 1 / r00
  const double r01 = g01 * i00, r02 = g02 * i00;
  const double i11 = rsqrt(g11 - r01 * r01);
  const double r12 = (g12 - r01 * r02) * i11;
  const double i22 = rsqrt(g22 - r02 * r02 - r12 * r12);
  const double q0 = x[0] * i00;
  const double q1 = (x[1] - q0 * r01) * i11;
  const double q2 = (x[2] - q0 * r02 - q1 * r12) * i22;
  x[0] = q0; x[1] = q1; x[2] = q2;
}

// ---------------------------------------------------------------------------
// grid-wide barrier + deterministic reduction (persistent cooperative kernel)
//
// One monotonically increasing 64-bit arrival counter (never reset: at kernel
// start it is a multiple of gridDim.x plus the arrivals of faster CTAs, so the
// epoch base is (value / gridDim.x) * gridDim.x).  Thread 0 of each CTA arrives
// with a release reduction and polls ONE address with acquire loads.
// Measured on B200 (148 CTAs): 2.2 us with three __threadfence() + atomicAdd,
// 1.25 us with release/acquire (this version).  A flag-per-CTA variant with the
// payload packed into the flags (NCCL-LL style) was tried and measured 2.4 us:
// 148 CTAs polling 148 slots each congest the few L2 lines that hold them.
// ---------------------------------------------------------------------------
struct BarState {
  unsigned long long next;  // meaningful in thread 0 only
  int parity;               // reduction slot parity (uniform)
  int dead;                 // thread 0 only: the counter was poisoned (a fabric wait timed out somewhere)
};
// A CTA that gives up on a peer adds this to the arrival counter: every grid barrier of the launch
// then falls through, so no CTA can be left spinning on one that already left.
constexpr unsigned long long kBarPoison = 1ull << 48;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void bar_init(const GridSync &gs, BarState &bs) {
  bs.next = 0;
  bs.parity = 0;
  bs.dead = 0;
  if (threadIdx.x == 0) {
    const unsigned long long start = ld_acquire_u64(gs.counter);
    bs.next = (start / gridDim.x) * gridDim.x + gridDim.x;
  }
}

__device__ __forceinline__ void grid_barrier(const GridSync &gs, BarState &bs) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add_u64(gs.counter, 1ull);
    unsigned long long v;
    while ((v = ld_acquire_u64(gs.counter)) < bs.next) {
    }
    bs.next += gridDim.x;
    if (v >= kBarPoison) bs.dead = 1;
  }
  __syncthreads();
}

// Sum `vals` over every thread of the grid; every thread gets the same totals,
// summed in a fixed order (bitwise reproducible run to run).  Includes a grid
// barrier, so it also orders global memory between phases.  Slots are double
// buffered by parity: a fast CTA may start writing the next reduction's slots
// while a slow one is still reading this one's.
// sm: 64 doubles of shared memory (per-warp partials in [0, 8 K), totals in [48, 48 + K))
template <int K>
__device__ __forceinline__ void grid_reduce(const GridSync &gs, BarState &bs, double (&vals)[K], double *sm) {
  static_assert(K <= kRed && 8 * K <= 48, "grid_reduce: scratch layout");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) vals[k] = wsum32(vals[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sm[warp * K + k] = vals[k];
  }
  double *slots = gs.slots + (size_t)bs.parity * gridDim.x * kRed;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) s += sm[w * K + k];
      slots[(size_t)blockIdx.x * kRed + k] = s;
    }
    red_release_add_u64(gs.counter, 1ull);
    unsigned long long v;
    while ((v = ld_acquire_u64(gs.counter)) < bs.next) {
    }
    bs.next += gridDim.x;
    if (v >= kBarPoison) bs.dead = 1;
  }
  __syncthreads();
  // warp 0 sums the per-CTA partials (fixed order: lane-strided, then the butterfly) and hands the totals to the other
  // warps through shared memory.  Until round 2 every warp read the slots itself: 1184 warps x 5 x K loads on the same
  // 37 cache lines cost 1.0 us (K = 1) to 3.2 us (K = 4) on top of the barrier (tools/probe_rtr_phases.py).
  if (warp == 0) {
    double s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
#pragma unroll
      for (int k = 0; k < K; ++k) s[k] += __ldcg(&slots[(size_t)b * kRed + k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = wsum32(s[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) sm[48 + k] = s[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) vals[k] = sm[48 + k];
  bs.parity ^= 1;
}

// ---------------------------------------------------------------------------
// fabric barrier (multi-GPU).  arrive: call right after a grid barrier that follows this rank's
// last stores into peer memory -- thread s of CTA 0 then fences at system scope and writes this
// rank's flag in peer s's window (release/acquire chain: peer stores of any CTA -> grid barrier
// -> fence.sys -> flag; every thread that stored into peer memory also fences at system scope before that grid
// barrier, so the chain does not lean on cumulativity across SMs over NVLink).  wait: lanes 0..world-1 of every CTA poll their peer's flag in the LOCAL
// window.  Split arrive / wait lets a rank announce "I am done reading my inbox" early.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct FabState {
  unsigned long long seq;  // barriers this rank arrived at so far (uniform)
};

__device__ __forceinline__ void fabric_arrive(const Fabric &F, FabState &fs, unsigned long long payload) {
  ++fs.seq;
  if (blockIdx.x == 0 && (int)threadIdx.x < F.world && (int)threadIdx.x != F.rank) {
    __threadfence_system();
    st_relaxed_sys_u64(F.peer_payload[threadIdx.x] + (fs.seq & 1) * kMaxRanks + F.rank, payload);
    st_release_sys_u64(F.peer_flags[threadIdx.x] + F.rank, fs.seq);
  }
}

// true: every peer arrived at barrier `upto`; false: timed out / the launch is being abandoned
__device__ __forceinline__ bool fabric_wait(const Fabric &F, const GridSync &gs, BarState &bs, unsigned long long upto) {
  int bad = (threadIdx.x == 0) ? bs.dead : 0;
  if (!bad && (int)threadIdx.x < F.world && (int)threadIdx.x != F.rank) {
    const unsigned long long *f = F.flags + threadIdx.x;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys_u64(f) < upto) {
      if ((++spins & 0x3ff) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > F.timeout_ns || ld_acquire_u64(gs.counter) >= kBarPoison) {
          bad = 1;
          break;
        }
      }
    }
  }
  bad = __syncthreads_or(bad);
  if (bad && threadIdx.x == 0 && !bs.dead) {
    red_release_add_u64(gs.counter, kBarPoison);  // release every local CTA that is (or will be) in a grid barrier
    bs.dead = 1;
  }
  return !bad;
}
// point-to-point progress words.  post: right after a grid barrier that follows this rank's stores into peer memory
// (same release chain as fabric_arrive); wait: lanes of every CTA poll the words of the ranks in `mask`.
__device__ __forceinline__ void fabric_post(const Fabric &F, unsigned mask, unsigned long long value) {
  if (blockIdx.x == 0 && (int)threadIdx.x < F.world && ((mask >> threadIdx.x) & 1u)) {
    __threadfence_system();
    st_release_sys_u64(F.peer_prog[threadIdx.x] + F.rank, value);
  }
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// acquire = false: the wait only guards this rank's later STORES (write-after-read on a neighbour's inbox); the
// branch on the polled value orders them, and skipping the system-scope acquire saves the ~3.5 us it takes to drain
// this SM's outstanding peer stores first (2-GPU profile)
__device__ __forceinline__ bool fabric_wait_prog(const Fabric &F, const GridSync &gs, BarState &bs, unsigned mask,
                                                 unsigned long long value, bool acquire = true) {
  int bad = (threadIdx.x == 0) ? bs.dead : 0;
  if (!bad && (int)threadIdx.x < F.world && ((mask >> threadIdx.x) & 1u)) {
    const unsigned long long *f = F.prog + threadIdx.x;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    // spin on relaxed loads (no ordering work per probe), acquire once the word is there
    while (ld_relaxed_sys_u64(f) < value) {
      if ((++spins & 0x3ff) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > F.timeout_ns || ld_acquire_u64(gs.counter) >= kBarPoison) {
          bad = 1;
          break;
        }
      }
    }
    if (acquire) (void)ld_acquire_sys_u64(f);
  }
  bad = __syncthreads_or(bad);
  if (bad && threadIdx.x == 0 && !bs.dead) {
    red_release_add_u64(gs.counter, kBarPoison);
    bs.dead = 1;
  }
  return !bad;
}
// OR of the words the ranks attached to barrier `seq` (call after fabric_wait(seq) succeeded)
__device__ __forceinline__ unsigned long long fabric_or_payload(const Fabric &F, unsigned long long seq,
                                                                unsigned long long mine) {
  unsigned long long v = mine;
  for (int s = 0; s < F.world; ++s)
    if (s != F.rank) v |= ld_acquire_sys_u64(F.payload + (seq & 1) * kMaxRanks + s);
  return v;
}

}  // namespace dpgo
