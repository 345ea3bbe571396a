// Host side of the B200 RBCD path: pose-graph bookkeeping, device data layout
// construction and launch sequencing.  Mirrors the DPGO::PGOAgent call surface
// that dpgo_ros drives; all arithmetic runs in the kernels.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dpgo_b200.h"
#include "kernels.h"

namespace dpgo {

struct Meas {
  int r1, p1, r2, p2;
  double R[9];  // column-major
  double t[3];
  double kappa, tau, weight;
  bool fixed;
};

struct Error {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string &msg);
// nested section of the per-entry-point library clock (capi.cu, dpgo_b200_debug_api_profile): reported under a name
// that starts with '.', which callers leave out of the library total (the enclosing entry point already counts it)
struct ProfSection {
  const char *name;
  double t0;
  explicit ProfSection(const char *n);
  ~ProfSection();
};
void cuda_check(cudaError_t e, const char *what);

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count, bool zero = true) {
    if (count != n) {
      release();
      if (count) cuda_check(cudaMalloc((void **)&p, count * sizeof(T)), "cudaMalloc");
      n = count;
    }
    if (zero && n) cuda_check(cudaMemset(p, 0, n * sizeof(T)), "cudaMemset");
  }
  void upload(const std::vector<T> &h) {
    alloc(h.size(), false);
    if (n) cuda_check(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice), "H2D");
  }
};

class Team;

class Agent {
 public:
  Agent(int id, const dpgo_b200_params &p, int device);
  ~Agent();

  // ---- pose graph
  void add_measurement(const Meas &m);
  Meas *find_measurement(int r1, int p1, int r2, int p2);
  int measurement_index(int r1, int p1, int r2, int p2) const;  // position in [odom | plc | slc], -1 if unknown
  std::vector<int> neighbors() const { return std::vector<int>(nbrs.begin(), nbrs.end()); }
  const std::vector<int> &my_public_frames(int nbr) const;  // sorted; cached until the pose graph changes

  // ---- lifecycle
  void set_lifting_matrix(const double *Y);
  void initialize(const double *T_rowmajor_or_null);
  void initialize_chordal();  // local_initialization_method "Chordal" (src/PGOAgentROSNode.cpp:106-112)
  void initialize_in_global_frame(const double *Tw_rowmajor);
  void reset();
  bool iterate(bool do_opt);

  // ---- exchange
  int get_shared_pose_dict(int nbr, bool aux, int *frames, double *poses, int cap);
  void update_neighbor_poses(int nbr, bool aux, const int *frames, const double *poses, int count);

  // ---- status / GNC
  dpgo_b200_status get_status() const;
  bool should_terminate() const;
  bool should_update_weights() const;
  void update_measurement_weights();  // standalone (team of one) path
  bool compute_residual(int r1, int p1, int r2, int p2, double *res);
  void refresh_residuals();  // one launch over every measurement, cached until X / the inbox / the graph change
  double robust_weight(double residual) const;

  // ---- device structures
  // allocate + build everything that is stale.  wait_precond = false: the dense inverse is only STARTED (on the agent's
  // own non-blocking stream) -- what iterate(false) needs, and what lets the inverses of several robots on one GPU
  // run side by side instead of one after the other (each alone fills a fraction of the SMs)
  // precond_mode 2: start + wait (default); 1: start only; 0: only reserve its buffers (Agent::warm)
  void ensure_device(int precond_mode = 2);
  void reserve_preconditioner(bool with_workspace);  // the allocations of the build (they serialise streams)
  bool precond_reserved() const;
  void warm();                    // allocate, assemble and wire everything except the dense inverse (at initialisation)
  void start_preconditioner();    // enqueue scatter + inverse on pstream, no host synchronisation
  void finish_preconditioner();   // wait for it, read the pivot flag
  void quiesce_preconditioner();  // a build in flight reads d_q_*: wait before those arrays change (result discarded)
  cudaStream_t pstream = nullptr;
  cudaEvent_t pevent = nullptr, pevent_in = nullptr;
  bool precon_inflight = false;
  void build_structure(); // graph topology -> CSR / slots / publication lists
  void build_values();    // Q blocks, G blocks, LC arrays (weights) -> device
  void build_preconditioner();
  bool need_preconditioner() const { return P.method == 0 || P.rgd_use_preconditioner; }
  AgentDev dev_view() const;
  bool all_inbox_valid(bool aux) const;
  bool weights_converged() const;

  // identity / params
  int id;
  dpgo_b200_params P;
  int device;
  int r;
  Team *team = nullptr;        // the team that launches for this agent
  std::unique_ptr<Team> own;   // implicit team of one
  int local_index = 0;

  // pose graph (host)
  int n = 0;
  std::vector<Meas> odom, plc, slc;
  // (r1, p1, r2, p2) -> (0 odom / 1 plc / 2 slc, position in that vector)
  std::map<std::pair<std::pair<int, int>, std::pair<int, int>>, std::pair<int, int>> have;
  std::set<int> nbrs;
  mutable std::map<int, std::vector<int>> pub_frames_cache;
  // per neighbour: the contiguous inbox slot range that holds its poses and their (sorted) frame ids
  struct NbrSlots {
    int first = 0;
    std::vector<int> frames;
  };
  std::map<int, NbrSlots> nbr_slots;

  // state machine
  int state = 0, instance = 0, iter = 0;
  bool have_lift = false;
  double ylift[8 * 3];
  std::vector<double> Tlocal;  // 12 n, column-major 3x4 per pose
  dpgo_b200_opt_result opt{};
  dpgo_b200_status status{};
  std::map<int, dpgo_b200_status> team_status;
  std::set<int> inactive_robots;   // setRobotActive(id, false): excluded from the leader's termination / re-weighting tests
  int weight_update_count = 0, robust_inner_iter = 0;
  double mu;
  bool publish_requested = false;
  bool outbox_stale = true;  // outbox staging not yet filled by a publish

  // neighbour slots
  std::map<std::pair<int, int>, int> slot_of;  // (robot, frame) -> inbox slot
  std::vector<std::pair<int, int>> slot_key;
  std::vector<char> inbox_valid_reg, inbox_valid_aux;
  // outbox staging (for neighbours that are not co-located): per neighbour offset into outbox arrays
  std::map<int, std::pair<int, int>> outbox_range;  // nbr -> (first index, count)
  int outbox_total = 0;

  // dirtiness
  bool structure_dirty = true, values_dirty = true, precon_dirty = true, wiring_dirty = true;
  bool lc_dirty = true;        // the set of re-weightable loop closures changed (fixedWeight flipped)
  bool weights_host_dirty = true;  // host weights newer than MeasDev::w (set by every host-side weight write)

  // device buffers
  DevBuf<double> dX, dY, dV, dXinit;
  DevBuf<int> d_q_rowptr, d_q_col, d_s_rowptr, d_s_slot, d_pub_rowptr;
  DevBuf<double> d_q_val, d_s_val;
  // ELL(8) / ELL(4) copies + CSR overflow read by the hot phases
  DevBuf<int> d_qe_col, d_qo_rowptr, d_qo_col, d_se_slot, d_so_rowptr, d_so_slot;
  DevBuf<double> d_qe_val, d_qo_val, d_se_val, d_so_val;
  // inbox / outbox: [reg | aux] contiguous on the device, mirrored in pinned host memory so the
  // per-robot exchange API (getSharedPoseDict / updateNeighborPoses) costs one async copy per iterate
  DevBuf<double> d_inbox;
  double *inbox_ext = nullptr;  // multi-GPU: the inbox lives in the team's peer-visible window instead
  double *inbox_base() const { return inbox_ext ? inbox_ext : d_inbox.p; }
  // ---- lookahead of the stand-alone accelerated path (phases.cuh, phase_lookahead)
  DevBuf<double> dLX;                 // [kLaMax][r x 4n] speculated states
  double *d_la_out = nullptr, *h_la_out = nullptr;  // [kLaMax][la_stride] public poses, in the team's mapped result block
  int la_stride = 0;
  int la_valid = 0, la_used = 0;      // speculated steps available / consumed by iterate(false) without a launch
  int la_vsrc = -1;                   // consumed restart step (V = X there)
  unsigned long long la_launch = 0;   // launch whose tail wrote the lookahead (TeamCtl::la_seq)
  double la_gamma[kLaMax];            // Nesterov gamma after each speculated step
  unsigned char la_restart[kLaMax];
  bool lookahead_usable() const;      // stand-alone, accelerated, initialised
  void materialize_lookahead();       // write the consumed speculated state back to X / Y / V (one tiny launch)
  void drop_lookahead() { la_valid = la_used = 0; la_vsrc = -1; }
  // ---- armed launch (Agent::arm): the solve kernel of the NEXT iterate(true) already sits on the GPU and waits for
  // the doorbell, so that call costs neither the launch latency nor the Nesterov phase
  bool armed = false;
  int arm_backoff = 0;          // iterate(true) calls to sit out after an armed launch expired or was aborted
  // called when neighbour poses arrive and the next call is expected to be iterate(true); from_nbr / from_aux: the
  // delivery that triggered it (several robots per device arm on their RoundRobin predecessor's auxiliary poses only)
  void maybe_arm(int from_nbr = -1, bool from_aux = false);
  void disarm();                // ring "abort", wait for the kernel to leave, take over the committed state
  bool stats_pending = false;  // fOpt / gradNormOpt of the last iterate(true) not evaluated yet
  void finish_opt_stats();
  DevBuf<double> d_stat_partials;
  double *d_inbox_reg() const { return inbox_base(); }
  double *d_inbox_aux() const { return inbox_base() + (size_t)slot_key.size() * 4 * r; }
  // outbox + AgentStat live in the owning team's result block (one D2H copy per launch)
  double *d_outbox = nullptr, *h_outbox = nullptr;
  AgentStat *d_stat = nullptr, *h_stat = nullptr;
  size_t outbox_doubles() const { return (size_t)2 * std::max(1, outbox_total) * 4 * r; }
  double *d_outbox_reg() const { return d_outbox; }
  double *d_outbox_aux() const { return d_outbox + (size_t)std::max(1, outbox_total) * 4 * r; }
  double *h_inbox = nullptr;  // pinned
  bool inbox_dirty = false;      // host staging newer than the device inbox
  bool outbox_mirror_valid = false;
  void free_pinned();
  DevBuf<double *> d_pub_dst_reg, d_pub_dst_aux;
  DevBuf<double> dPinv, dPwork;           // dense preconditioner + the factorisation workspace
  DevBuf<int> dPinfo;
  DevBuf<double> dG, dRg, dRgT, dZ, dEta, dDlt0, dDlt1, dHd, dHdT, dRv, dRvT, dRw, dRwT, dX2, dX3, dRg2, dRg2T, dZeta, dS, dS2;
  // the measurements on the device, [odom | plc | slc] (assemble.cu: MeasDev), and the slot lists of the
  // weight-dependent blocks (AssembleDev)
  DevBuf<double> d_m_R, d_m_t, d_m_kappa, d_m_tau, d_m_w, d_m_resid;
  DevBuf<int> d_m_src, d_m_dst, d_qc_ptr, d_qc_item, d_q_dst, d_s_item, d_s_dst;
  DevBuf<unsigned char> d_m_flags, d_m_skip;
  MeasDev meas_view() const;
  AssembleDev assemble_view() const;
  int num_meas() const { return (int)(odom.size() + plc.size() + slc.size()); }
  Meas &meas_at(int idx) {
    return idx < (int)odom.size() ? odom[idx]
                                  : (idx < (int)(odom.size() + plc.size()) ? plc[idx - odom.size()]
                                                                           : slc[idx - odom.size() - plc.size()]);
  }
  // loop closures subject to re-weighting (non-fixed): measurement index + "this agent owns the weight"
  DevBuf<int> d_lc_meas;
  DevBuf<unsigned char> d_lc_mask;
  DevBuf<double> d_lc_residual;
  std::vector<int> lc_meas;
  std::vector<unsigned char> lc_mask;
  void build_lc_list();
  void upload_weights();
  // LARGE agents: 128-byte edge records + per-pose incidence lists for k_edge_grad (edge_grad.cu)
  static constexpr int kEdgeGradMinPoses = 1024;
  DevBuf<double> d_er_rec, d_eg_partials;
  DevBuf<int> d_inc_ptr;
  DevBuf<int2> d_inc_item;
  DevBuf<unsigned long long> d_eg_marks;   // [0] min start, [1] max end of the last profiled k_edge_grad (globaltimer ns)
  bool has_edge_arrays() const { return d_inc_ptr.n != 0; }
  // ... and the symmetric streaming pass of the dense preconditioner (sym_precond.cu): tile table, per-tile partials,
  // the result Z^T
  DevBuf<double> d_sym_partials, dZt;
  DevBuf<int> d_sym_first_tile, d_sym_counter;
  int sym_ntiles = 0, sym_npanels = 0;
  void ensure_sym_buffers();
  bool eg_profile = false;
  // cached residuals of every measurement (TERMINATE handler, src/PGOAgentROS.cpp:1044-1057)
  std::vector<double> h_resid;
  bool resid_valid = false;
  // host copies of structure used for wiring
  std::vector<int> h_pub_rowptr;
  std::vector<std::pair<int, int>> h_pub_entries;  // (neighbour, my frame) per publication entry
  std::vector<int> h_q_rowptr, h_q_col, h_s_rowptr, h_s_slot;
};

class Team {
 public:
  explicit Team(int device);
  ~Team();
  void add(Agent *a);
  void remove(Agent *a);
  // ensure every agent's device data + wiring + TeamDev are current; need_inbox: the next launch reads the
  // neighbour poses, so staged host inboxes are uploaded (iterate(false) does not pay for that copy)
  void prepare(bool need_inbox = true, bool keep_lookahead = false);
  void exchange_all();
  dpgo_b200_run_result run(int max_iters, bool stop_on_terminate);
  int parallel_schedule_checked() const;
  int schedule = 0;  // 0: synchronous RoundRobin token (:464-472); 1: parallel ticks (asynchronous mode, :119-127)
  // one iteration with a forced selection (standalone iterate path); returns kernel ms
  void run_forced(int sel_local);
  void gnc_update_all();
  double global_cost();
  void sync_ctl_from_agents();
  void read_back();  // ctl + per-agent stats -> host

  int device;
  int grid = 0;
  cudaStream_t stream = nullptr;
  // result block: [TeamCtl | AgentStat per agent | outbox per agent], device + pinned mirror
  unsigned char *d_result = nullptr, *h_result = nullptr;
  size_t result_bytes = 0;
  unsigned long long seq = 0;
  // explicit (multi-agent / multi-GPU) teams keep the outboxes in DEVICE memory so that NCCL / peer
  // copies can read them; the implicit team of a stand-alone agent keeps them in the mapped host block
  bool device_outbox = false;
  DevBuf<double> dOutboxAll;
  double last_gamma_use = 0, last_alpha_use = 0;
  void step(int selected_robot, int mode);
  // ---- multi-GPU fabric (struct Fabric, device.cuh): this team is rank `fab_rank` of `fab_world`
  // processes, one per GPU, that run the global schedule together in their own persistent kernels
  struct Route {
    int peer;            // rank that owns the neighbour
    size_t off_reg, off_aux;  // byte offsets, in the peer's window, of my contiguous range in its inbox
  };
  int fab_world = 0, fab_rank = 0;
  unsigned char *window = nullptr;  // [flags | payload | inbox of every local agent], cudaMalloc'ed (IPC-exportable)
  size_t window_bytes = 0;
  unsigned char *peer_base[kMaxRanks] = {};
  bool peer_is_ipc[kMaxRanks] = {};
  unsigned long long fab_seq = 0;
  unsigned long long fab_step = 0;   // global steps run over the fabric (base of the point-to-point progress words)
  double fab_timeout_s = 20.0;
  std::map<std::pair<int, int>, Route> routes;  // (local robot, remote neighbour)
  void fabric_init(int world, int rank);
  void fabric_import(int peer, const cudaIpcMemHandle_t *handle, void *same_process_base);
  void fabric_route(int robot, int nbr, int peer, size_t off_reg, size_t off_aux);
  dpgo_b200_run_result fabric_run(int max_iters, bool stop_on_terminate);
  void fabric_close();
  // GNC weight update split in two so that the caller can carry weights across ranks in between
  void gnc_compute_weights();
  void gnc_finish_update();
  void wait_result(unsigned long long expect);
  TeamCtl *h_ctl() const { return reinterpret_cast<TeamCtl *>(h_result); }
  void layout_result();
  int small_grid = 1;
  void flush_inboxes();
  void launch_and_read(const RunArgs &args, int use_grid, bool timed, float *ms);
  // the two halves of launch_and_read (an armed launch is begun by Agent::arm and finished by iterate(true))
  struct PendingLaunch {
    RunArgs args;
    Agent *la = nullptr;
    int la_depth = 0;
    bool timed = false;
    std::chrono::steady_clock::time_point hp0, hp1;
  };
  PendingLaunch pending;
  void launch_begin(const RunArgs &args, int use_grid, bool timed, PendingLaunch &pl);
  void launch_finish(PendingLaunch &pl, float *ms);
  volatile unsigned long long *doorbell() const {
    return reinterpret_cast<volatile unsigned long long *>(h_result + kCtlDoorbellOff);
  }
  volatile unsigned long long *arm_state() const {
    return reinterpret_cast<volatile unsigned long long *>(h_result + kCtlArmStateOff);
  }
  // LARGE agents without acceleration run one iteration per launch with the gradient in k_edge_grad (edge_grad.cu)
  bool edge_grad_loop(int use_grid) const;
  unsigned ext_grad_mask(const RunArgs &args, int use_grid) const;
  std::vector<Agent *> agents;
  TeamDev T{};
  TeamCtl ctl{};
  DevBuf<unsigned long long> dBar;
  DevBuf<double> dSlots;
  DevBuf<double> dDefer;
  DevBuf<double2> dGammaTab;
  std::vector<double2> h_gamma_tab;
  std::vector<double> h_gamma_state;
  double gamma_state = 0;  // Nesterov gamma after the last executed iteration (0 after a restart)
  static double next_gamma(double g, int N) { return (1.0 + std::sqrt(1.0 + 4.0 * N * N * g * g)) / (2.0 * N); }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool team_dirty = true;
  int precond_mode = 2;  // prepare(): 2 = dense inverses built and waited for, 1 = started, 0 = buffers only (warm, iterate(false))
  int launches = 0;
  double host_prof[4] = {0, 0, 0, 0};  // diagnostics: seconds in the launch call, seconds until the result, launches
  DevBuf<long long> dProf;
  int prof_iters = 0, prof_cta = 0;
};

}  // namespace dpgo
