// extern "C" surface declared in include/dpgo_b200.h.  Every entry point
// converts exceptions into negative error codes; nothing here computes on the
// CPU -- a missing / unusable CUDA device surfaces as DPGO_B200_ERR_CUDA.
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "host.hpp"

using namespace dpgo;

struct dpgo_b200_agent_s {
  Agent *a;
};
struct dpgo_b200_team_s {
  Team *t;
};

static thread_local std::string g_last_error;

// seconds spent inside each entry point (dpgo_b200_debug_api_profile): lets a caller split its wall clock into
// "inside the library" and "its own host code" (the reference wrapper's message handling)
namespace {
struct ApiClock {
  std::mutex mu;
  struct Ev {
    const char *name;
    double t0, t1;   // steady_clock seconds
  };
  std::vector<Ev> ev;
};
ApiClock &api_clock() {
  static ApiClock c;
  return c;
}
inline double api_now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// OFF until a caller asks for it (dpgo_b200_debug_api_profile with reset = 2): every entry point would otherwise pay two
// clock reads and a process-wide mutex -- with one thread per robot (the per-robot e2e path: ~190 calls per global
// iteration from 8 threads) that lock is contended on the critical path.
std::atomic<bool> g_api_clock_on{false};
struct ApiTimer {
  const char *name;
  double t0;
  bool on;
  explicit ApiTimer(const char *n) : name(n), t0(0), on(g_api_clock_on.load(std::memory_order_relaxed)) {
    if (on) t0 = api_now();
  }
  ~ApiTimer() {
    if (!on) return;
    const double t1 = api_now();
    ApiClock &c = api_clock();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.ev.size() < (size_t)4 << 20) c.ev.push_back({name, t0, t1});
  }
};
}  // namespace

namespace dpgo {
ProfSection::ProfSection(const char *n) : name(n), t0(g_api_clock_on.load(std::memory_order_relaxed) ? api_now() : -1.0) {}
ProfSection::~ProfSection() {
  if (t0 < 0) return;
  const double t1 = api_now();
  ApiClock &c = api_clock();
  std::lock_guard<std::mutex> lock(c.mu);
  if (c.ev.size() < (size_t)4 << 20) c.ev.push_back({name, t0, t1});
}
}  // namespace dpgo

#define API_BEGIN   \
  ApiTimer api_timer_(__func__); \
  try {
#define API_END                                    \
  }                                                \
  catch (const Error &e) {                         \
    g_last_error = e.msg;                          \
    return e.code;                                 \
  }                                                \
  catch (const std::exception &e) {                \
    g_last_error = e.what();                       \
    return DPGO_B200_ERR_INVALID;                  \
  }                                                \
  return DPGO_B200_OK;

static Agent *A(dpgo_b200_agent_t h) {
  if (!h || !h->a) fail(DPGO_B200_ERR_INVALID, "null agent handle");
  return h->a;
}
static Team *TT(dpgo_b200_team_t h) {
  if (!h || !h->t) fail(DPGO_B200_ERR_INVALID, "null team handle");
  return h->t;
}

#pragma GCC visibility push(default)
extern "C" {

const char *dpgo_b200_version(void) { return "dpgo_b200 0.1 (sm_100a)"; }
const char *dpgo_b200_last_error(void) { return g_last_error.c_str(); }
int dpgo_b200_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}
long long dpgo_b200_kernel_launch_count(void) { return kernel_launch_count() + dense_inverse_launch_count(); }

int dpgo_b200_agent_create(int id, const dpgo_b200_params *params, int device, dpgo_b200_agent_t *out) {
  API_BEGIN
  if (!params || !out) fail(DPGO_B200_ERR_INVALID, "null argument");
  Agent *a = new Agent(id, *params, device);
  *out = new dpgo_b200_agent_s{a};
  API_END
}
int dpgo_b200_agent_destroy(dpgo_b200_agent_t h) {
  API_BEGIN
  if (h) {
    delete h->a;
    delete h;
  }
  API_END
}
int dpgo_b200_reset(dpgo_b200_agent_t h) {
  API_BEGIN
  A(h)->reset();
  API_END
}

int dpgo_b200_add_measurements(dpgo_b200_agent_t h, int m, const int *r1, const int *p1, const int *r2,
                               const int *p2, const double *R, const double *t, const double *kappa,
                               const double *tau, const double *weight, const unsigned char *fixed) {
  API_BEGIN
  Agent *a = A(h);
  for (int e = 0; e < m; ++e) {
    Meas ms;
    ms.r1 = r1[e]; ms.p1 = p1[e]; ms.r2 = r2[e]; ms.p2 = p2[e];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) ms.R[j * 3 + i] = R[(size_t)e * 9 + i * 3 + j];
    for (int i = 0; i < 3; ++i) ms.t[i] = t[(size_t)e * 3 + i];
    ms.kappa = kappa[e];
    ms.tau = tau[e];
    ms.weight = weight ? weight[e] : 1.0;
    ms.fixed = fixed ? fixed[e] != 0 : false;
    a->add_measurement(ms);
  }
  API_END
}
int dpgo_b200_num_poses(dpgo_b200_agent_t h) { return h && h->a ? h->a->n : DPGO_B200_ERR_INVALID; }
int dpgo_b200_iteration_number(dpgo_b200_agent_t h) { return h && h->a ? h->a->iter : DPGO_B200_ERR_INVALID; }
int dpgo_b200_num_neighbors(dpgo_b200_agent_t h) {
  return h && h->a ? (int)h->a->nbrs.size() : DPGO_B200_ERR_INVALID;
}
int dpgo_b200_get_neighbors(dpgo_b200_agent_t h, int *ids, int cap) {
  API_BEGIN
  int k = 0;
  for (int b : A(h)->nbrs) {
    if (k >= cap) fail(DPGO_B200_ERR_INVALID, "buffer too small");
    ids[k++] = b;
  }
  API_END
}
int dpgo_b200_measurement_counts(dpgo_b200_agent_t h, int *o, int *p, int *s) {
  API_BEGIN
  Agent *a = A(h);
  if (o) *o = (int)a->odom.size();
  if (p) *p = (int)a->plc.size();
  if (s) *s = (int)a->slc.size();
  API_END
}

int dpgo_b200_set_lifting_matrix(dpgo_b200_agent_t h, const double *Y) {
  API_BEGIN
  A(h)->set_lifting_matrix(Y);
  API_END
}
int dpgo_b200_get_lifting_matrix(dpgo_b200_agent_t h, double *Y) {
  API_BEGIN
  Agent *a = A(h);
  if (!a->have_lift) fail(DPGO_B200_ERR_STATE, "lifting matrix not set");
  std::memcpy(Y, a->ylift, sizeof(double) * a->r * 3);
  API_END
}
int dpgo_b200_initialize(dpgo_b200_agent_t h, const double *T) {
  API_BEGIN
  A(h)->initialize(T);
  API_END
}
int dpgo_b200_initialize_chordal(dpgo_b200_agent_t h) {
  API_BEGIN
  A(h)->initialize_chordal();
  API_END
}
int dpgo_b200_get_local_trajectory(dpgo_b200_agent_t h, double *out) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state == 0 || a->Tlocal.size() != (size_t)12 * a->n) fail(DPGO_B200_ERR_STATE, "no local trajectory yet");
  for (int i = 0; i < a->n; ++i)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) out[(size_t)i * 12 + r * 4 + c] = a->Tlocal[(size_t)i * 12 + c * 3 + r];
  API_END
}
int dpgo_b200_initialize_in_global_frame(dpgo_b200_agent_t h, const double *Tw) {
  API_BEGIN
  A(h)->initialize_in_global_frame(Tw);
  API_END
}

int dpgo_b200_iterate(dpgo_b200_agent_t h, int do_opt) {
  API_BEGIN
  Agent *a = A(h);
  const bool optimized = a->iterate(do_opt != 0);
  // upstream returns the success of the local solve (the wrapper logs "iteration not successful", :160-175): an
  // iterate(true) that had to be skipped because a neighbour's poses have not arrived yet advances the iteration
  // counter like any other and reports DPGO_B200_ERR_MISSING
  if (do_opt && !optimized && a->state == 2) fail(DPGO_B200_ERR_MISSING, "iterate: neighbour poses missing, local solve skipped");
  API_END
}
int dpgo_b200_get_opt_result(dpgo_b200_agent_t h, dpgo_b200_opt_result *out) {
  API_BEGIN
  A(h)->finish_opt_stats();
  *out = A(h)->opt;
  API_END
}
int dpgo_b200_get_opt_result_lazy(dpgo_b200_agent_t h, dpgo_b200_opt_result *out) {
  API_BEGIN
  *out = A(h)->opt;
  API_END
}
int dpgo_b200_get_status(dpgo_b200_agent_t h, dpgo_b200_status *out) {
  API_BEGIN
  *out = A(h)->get_status();
  API_END
}
int dpgo_b200_set_neighbor_status(dpgo_b200_agent_t h, const dpgo_b200_status *s) {
  API_BEGIN
  A(h)->team_status[s->agent_id] = *s;
  API_END
}
int dpgo_b200_set_robot_active(dpgo_b200_agent_t h, int robot, int active) {
  API_BEGIN
  Agent *a = A(h);
  if (robot < 0 || robot >= a->P.num_robots) fail(DPGO_B200_ERR_INVALID, "setRobotActive: robot id out of range");
  const bool changed = active ? a->inactive_robots.erase(robot) != 0 : a->inactive_robots.insert(robot).second;
  if (changed) {   // the robot's shared loop closures enter / leave Q, G and the preconditioner
    a->values_dirty = a->precon_dirty = true;
    if (a->team) a->team->team_dirty = true;
  }
  API_END
}
int dpgo_b200_should_terminate(dpgo_b200_agent_t h) {
  try {
    return A(h)->should_terminate() ? 1 : 0;
  } catch (const Error &e) {
    g_last_error = e.msg;
    return e.code;
  }
}
int dpgo_b200_should_update_measurement_weights(dpgo_b200_agent_t h) {
  try {
    return A(h)->should_update_weights() ? 1 : 0;
  } catch (const Error &e) {
    g_last_error = e.msg;
    return e.code;
  }
}

int dpgo_b200_get_x(dpgo_b200_agent_t h, int which, double *out) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "getX: agent not initialized");
  cuda_check(cudaSetDevice(a->device), "cudaSetDevice");
  a->materialize_lookahead();
  const DevBuf<double> &b = which == 0 ? a->dX : (which == 1 ? a->dY : a->dV);
  cuda_check(cudaMemcpy(out, b.p, sizeof(double) * a->r * 4 * a->n, cudaMemcpyDeviceToHost), "D2H X");
  API_END
}
int dpgo_b200_set_iteration_number(dpgo_b200_agent_t h, int iteration) {
  API_BEGIN
  Agent *a = A(h);
  if (iteration < 0) fail(DPGO_B200_ERR_INVALID, "iteration number must be non-negative");
  if (a->state == 2) {
    cuda_check(cudaSetDevice(a->device), "cudaSetDevice");
    a->materialize_lookahead();  // speculated steps belong to the old numbering (restart iterations depend on it)
  }
  a->drop_lookahead();
  a->iter = iteration;
  a->status.iteration_number = iteration;
  if (a->team) a->team->ctl.iter = iteration;
  API_END
}
int dpgo_b200_get_pose(dpgo_b200_agent_t h, int which, int index, double *out) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "getSharedPose: agent not initialized");
  if (index < 0 || index >= a->n || which < 0 || which > 2) fail(DPGO_B200_ERR_INVALID, "getSharedPose: index out of range");
  cuda_check(cudaSetDevice(a->device), "cudaSetDevice");
  a->materialize_lookahead();
  const DevBuf<double> &b = which == 0 ? a->dX : (which == 1 ? a->dY : a->dV);
  cuda_check(cudaMemcpy(out, b.p + (size_t)index * 4 * a->r, sizeof(double) * 4 * a->r, cudaMemcpyDeviceToHost), "D2H pose");
  API_END
}
int dpgo_b200_set_x(dpgo_b200_agent_t h, const double *X) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "setX: agent not initialized");
  cuda_check(cudaSetDevice(a->device), "cudaSetDevice");
  a->materialize_lookahead();
  a->drop_lookahead();
  const size_t bytes = sizeof(double) * a->r * 4 * a->n;
  a->resid_valid = false;
  cuda_check(cudaMemcpy(a->dX.p, X, bytes, cudaMemcpyHostToDevice), "H2D X");
  if (a->P.acceleration) {
    cuda_check(cudaMemcpy(a->dV.p, X, bytes, cudaMemcpyHostToDevice), "H2D V");
    cuda_check(cudaMemcpy(a->dY.p, X, bytes, cudaMemcpyHostToDevice), "H2D Y");
    a->team->ctl.gamma = a->team->ctl.alpha = 0;
    a->team->gamma_state = 0;
  }
  a->outbox_stale = true;
  API_END
}

int dpgo_b200_num_shared_poses(dpgo_b200_agent_t h, int nbr) {
  try {
    return (int)A(h)->my_public_frames(nbr).size();
  } catch (const Error &e) {
    g_last_error = e.msg;
    return e.code;
  }
}
int dpgo_b200_get_shared_pose_dict(dpgo_b200_agent_t h, int nbr, int aux, int *frames, double *poses, int cap,
                                   int *count) {
  API_BEGIN
  *count = A(h)->get_shared_pose_dict(nbr, aux != 0, frames, poses, cap);
  API_END
}
int dpgo_b200_update_neighbor_poses(dpgo_b200_agent_t h, int nbr, int aux, const int *frames, const double *poses,
                                    int count) {
  API_BEGIN
  A(h)->update_neighbor_poses(nbr, aux != 0, frames, poses, count);
  API_END
}
int dpgo_b200_outbox_device_ptr(dpgo_b200_agent_t h, int nbr, int aux, void **ptr, size_t *bytes) {
  API_BEGIN
  Agent *a = A(h);
  a->team->prepare();
  auto it = a->outbox_range.find(nbr);
  if (it == a->outbox_range.end()) fail(DPGO_B200_ERR_MISSING, "not a neighbour");
  *ptr = (aux ? a->d_outbox_aux() : a->d_outbox_reg()) + (size_t)it->second.first * 4 * a->r;
  *bytes = (size_t)it->second.second * 4 * a->r * sizeof(double);
  API_END
}
int dpgo_b200_inbox_device_ptr(dpgo_b200_agent_t h, int nbr, int aux, void **ptr, size_t *bytes) {
  API_BEGIN
  Agent *a = A(h);
  a->team->prepare();
  int first = -1, cnt = 0;
  for (size_t s = 0; s < a->slot_key.size(); ++s)
    if (a->slot_key[s].first == nbr) {
      if (first < 0) first = (int)s;
      ++cnt;
    }
  if (first < 0) fail(DPGO_B200_ERR_MISSING, "not a neighbour");
  *ptr = (aux ? a->d_inbox_aux() : a->d_inbox_reg()) + (size_t)first * 4 * a->r;
  *bytes = (size_t)cnt * 4 * a->r * sizeof(double);
  API_END
}
int dpgo_b200_mark_inbox_updated(dpgo_b200_agent_t h, int nbr, int aux) {
  API_BEGIN
  Agent *a = A(h);
  auto &v = aux ? a->inbox_valid_aux : a->inbox_valid_reg;
  for (size_t s = 0; s < a->slot_key.size(); ++s)
    if (a->slot_key[s].first == nbr) v[s] = 1;
  API_END
}

int dpgo_b200_update_measurement_weights(dpgo_b200_agent_t h) {
  API_BEGIN
  A(h)->update_measurement_weights();
  API_END
}
int dpgo_b200_set_measurement_weight(dpgo_b200_agent_t h, int r1, int p1, int r2, int p2, double w, int fixed) {
  API_BEGIN
  Agent *a = A(h);
  Meas *m = a->find_measurement(r1, p1, r2, p2);
  if (!m) fail(DPGO_B200_ERR_MISSING, "setMeasurementWeight: no such measurement");
  const bool fixed_changed = m->fixed != (fixed != 0);
  if (m->weight == w && !fixed_changed) return DPGO_B200_OK;
  m->weight = w;
  m->fixed = fixed != 0;
  // the weight sits in Q, G, the preconditioner and the device-side loop-closure arrays; `fixed` decides whether
  // the edge is in the re-weighting list at all
  a->values_dirty = a->precon_dirty = a->weights_host_dirty = true;
  if (fixed_changed) a->lc_dirty = true;
  if (a->team) a->team->team_dirty = true;
  API_END
}
int dpgo_b200_compute_measurement_residual(dpgo_b200_agent_t h, int r1, int p1, int r2, int p2, double *res) {
  API_BEGIN
  Agent *a = A(h);
  if (!a->compute_residual(r1, p1, r2, p2, res)) fail(DPGO_B200_ERR_MISSING, "computeMeasurementResidual: pose unavailable");
  API_END
}
double dpgo_b200_robust_weight(dpgo_b200_agent_t h, double residual) {
  return h && h->a ? h->a->robust_weight(residual) : -1.0;
}
int dpgo_b200_clear_data_matrices(dpgo_b200_agent_t h) {
  API_BEGIN
  Agent *a = A(h);
  a->values_dirty = a->precon_dirty = true;
  if (a->team) a->team->team_dirty = true;
  API_END
}
int dpgo_b200_get_lc_weights(dpgo_b200_agent_t h, double *out, int cap) {
  try {
    Agent *a = A(h);
    int k = 0;
    for (auto &m : a->plc)
      if (k < cap) out[k++] = m.weight;
    for (auto &m : a->slc)
      if (k < cap) out[k++] = m.weight;
    return k;
  } catch (const Error &e) {
    g_last_error = e.msg;
    return e.code;
  }
}
int dpgo_b200_get_shared_loop_closures(dpgo_b200_agent_t h, int *r1, int *p1, int *r2, int *p2, double *weight,
                                       unsigned char *fixed, int cap) {
  try {
    Agent *a = A(h);
    int k = 0;
    for (auto &m : a->slc) {
      if (k < cap) {
        if (r1) r1[k] = m.r1;
        if (p1) p1[k] = m.p1;
        if (r2) r2[k] = m.r2;
        if (p2) p2[k] = m.p2;
        if (weight) weight[k] = m.weight;
        if (fixed) fixed[k] = m.fixed ? 1 : 0;
      }
      ++k;
    }
    return k;
  } catch (const Error &e) {
    g_last_error = e.msg;
    return e.code;
  }
}
int dpgo_b200_weight_update_count(dpgo_b200_agent_t h) {
  return h && h->a ? h->a->weight_update_count : DPGO_B200_ERR_INVALID;
}

// ---- parity hooks -----------------------------------------------------------------
int dpgo_b200_eval(dpgo_b200_agent_t h, const double *X, double *f, double *egrad, double *rgrad) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "eval: agent not initialized");
  a->team->prepare();
  if (!a->all_inbox_valid(false)) fail(DPGO_B200_ERR_MISSING, "eval: neighbour poses missing");
  const size_t cnt = (size_t)a->r * 4 * a->n;
  const int grid = a->team->grid;
  DevBuf<double> dX, dE, dR, dP;
  dX.upload(std::vector<double>(X, X + cnt));
  dE.alloc(cnt);
  dR.alloc(cnt);
  dP.alloc((size_t)grid * 2);
  const AgentDev view = a->team->T.ag[a->local_index];
  cuda_check(launch_eval(view, dX.p, a->d_inbox_reg(), dE.p, dR.p, dP.p, grid, 0), "k_eval");
  std::vector<double> part((size_t)grid * 2);
  cuda_check(cudaMemcpy(part.data(), dP.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H partials");
  if (f) {
    double s = 0;
    for (int b = 0; b < grid; ++b) s += part[(size_t)b * 2];
    *f = s;
  }
  if (egrad) cuda_check(cudaMemcpy(egrad, dE.p, cnt * sizeof(double), cudaMemcpyDeviceToHost), "D2H egrad");
  if (rgrad) cuda_check(cudaMemcpy(rgrad, dR.p, cnt * sizeof(double), cudaMemcpyDeviceToHost), "D2H rgrad");
  API_END
}

int dpgo_b200_hess(dpgo_b200_agent_t h, const double *X, const double *V, double *out) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "hess: agent not initialized");
  a->team->prepare();
  if (!a->all_inbox_valid(false)) fail(DPGO_B200_ERR_MISSING, "hess: neighbour poses missing");
  const size_t cnt = (size_t)a->r * 4 * a->n;
  const int grid = a->team->grid;
  DevBuf<double> dX, dV, dE, dR, dP, dH;
  dX.upload(std::vector<double>(X, X + cnt));
  dV.upload(std::vector<double>(V, V + cnt));
  dE.alloc(cnt);
  dR.alloc(cnt);
  dH.alloc(cnt);
  dP.alloc((size_t)grid * 2);
  const AgentDev view = a->team->T.ag[a->local_index];
  cuda_check(launch_eval(view, dX.p, a->d_inbox_reg(), dE.p, dR.p, dP.p, grid, 0), "k_eval");
  cuda_check(launch_hess(view, dX.p, dV.p, dH.p, grid, 0), "k_hess");
  cuda_check(cudaMemcpy(out, dH.p, cnt * sizeof(double), cudaMemcpyDeviceToHost), "D2H hess");
  API_END
}

int dpgo_b200_precond(dpgo_b200_agent_t h, const double *X, const double *V, double *out) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "precond: agent not initialized");
  if (!a->need_preconditioner()) fail(DPGO_B200_ERR_STATE, "precond: preconditioner disabled by the parameters");
  a->team->prepare();
  const size_t cnt = (size_t)a->r * 4 * a->n;
  const int grid = a->team->grid;
  DevBuf<double> dX, dV, dVT, dZ;
  dX.upload(std::vector<double>(X, X + cnt));
  dV.upload(std::vector<double>(V, V + cnt));
  dVT.alloc(cnt);
  dZ.alloc(cnt);
  const AgentDev view = a->team->T.ag[a->local_index];
  cuda_check(launch_transpose_rows(dV.p, dVT.p, a->r, 4 * a->n, 0), "k_transpose_rows");
  cuda_check(launch_precond(view, dX.p, dV.p, dVT.p, dZ.p, grid, 0), "k_precond");
  cuda_check(cudaMemcpy(out, dZ.p, cnt * sizeof(double), cudaMemcpyDeviceToHost), "D2H precond");
  API_END
}

static int manifold_op(int device, int op, int r, int n, const double *Ain, const double *Bin, double *out) {
  API_BEGIN
  if (r < 3 || r > 8 || n < 0) fail(DPGO_B200_ERR_INVALID, "manifold op: bad shape");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    fail(DPGO_B200_ERR_CUDA, "no usable CUDA device (the RBCD path has no CPU fallback)");
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  if (n == 0) return DPGO_B200_OK;
  const size_t cnt = (size_t)r * 4 * n;
  DevBuf<double> dA, dB, dO;
  dA.upload(std::vector<double>(Ain, Ain + cnt));
  if (Bin) dB.upload(std::vector<double>(Bin, Bin + cnt));
  dO.alloc(cnt);
  const int grid = std::min(max_coop_grid(device), (n + 31) / 32);
  cuda_check(launch_manifold_op(op, r, n, dA.p, dB.p, dO.p, std::max(1, grid), 0), "k_manifold_op");
  cuda_check(cudaMemcpy(out, dO.p, cnt * sizeof(double), cudaMemcpyDeviceToHost), "D2H manifold op");
  API_END
}
int dpgo_b200_manifold_project(int device, int r, int n, const double *M, double *out) {
  return manifold_op(device, 0, r, n, M, nullptr, out);
}
int dpgo_b200_tangent_project(int device, int r, int n, const double *X, const double *Z, double *out) {
  return manifold_op(device, 1, r, n, X, Z, out);
}
int dpgo_b200_retract(int device, int r, int n, const double *X, const double *xi, double *out) {
  return manifold_op(device, 2, r, n, X, xi, out);
}

// ---- team ---------------------------------------------------------------------------
int dpgo_b200_team_create(int device, dpgo_b200_team_t *out) {
  API_BEGIN
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    fail(DPGO_B200_ERR_CUDA, "no usable CUDA device (the RBCD path has no CPU fallback)");
  Team *t = new Team(device);
  t->device_outbox = true;
  *out = new dpgo_b200_team_s{t};
  API_END
}
int dpgo_b200_team_step(dpgo_b200_team_t h, int selected_robot, int mode) {
  API_BEGIN
  if (mode < 0 || mode > 2) fail(DPGO_B200_ERR_INVALID, "team_step: mode must be 0, 1 or 2");
  TT(h)->step(selected_robot, mode);
  API_END
}
int dpgo_b200_team_destroy(dpgo_b200_team_t h) {
  API_BEGIN
  if (h) {
    delete h->t;
    delete h;
  }
  API_END
}
int dpgo_b200_team_add_agent(dpgo_b200_team_t h, dpgo_b200_agent_t a) {
  API_BEGIN
  TT(h)->add(A(a));
  API_END
}
int dpgo_b200_team_exchange_all(dpgo_b200_team_t h) {
  API_BEGIN
  TT(h)->exchange_all();
  API_END
}
int dpgo_b200_team_run(dpgo_b200_team_t h, int max_iters, int stop_on_terminate, dpgo_b200_run_result *out) {
  API_BEGIN
  dpgo_b200_run_result r = TT(h)->run(max_iters, stop_on_terminate != 0);
  if (out) *out = r;
  API_END
}
double dpgo_b200_team_global_cost(dpgo_b200_team_t h, int *status) {
  try {
    const double c = TT(h)->global_cost();
    if (status) *status = 0;
    return c;
  } catch (const Error &e) {
    g_last_error = e.msg;
    if (status) *status = e.code;
    return 0.0;
  }
}
int dpgo_b200_team_fabric_init(dpgo_b200_team_t h, int world, int rank) {
  API_BEGIN
  TT(h)->fabric_init(world, rank);
  API_END
}
int dpgo_b200_team_fabric_window(dpgo_b200_team_t h, void **base, size_t *bytes, void *ipc_handle_64) {
  API_BEGIN
  Team *t = TT(h);
  if (!t->window) fail(DPGO_B200_ERR_STATE, "fabric_window before fabric_init");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  if (base) *base = t->window;
  if (bytes) *bytes = t->window_bytes;
  if (ipc_handle_64) {
    cuda_check(cudaSetDevice(t->device), "cudaSetDevice");
    cudaIpcMemHandle_t hd;
    cuda_check(cudaIpcGetMemHandle(&hd, t->window), "cudaIpcGetMemHandle");
    std::memcpy(ipc_handle_64, &hd, sizeof(hd));
  }
  API_END
}
int dpgo_b200_team_fabric_import(dpgo_b200_team_t h, int peer_rank, const void *ipc_handle_64,
                                 void *same_process_base) {
  API_BEGIN
  cudaIpcMemHandle_t hd;
  if (ipc_handle_64) std::memcpy(&hd, ipc_handle_64, sizeof(hd));
  TT(h)->fabric_import(peer_rank, ipc_handle_64 ? &hd : nullptr, same_process_base);
  API_END
}
int dpgo_b200_team_fabric_route(dpgo_b200_team_t h, int robot, int neighbor, int peer_rank, size_t off_reg,
                                size_t off_aux) {
  API_BEGIN
  TT(h)->fabric_route(robot, neighbor, peer_rank, off_reg, off_aux);
  API_END
}
int dpgo_b200_team_fabric_run(dpgo_b200_team_t h, int max_iters, int stop_on_terminate, dpgo_b200_run_result *out) {
  API_BEGIN
  dpgo_b200_run_result r = TT(h)->fabric_run(max_iters, stop_on_terminate != 0);
  if (out) *out = r;
  API_END
}
int dpgo_b200_team_fabric_set_timeout(dpgo_b200_team_t h, double seconds) {
  API_BEGIN
  if (!(seconds > 0)) fail(DPGO_B200_ERR_INVALID, "fabric timeout must be positive");
  TT(h)->fab_timeout_s = seconds;
  TT(h)->team_dirty = true;
  API_END
}
int dpgo_b200_team_fabric_close(dpgo_b200_team_t h) {
  API_BEGIN
  TT(h)->fabric_close();
  API_END
}
int dpgo_b200_team_gnc_compute_weights(dpgo_b200_team_t h) {
  API_BEGIN
  TT(h)->gnc_compute_weights();
  API_END
}
int dpgo_b200_team_gnc_finish_update(dpgo_b200_team_t h) {
  API_BEGIN
  TT(h)->gnc_finish_update();
  API_END
}
int dpgo_b200_team_set_schedule(dpgo_b200_team_t h, int schedule) {
  API_BEGIN
  if (schedule != 0 && schedule != 1) fail(DPGO_B200_ERR_INVALID, "schedule must be 0 (RoundRobin) or 1 (parallel)");
  TT(h)->schedule = schedule;
  API_END
}
int dpgo_b200_team_set_grid(dpgo_b200_team_t h, int num_ctas) {
  API_BEGIN
  Team *t = TT(h);
  const int mx = max_coop_grid(t->device);
  t->grid = (num_ctas <= 0 || num_ctas > mx) ? mx : num_ctas;
  t->team_dirty = true;
  API_END
}

// ---- diagnostics (not part of the reference surface) ---------------------------------
int dpgo_b200_debug_barrier_bench(int device, int grid, int iters, int mode, float *ms) {
  API_BEGIN
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  DevBuf<unsigned long long> bar;
  DevBuf<double> slots, out;
  bar.alloc(1);
  slots.alloc((size_t)2 * grid * kRed);
  out.alloc(1);
  GridSync gs{bar.p, slots.p};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cuda_check(launch_barrier_bench(gs, 10, mode, 0u, out.p, grid, 0), "barrier_bench warmup");
  cudaEventRecord(e0, 0);
  cuda_check(launch_barrier_bench(gs, iters, mode, 10u, out.p, grid, 0), "barrier_bench");
  cudaEventRecord(e1, 0);
  cuda_check(cudaEventSynchronize(e1), "barrier_bench sync");
  cudaEventElapsedTime(ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  API_END
}
// the data matrices as the device assembled them (k_assemble_values): Q dense 4n x 4n column-major from the block-CSR
// copy AND, separately, from the ELL + overflow copy the hot phases read; G_lin = the r x 4n linear term for the
// current inbox.  A read-out for the parity test against PoseGraph::denseQ() of the oracle; nothing is computed here.
int dpgo_b200_debug_dense_q(dpgo_b200_agent_t h, double *Q_csr, double *Q_ell) {
  API_BEGIN
  Agent *a = A(h);
  cuda_check(cudaSetDevice(a->device), "cudaSetDevice");
  a->team->prepare();
  const int n = a->n;
  const size_t N = (size_t)4 * n;
  auto put = [&](double *Q, int out_pose, int col_pose, const double *blk) {
    // out_j += X_col * blk  =>  Q(4 col + p, 4 out + q) = blk(p, q)
    for (int q = 0; q < 4; ++q)
      for (int p = 0; p < 4; ++p) Q[((size_t)4 * out_pose + q) * N + (size_t)4 * col_pose + p] += blk[q * 4 + p];
  };
  if (Q_csr) {
    std::memset(Q_csr, 0, sizeof(double) * N * N);
    std::vector<double> val(a->h_q_col.size() * 16);
    cuda_check(cudaMemcpy(val.data(), a->d_q_val.p, val.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H Q");
    for (int j = 0; j < n; ++j)
      for (int e = a->h_q_rowptr[j]; e < a->h_q_rowptr[j + 1]; ++e) put(Q_csr, j, a->h_q_col[e], &val[(size_t)e * 16]);
  }
  if (Q_ell) {
    std::memset(Q_ell, 0, sizeof(double) * N * N);
    std::vector<int> ec((size_t)n * 8), orp(n + 1), oc(a->d_qo_col.n);
    std::vector<double> ev((size_t)n * 8 * 16), ov(a->d_qo_val.n);
    cuda_check(cudaMemcpy(ec.data(), a->d_qe_col.p, ec.size() * sizeof(int), cudaMemcpyDeviceToHost), "D2H");
    cuda_check(cudaMemcpy(ev.data(), a->d_qe_val.p, ev.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H");
    cuda_check(cudaMemcpy(orp.data(), a->d_qo_rowptr.p, orp.size() * sizeof(int), cudaMemcpyDeviceToHost), "D2H");
    cuda_check(cudaMemcpy(oc.data(), a->d_qo_col.p, oc.size() * sizeof(int), cudaMemcpyDeviceToHost), "D2H");
    cuda_check(cudaMemcpy(ov.data(), a->d_qo_val.p, ov.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H");
    for (int j = 0; j < n; ++j) {
      for (int k = 0; k < 8; ++k)
        if (ec[(size_t)j * 8 + k] >= 0) put(Q_ell, j, ec[(size_t)j * 8 + k], &ev[((size_t)j * 8 + k) * 16]);
      for (int e = orp[j]; e < orp[j + 1]; ++e) put(Q_ell, j, oc[e], &ov[(size_t)e * 16]);
    }
  }
  API_END
}
int dpgo_b200_debug_spd_inverse(int device, int N, const double *Ah, double *Ph, double *device_ms) {
  API_BEGIN
  if (N <= 0 || N % 32 != 0 || !Ah || !Ph) fail(DPGO_B200_ERR_INVALID, "spd_inverse: N must be a positive multiple of 32");
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  const size_t NN = (size_t)N * N;
  DevBuf<double> dA, dW;
  DevBuf<int> dinfo;
  dA.alloc(NN, false);
  dW.alloc(NN, false);
  dinfo.alloc(1);
  cuda_check(cudaMemcpy(dA.p, Ah, NN * sizeof(double), cudaMemcpyHostToDevice), "H2D A");
  cudaEvent_t e0, e1;
  cuda_check(cudaEventCreate(&e0), "event");
  cuda_check(cudaEventCreate(&e1), "event");
  cuda_check(cudaEventRecord(e0, 0), "event record");
  cuda_check(spd_inverse(dA.p, dW.p, N, dinfo.p, 0), "spd_inverse");
  cuda_check(cudaEventRecord(e1, 0), "event record");
  cuda_check(cudaDeviceSynchronize(), "spd_inverse kernels");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (device_ms) *device_ms = ms;
  int info = 0;
  cuda_check(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost), "D2H info");
  if (info != 0) fail(DPGO_B200_ERR_NUMERIC, "spd_inverse: the matrix is not positive definite (pivot " + std::to_string(info) + ")");
  cuda_check(cudaMemcpy(Ph, dA.p, NN * sizeof(double), cudaMemcpyDeviceToHost), "D2H P");
  API_END
}
// "name seconds calls\n" per entry point, restricted to the part of every call that fell inside [t_begin, t_end]
// (std::chrono::steady_clock seconds); returns the number of bytes the full report needs
int dpgo_b200_debug_api_profile(double t_begin, double t_end, char *buf, int cap, int reset) {
  if (reset == 2) g_api_clock_on.store(true);
  ApiClock &c = api_clock();
  std::lock_guard<std::mutex> lock(c.mu);
  std::map<std::string, std::pair<double, long long>> acc;
  for (const auto &e : c.ev) {
    const double a = std::max(e.t0, t_begin), b = std::min(e.t1, t_end);
    if (b <= a) continue;
    auto &x = acc[e.name];
    x.first += b - a;
    x.second += 1;
  }
  std::string out;
  for (const auto &kv : acc) {
    char line[160];
    snprintf(line, sizeof line, "%s %.9f %lld\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t k = std::min(out.size(), (size_t)cap - 1);
    std::memcpy(buf, out.data(), k);
    buf[k] = 0;
  }
  if (reset) c.ev.clear();
  return (int)out.size() + 1;
}
// The LARGE-agent gradient kernel (k_edge_grad, edge_grad.cu) on its own: f, rgrad at X against the current inbox, and
// the kernel's device time between the first CTA's start and the last CTA's end (globaltimer) -- the in-situ figure
// the roofline of the HBM-bound regime is quoted on.  flush_l2 != 0 rewrites a 512 MB scratch buffer first, so the
// records come from HBM as they do inside a step (where 10-20 GB of preconditioner have just streamed through L2).
int dpgo_b200_debug_edge_grad(dpgo_b200_agent_t h, const double *X, int flush_l2, double *f, double *rgrad,
                              double *kernel_ns, double *event_ns) {
  API_BEGIN
  Agent *a = A(h);
  if (a->state != 2) fail(DPGO_B200_ERR_STATE, "edge_grad: agent not initialized");
  a->team->prepare();
  if (!a->has_edge_arrays()) fail(DPGO_B200_ERR_STATE, "edge_grad: the agent is below the edge-record threshold");
  if (!a->all_inbox_valid(false)) fail(DPGO_B200_ERR_MISSING, "edge_grad: neighbour poses missing");
  const size_t cnt = (size_t)a->r * 4 * a->n;
  DevBuf<double> dX, dG, dR, dRT;
  DevBuf<unsigned char> scratch;
  if (X)
    dX.upload(std::vector<double>(X, X + cnt));
  dG.alloc(cnt);
  dR.alloc(cnt);
  dRT.alloc(cnt);
  cudaStream_t st = a->team->stream;
  if (flush_l2) {
    scratch.alloc((size_t)512 << 20, false);
    cuda_check(cudaMemsetAsync(scratch.p, 1, scratch.n, st), "flush L2");
  }
  EdgeGradArgs eg{};
  eg.n = a->n;
  eg.build_g = 1;
  eg.rec = a->d_er_rec.p;
  eg.inc_ptr = a->d_inc_ptr.p;
  eg.inc_item = a->d_inc_item.p;
  eg.Xin = X ? dX.p : a->dX.p;
  eg.inbox = a->d_inbox_reg();
  eg.G = dG.p;
  eg.Rg = dR.p;
  eg.RgT = dRT.p;
  eg.partials = a->d_eg_partials.p;
  eg.tmarks = a->d_eg_marks.p;
  cuda_check(cudaMemsetAsync(a->d_eg_marks.p, 0xff, sizeof(unsigned long long), st), "marks");
  cuda_check(cudaMemsetAsync(a->d_eg_marks.p + 1, 0, sizeof(unsigned long long), st), "marks");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  cuda_check(launch_edge_grad(eg, a->r, st), "k_edge_grad");
  cudaEventRecord(e1, st);
  cuda_check(cudaStreamSynchronize(st), "k_edge_grad");
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  unsigned long long marks[2];
  cuda_check(cudaMemcpy(marks, a->d_eg_marks.p, sizeof marks, cudaMemcpyDeviceToHost), "D2H marks");
  if (kernel_ns) *kernel_ns = (double)(marks[1] - marks[0]);
  if (event_ns) *event_ns = ms * 1e6;
  const int grid = edge_grad_grid(a->n);
  std::vector<double> part((size_t)grid * 2);
  cuda_check(cudaMemcpy(part.data(), a->d_eg_partials.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H partials");
  if (f) {
    double s0 = 0;
    for (int b = 0; b < grid; ++b) s0 += part[(size_t)b * 2];
    *f = s0;
  }
  if (rgrad) cuda_check(cudaMemcpy(rgrad, dR.p, cnt * sizeof(double), cudaMemcpyDeviceToHost), "D2H rgrad");
  API_END
}
int dpgo_b200_debug_host_profile(dpgo_b200_agent_t h, double *out3, int reset) {
  API_BEGIN
  Team *t = A(h)->team;
  for (int i = 0; i < 4; ++i) out3[i] = t->host_prof[i];
  if (reset) t->host_prof[0] = t->host_prof[1] = t->host_prof[2] = t->host_prof[3] = 0;
  API_END
}
// clock64 marks of CTA 0 along the RTR solve(s) of the last profiled launch (rtr_solve, team_run.cuh): 160 slots
int dpgo_b200_debug_team_marks(dpgo_b200_team_t h, long long *out160) {
  API_BEGIN
  Team *t = TT(h);
  if (t->dProf.n < 4096 + 256) fail(DPGO_B200_ERR_STATE, "debug_team_marks: run dpgo_b200_debug_team_profile first");
  cuda_check(cudaMemcpy(out160, t->dProf.p + 4096 + 64, sizeof(long long) * 160, cudaMemcpyDeviceToHost), "D2H marks");
  API_END
}
int dpgo_b200_debug_team_profile(dpgo_b200_team_t h, int iters, int cta, long long *out /* iters*16 */) {
  API_BEGIN
  Team *t = TT(h);
  t->prof_iters = iters;
  t->prof_cta = cta;
  t->dProf.alloc((size_t)4096 + 256);
  t->team_dirty = true;
  t->run(iters, false);
  cuda_check(cudaMemcpy(out, t->dProf.p, sizeof(long long) * iters * 16, cudaMemcpyDeviceToHost), "D2H prof");
  cuda_check(cudaMemcpy(out + iters * 16, t->dProf.p + 4096, sizeof(long long) * 32, cudaMemcpyDeviceToHost), "D2H dbg");
  t->prof_iters = 0;
  t->team_dirty = true;
  API_END
}

}  // extern "C"
#pragma GCC visibility pop
