// =============================================================================
// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see dpgo_oracle.hpp).
//
// The per-robot slice of the C ABI of include/dpgo_b200.h (the entry points the DPGO:: shim in
// include/DPGO/PGOAgent.h calls) implemented on the CPU oracle.  It exists for ONE purpose: to let the
// reference's UNMODIFIED wrapper (src/PGOAgentROS.cpp etc., built by oracle/Makefile.ref against the ROS
// stand-in of tests/cpp/ros_stub) run to completion in the CPU-only container, so that
//   * the stand-in, the shim and the wrapper's protocol can be debugged without a GPU, and
//   * the same wrapper binary linked against libdpgo_b200.so on a B200 has a CPU run of the very same
//     control flow to be compared with (tests/test_zz_reference_wrapper.py).
// It is linked ONLY into oracle/_ref/dpgo_ros_inproc_oracle.  libdpgo_b200.so, the shim and the Python
// package never see it; the product has no CPU path.
// =============================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/dpgo_b200.h"
#include "dpgo_oracle.hpp"

using namespace dpgo_oracle;

struct dpgo_b200_agent_s {
  Agent *a;
};

static thread_local std::string g_err;

// same per-entry-point clock as the CUDA library's (dpgo_b200_debug_api_profile), so that the wrapper's run time
// splits into "library" and "wrapper host code" the same way on both back ends
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
std::atomic<bool> g_api_clock_on{false};  // switched on by dpgo_b200_debug_api_profile(..., reset = 2), as in the CUDA library
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
extern "C" int dpgo_b200_debug_api_profile(double t_begin, double t_end, char *buf, int cap, int reset) {
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

#define TRY_                     \
  ApiTimer api_timer_(__func__); \
  try {
#define CATCH_                         \
  }                                    \
  catch (const std::exception &e) {    \
    g_err = e.what();                  \
    return DPGO_B200_ERR_INVALID;      \
  }                                    \
  return DPGO_B200_OK;

static Params toParams(const dpgo_b200_params *p) {
  Params q;
  q.d = p->d;
  q.r = p->r;
  q.numRobots = p->num_robots;
  q.method = p->method == 0 ? OptMethod::RTR : OptMethod::RGD;
  q.RGD_stepsize = p->rgd_stepsize;
  q.RGD_use_preconditioner = p->rgd_use_preconditioner != 0;
  q.RTR_iterations = p->rtr_iterations;
  q.RTR_tCG_iterations = p->rtr_tcg_iterations;
  q.RTR_initial_radius = p->rtr_initial_radius;
  q.gradnorm_tol = p->gradnorm_tol;
  q.acceleration = p->acceleration != 0;
  q.restartInterval = p->restart_interval;
  q.costType = (CostType)p->cost_type;
  q.GNCBarc = p->gnc_barc;
  q.GNCMuStep = p->gnc_mu_step;
  q.GNCInitMu = p->gnc_init_mu;
  q.robustOptNumWeightUpdates = p->robust_opt_num_weight_updates;
  q.robustOptNumResets = p->robust_opt_num_resets;
  q.robustOptInnerIters = p->robust_opt_inner_iters;
  q.robustOptMinConvergenceRatio = p->robust_opt_min_convergence_ratio;
  q.maxNumIters = p->max_num_iters;
  q.relChangeTol = p->rel_change_tol;
  q.precondLambda = p->precond_lambda;
  return q;
}

static void rowMajorPosesToColMajor(const double *T, int n, std::vector<double> &out) {
  out.resize((size_t)12 * n);
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 4; ++c) out[(size_t)i * 12 + c * 3 + a] = T[(size_t)i * 12 + a * 4 + c];
}

static const Mat &pick(Agent *a, int which) { return which == 0 ? a->X() : (which == 1 ? a->Yaux() : a->V()); }

extern "C" {

const char *dpgo_b200_version(void) { return "dpgo_b200 ABI on the CPU oracle (test infrastructure)"; }
const char *dpgo_b200_last_error(void) { return g_err.c_str(); }
int dpgo_b200_device_count(void) { return 0; }
long long dpgo_b200_kernel_launch_count(void) { return 0; }

int dpgo_b200_agent_create(int id, const dpgo_b200_params *params, int /*device*/, dpgo_b200_agent_t *out) {
  TRY_
  if (!params || !out) throw std::runtime_error("null argument");
  *out = new dpgo_b200_agent_s{new Agent(id, toParams(params))};
  CATCH_
}
int dpgo_b200_agent_destroy(dpgo_b200_agent_t h) {
  if (h) {
    delete h->a;
    delete h;
  }
  return DPGO_B200_OK;
}
int dpgo_b200_reset(dpgo_b200_agent_t h) {
  TRY_ h->a->reset();
  CATCH_
}
int dpgo_b200_add_measurements(dpgo_b200_agent_t h, int m, const int *r1, const int *p1, const int *r2, const int *p2,
                               const double *R, const double *t, const double *kappa, const double *tau,
                               const double *weight, const unsigned char *fixed) {
  TRY_
  for (int e = 0; e < m; ++e) {
    Measurement ms;
    ms.r1 = r1[e];
    ms.p1 = p1[e];
    ms.r2 = r2[e];
    ms.p2 = p2[e];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) ms.R[j * 3 + i] = R[(size_t)e * 9 + i * 3 + j];
    for (int i = 0; i < 3; ++i) ms.t[i] = t[(size_t)e * 3 + i];
    ms.kappa = kappa[e];
    ms.tau = tau[e];
    ms.weight = weight ? weight[e] : 1.0;
    ms.fixedWeight = fixed ? fixed[e] != 0 : false;
    h->a->addMeasurement(ms);
  }
  CATCH_
}
int dpgo_b200_num_poses(dpgo_b200_agent_t h) { return h->a->numPoses(); }
int dpgo_b200_iteration_number(dpgo_b200_agent_t h) { return h->a->iterationNumber(); }
int dpgo_b200_set_iteration_number(dpgo_b200_agent_t h, int it) {
  TRY_ h->a->setIterationNumber(it);
  CATCH_
}
int dpgo_b200_set_lifting_matrix(dpgo_b200_agent_t h, const double *Y) {
  TRY_ h->a->setLiftingMatrix(Y);
  CATCH_
}
int dpgo_b200_initialize(dpgo_b200_agent_t h, const double *T) {
  TRY_
  if (T) {
    std::vector<double> cm;
    rowMajorPosesToColMajor(T, h->a->numPoses(), cm);
    h->a->initialize(cm.data());
  } else {
    h->a->initialize(nullptr);
  }
  CATCH_
}
int dpgo_b200_initialize_chordal(dpgo_b200_agent_t h) {
  TRY_ h->a->initializeChordal();
  CATCH_
}
int dpgo_b200_get_local_trajectory(dpgo_b200_agent_t h, double *out) {
  TRY_
  const Mat &T = h->a->localTrajectory();
  if (h->a->state() == AgentState::WAIT_FOR_DATA || T.cols != 4 * h->a->numPoses()) {
    g_err = "no local trajectory yet";
    return DPGO_B200_ERR_STATE;
  }
  for (int i = 0; i < T.cols / 4; ++i)
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 4; ++c) out[(size_t)i * 12 + a * 4 + c] = T(a, 4 * i + c);
  CATCH_
}
int dpgo_b200_initialize_in_global_frame(dpgo_b200_agent_t h, const double *Tw) {
  TRY_
  std::vector<double> cm;
  rowMajorPosesToColMajor(Tw, 1, cm);
  h->a->initializeInGlobalFrame(cm.data());
  CATCH_
}
int dpgo_b200_iterate(dpgo_b200_agent_t h, int do_opt) {
  TRY_
  const bool ok = h->a->iterate(do_opt != 0);
  if (!ok && do_opt && h->a->state() == AgentState::INITIALIZED) {
    g_err = "iterate: neighbour poses missing";
    return DPGO_B200_ERR_MISSING;
  }
  CATCH_
}
static void fillOpt(Agent *a, dpgo_b200_opt_result *o) {
  const OptResult &q = a->localOptResult();
  o->success = q.success;
  o->f_init = q.fInit;
  o->f_opt = q.fOpt;
  o->gradnorm_init = q.gradNormInit;
  o->gradnorm_opt = q.gradNormOpt;
  o->relative_change = q.relativeChange;
  o->rtr_outer_iters = q.rtrOuterIters;
  o->tcg_iters = q.tcgIters;
  o->rtr_rejections = q.rtrRejections;
}
int dpgo_b200_get_opt_result(dpgo_b200_agent_t h, dpgo_b200_opt_result *o) {
  fillOpt(h->a, o);
  return DPGO_B200_OK;
}
int dpgo_b200_get_opt_result_lazy(dpgo_b200_agent_t h, dpgo_b200_opt_result *o) {
  fillOpt(h->a, o);
  return DPGO_B200_OK;
}
int dpgo_b200_get_status(dpgo_b200_agent_t h, dpgo_b200_status *s) {
  const Status q = h->a->getStatus();
  s->agent_id = q.agentID;
  s->state = (int)q.state;
  s->instance_number = q.instanceNumber;
  s->iteration_number = q.iterationNumber;
  s->ready_to_terminate = q.readyToTerminate;
  s->relative_change = q.relativeChange;
  return DPGO_B200_OK;
}
int dpgo_b200_set_neighbor_status(dpgo_b200_agent_t h, const dpgo_b200_status *s) {
  Status q;
  q.agentID = s->agent_id;
  q.state = (AgentState)s->state;
  q.instanceNumber = s->instance_number;
  q.iterationNumber = s->iteration_number;
  q.readyToTerminate = s->ready_to_terminate != 0;
  q.relativeChange = s->relative_change;
  h->a->setNeighborStatus(q);
  return DPGO_B200_OK;
}
int dpgo_b200_set_robot_active(dpgo_b200_agent_t h, int robot, int active) {
  h->a->setRobotActive(robot, active != 0);
  return DPGO_B200_OK;
}
int dpgo_b200_should_terminate(dpgo_b200_agent_t h) { return h->a->shouldTerminate() ? 1 : 0; }
int dpgo_b200_should_update_measurement_weights(dpgo_b200_agent_t h) { return h->a->shouldUpdateMeasurementWeights() ? 1 : 0; }

int dpgo_b200_get_x(dpgo_b200_agent_t h, int which, double *out) {
  if (h->a->state() != AgentState::INITIALIZED) return DPGO_B200_ERR_STATE;
  const Mat &M = pick(h->a, which);
  std::memcpy(out, M.a.data(), sizeof(double) * M.a.size());
  return DPGO_B200_OK;
}
int dpgo_b200_get_pose(dpgo_b200_agent_t h, int which, int index, double *out) {
  if (h->a->state() != AgentState::INITIALIZED) return DPGO_B200_ERR_STATE;
  const Mat &M = pick(h->a, which);
  if (index < 0 || 4 * index + 3 >= M.cols) return DPGO_B200_ERR_INVALID;
  std::memcpy(out, M.col(4 * index), sizeof(double) * 4 * M.rows);
  return DPGO_B200_OK;
}

int dpgo_b200_num_shared_poses(dpgo_b200_agent_t h, int nbr) { return (int)h->a->poseGraph().myPublicPoseIDs(nbr).size(); }
int dpgo_b200_get_shared_pose_dict(dpgo_b200_agent_t h, int nbr, int aux, int *frames, double *poses, int cap, int *count) {
  TRY_
  PoseDict d;
  const bool ok = aux ? h->a->getAuxSharedPoseDictWithNeighbor(d, nbr) : h->a->getSharedPoseDictWithNeighbor(d, nbr);
  if (!ok) {
    g_err = "getSharedPoseDict: agent not initialized";
    return DPGO_B200_ERR_STATE;
  }
  if ((int)d.size() > cap) throw std::runtime_error("buffer too small");
  const size_t len = (size_t)4 * h->a->params().r;
  int k = 0;
  for (const auto &kv : d) {   // PoseDict is ordered by (robot, frame)
    frames[k] = kv.first.frame;
    std::memcpy(poses + (size_t)k * len, kv.second.data(), sizeof(double) * len);
    ++k;
  }
  *count = k;
  CATCH_
}
int dpgo_b200_update_neighbor_poses(dpgo_b200_agent_t h, int nbr, int aux, const int *frames, const double *poses, int count) {
  TRY_
  const size_t len = (size_t)4 * h->a->params().r;
  PoseDict d;
  for (int k = 0; k < count; ++k) d[{nbr, frames[k]}] = std::vector<double>(poses + (size_t)k * len, poses + (size_t)(k + 1) * len);
  if (aux)
    h->a->updateAuxNeighborPoses(nbr, d);
  else
    h->a->updateNeighborPoses(nbr, d);
  CATCH_
}

int dpgo_b200_update_measurement_weights(dpgo_b200_agent_t h) {
  TRY_ h->a->updateMeasurementWeights();
  CATCH_
}
int dpgo_b200_set_measurement_weight(dpgo_b200_agent_t h, int r1, int p1, int r2, int p2, double w, int fixed) {
  return h->a->setMeasurementWeight(r1, p1, r2, p2, w, fixed != 0) ? DPGO_B200_OK : DPGO_B200_ERR_MISSING;
}
int dpgo_b200_compute_measurement_residual(dpgo_b200_agent_t h, int r1, int p1, int r2, int p2, double *res) {
  const Measurement *m = h->a->poseGraph().findMeasurement(r1, p1, r2, p2);
  if (!m) return DPGO_B200_ERR_MISSING;
  return h->a->computeMeasurementResidual(*m, res) ? DPGO_B200_OK : DPGO_B200_ERR_MISSING;
}
double dpgo_b200_robust_weight(dpgo_b200_agent_t h, double residual) { return h->a->robustCost().weight(residual); }
int dpgo_b200_clear_data_matrices(dpgo_b200_agent_t h) {
  h->a->poseGraph().clearDataMatrices();
  return DPGO_B200_OK;
}
int dpgo_b200_get_lc_weights(dpgo_b200_agent_t h, double *out, int cap) {
  int k = 0;
  for (const auto &m : h->a->poseGraph().privateLoopClosures())
    if (k < cap) out[k++] = m.weight;
  for (const auto &m : h->a->poseGraph().sharedLoopClosures())
    if (k < cap) out[k++] = m.weight;
  return k;
}
int dpgo_b200_weight_update_count(dpgo_b200_agent_t h) { return h->a->weightUpdateCount(); }

}  // extern "C"
