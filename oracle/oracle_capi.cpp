// CPU ORACLE -- test infrastructure only (see dpgo_oracle.hpp).  Plain C entry
// points so tests / bench.py can drive the oracle through ctypes.  Array
// conventions match include/dpgo_b200.h so the same numpy buffers feed both:
//   R  : m x 9, row-major 3x3 per measurement;  t : m x 3
//   X  : r x 4n column-major (pose i = columns 4i..4i+3)
//   T  : n x 3 x 4 row-major per pose (numpy [n,3,4])
#include <cstring>
#include <exception>
#include <cstdio>

#include "dpgo_oracle.hpp"

using namespace dpgo_oracle;

extern "C" {

struct orc_params {
  int d, r, num_robots;
  int method;  // 0 RTR, 1 RGD
  double rgd_stepsize;
  int rgd_use_preconditioner;
  int rtr_iterations, rtr_tcg_iterations;
  double rtr_initial_radius, gradnorm_tol;
  int acceleration, restart_interval;
  int cost_type;
  double gnc_barc, gnc_mu_step, gnc_init_mu;
  int robust_opt_num_weight_updates, robust_opt_num_resets, robust_opt_inner_iters;
  double robust_opt_min_convergence_ratio;
  int max_num_iters;
  double rel_change_tol;
  double precond_lambda;
};

struct orc_run_result {
  int iterations, terminated, weight_updates;
  double wall_seconds;
};

struct orc_opt_result {
  int success;
  double f_init, f_opt, gradnorm_init, gradnorm_opt, relative_change;
  int rtr_outer_iters, tcg_iters, rtr_rejections;
};

struct orc_status {
  int agent_id, state, instance_number, iteration_number, ready_to_terminate;
  double relative_change;
};

static Params toParams(const orc_params *p) {
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

#define ORC_TRY try {
#define ORC_CATCH                                        \
  }                                                      \
  catch (const std::exception &e) {                      \
    std::fprintf(stderr, "[oracle] %s\n", e.what());     \
    return -1;                                           \
  }                                                      \
  return 0;

void *orc_team_create(const orc_params *p) {
  try {
    return new Team(toParams(p));
  } catch (...) {
    return nullptr;
  }
}
void orc_team_destroy(void *t) { delete (Team *)t; }

int orc_add_measurements(void *t, int agent, int m, const int *r1, const int *p1, const int *r2, const int *p2,
                         const double *R, const double *tt, const double *kappa, const double *tau,
                         const double *weight, const unsigned char *fixed) {
  ORC_TRY
  Team *team = (Team *)t;
  for (int e = 0; e < m; ++e) {
    Measurement ms;
    ms.r1 = r1[e];
    ms.p1 = p1[e];
    ms.r2 = r2[e];
    ms.p2 = p2[e];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) ms.R[j * 3 + i] = R[(size_t)e * 9 + i * 3 + j];
    for (int i = 0; i < 3; ++i) ms.t[i] = tt[(size_t)e * 3 + i];
    ms.kappa = kappa[e];
    ms.tau = tau[e];
    ms.weight = weight[e];
    ms.fixedWeight = fixed[e] != 0;
    team->agent(agent).addMeasurement(ms);
  }
  ORC_CATCH
}

int orc_num_poses(void *t, int agent) { return ((Team *)t)->agent(agent).numPoses(); }

int orc_set_lifting_matrix(void *t, int agent, const double *Y) {
  ORC_TRY((Team *)t)->agent(agent).setLiftingMatrix(Y);
  ORC_CATCH
}

static void rowMajorPosesToColMajor(const double *T, int n, std::vector<double> &out) {
  out.resize((size_t)12 * n);
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 4; ++c) out[(size_t)i * 12 + c * 3 + a] = T[(size_t)i * 12 + a * 4 + c];
}

int orc_initialize(void *t, int agent, const double *T_local) {
  ORC_TRY Agent &a = ((Team *)t)->agent(agent);
  if (T_local) {
    std::vector<double> cm;
    rowMajorPosesToColMajor(T_local, a.numPoses(), cm);
    a.initialize(cm.data());
  } else {
    a.initialize(nullptr);
  }
  ORC_CATCH
}

int orc_set_iteration_number(void *t, int agent, int it) {
  ORC_TRY((Team *)t)->agent(agent).setIterationNumber(it);
  ORC_CATCH
}

int orc_set_robot_active(void *t, int agent, int robot, int active) {
  ORC_TRY((Team *)t)->agent(agent).setRobotActive(robot, active != 0);
  ORC_CATCH
}

int orc_initialize_chordal(void *t, int agent) {
  ORC_TRY((Team *)t)->agent(agent).initializeChordal();
  ORC_CATCH
}

// local-frame trajectory after initialize / initializeChordal: n x 3 x 4 row-major per pose
int orc_get_local_trajectory(void *t, int agent, double *out) {
  ORC_TRY const Mat &T = ((Team *)t)->agent(agent).localTrajectory();
  const int n = T.cols / 4;
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 4; ++c) out[(size_t)i * 12 + a * 4 + c] = T(a, 4 * i + c);
  ORC_CATCH
}

int orc_initialize_in_global_frame(void *t, int agent, const double *Tw /*3x4 row-major*/) {
  ORC_TRY std::vector<double> cm;
  rowMajorPosesToColMajor(Tw, 1, cm);
  ((Team *)t)->agent(agent).initializeInGlobalFrame(cm.data());
  ORC_CATCH
}

int orc_exchange_all(void *t) {
  ORC_TRY((Team *)t)->exchangeAll();
  ORC_CATCH
}

int orc_run(void *t, int max_iters, int threads, int stop_on_terminate, orc_run_result *out) {
  ORC_TRY TeamRunResult r = ((Team *)t)->run(max_iters, threads, stop_on_terminate != 0);
  out->iterations = r.iterations;
  out->terminated = r.terminated;
  out->weight_updates = r.weightUpdates;
  out->wall_seconds = r.wallSeconds;
  ORC_CATCH
}

int orc_run_parallel(void *t, int ticks, int threads, orc_run_result *out) {
  ORC_TRY TeamRunResult r = ((Team *)t)->runParallel(ticks, threads);
  out->iterations = r.iterations;
  out->terminated = r.terminated;
  out->weight_updates = r.weightUpdates;
  out->wall_seconds = r.wallSeconds;
  ORC_CATCH
}

int orc_agent_iterate(void *t, int agent, int do_opt) {
  ORC_TRY((Team *)t)->agent(agent).iterate(do_opt != 0);
  ORC_CATCH
}

int orc_get_x(void *t, int agent, int which, double *out) {
  ORC_TRY Agent &a = ((Team *)t)->agent(agent);
  const Mat &M = which == 0 ? a.X() : (which == 1 ? a.Yaux() : a.V());
  std::memcpy(out, M.a.data(), sizeof(double) * M.a.size());
  ORC_CATCH
}

int orc_get_opt_result(void *t, int agent, orc_opt_result *o) {
  const OptResult &r = ((Team *)t)->agent(agent).localOptResult();
  o->success = r.success;
  o->f_init = r.fInit;
  o->f_opt = r.fOpt;
  o->gradnorm_init = r.gradNormInit;
  o->gradnorm_opt = r.gradNormOpt;
  o->relative_change = r.relativeChange;
  o->rtr_outer_iters = r.rtrOuterIters;
  o->tcg_iters = r.tcgIters;
  o->rtr_rejections = r.rtrRejections;
  return 0;
}

int orc_get_status(void *t, int agent, orc_status *s) {
  const Status st = ((Team *)t)->agent(agent).getStatus();
  s->agent_id = st.agentID;
  s->state = (int)st.state;
  s->instance_number = st.instanceNumber;
  s->iteration_number = st.iterationNumber;
  s->ready_to_terminate = st.readyToTerminate;
  s->relative_change = st.relativeChange;
  return 0;
}

double orc_global_cost(void *t) { return ((Team *)t)->globalCost(); }
int orc_weight_update_count(void *t, int agent) { return ((Team *)t)->agent(agent).weightUpdateCount(); }
double orc_robust_mu(void *t, int agent) { return ((Team *)t)->agent(agent).robustCost().mu; }

// loop-closure weights in insertion order: private LCs then shared LCs
int orc_get_lc_weights(void *t, int agent, double *out, int cap) {
  PoseGraph &pg = ((Team *)t)->agent(agent).poseGraph();
  int k = 0;
  for (auto &m : pg.privateLoopClosures())
    if (k < cap) out[k++] = m.weight;
  for (auto &m : pg.sharedLoopClosures())
    if (k < cap) out[k++] = m.weight;
  return k;
}

// ---- problem-level evaluation at an arbitrary X (uses the agent's current
// neighbour poses; aux selects the auxiliary dictionary)
struct EvalCtx {
  PoseDict nbr;
};

static bool buildProblem(Team *team, int agent, bool needPre) {
  Agent &a = team->agent(agent);
  // neighbour poses as currently stored in the *other* agents (fresh exchange)
  PoseDict nbr;
  for (int b = 0; b < team->size(); ++b) {
    if (b == agent) continue;
    PoseDict d;
    if (team->agent(b).getSharedPoseDictWithNeighbor(d, agent))
      for (auto &kv : d) nbr[kv.first] = kv.second;
  }
  return a.poseGraph().constructDataMatrices(nbr, needPre, a.params().precondLambda);
}

// f, Euclidean gradient, Riemannian gradient at X
int orc_eval(void *t, int agent, const double *X, double *f, double *egrad, double *rgrad) {
  ORC_TRY Team *team = (Team *)t;
  Agent &a = team->agent(agent);
  if (!buildProblem(team, agent, false)) return -2;
  const int r = a.params().r, n = a.numPoses();
  Mat Xm(r, 4 * n), g, rg;
  std::memcpy(Xm.a.data(), X, sizeof(double) * Xm.a.size());
  QuadraticProblem prob(&a.poseGraph(), r);
  if (f) *f = prob.f(Xm);
  prob.eucGrad(Xm, g);
  if (egrad) std::memcpy(egrad, g.a.data(), sizeof(double) * g.a.size());
  if (rgrad) {
    tangentProject(Xm, g, rg);
    std::memcpy(rgrad, rg.a.data(), sizeof(double) * rg.a.size());
  }
  ORC_CATCH
}

int orc_hess(void *t, int agent, const double *X, const double *V, double *out) {
  ORC_TRY Team *team = (Team *)t;
  Agent &a = team->agent(agent);
  if (!buildProblem(team, agent, false)) return -2;
  const int r = a.params().r, n = a.numPoses();
  Mat Xm(r, 4 * n), Vm(r, 4 * n), g, H;
  std::memcpy(Xm.a.data(), X, sizeof(double) * Xm.a.size());
  std::memcpy(Vm.a.data(), V, sizeof(double) * Vm.a.size());
  QuadraticProblem prob(&a.poseGraph(), r);
  prob.eucGrad(Xm, g);
  prob.rieHess(Xm, g, Vm, H);
  std::memcpy(out, H.a.data(), sizeof(double) * H.a.size());
  ORC_CATCH
}

int orc_precond(void *t, int agent, const double *X, const double *V, double *out) {
  ORC_TRY Team *team = (Team *)t;
  Agent &a = team->agent(agent);
  if (!buildProblem(team, agent, true)) return -2;
  const int r = a.params().r, n = a.numPoses();
  Mat Xm(r, 4 * n), Vm(r, 4 * n), Z;
  std::memcpy(Xm.a.data(), X, sizeof(double) * Xm.a.size());
  std::memcpy(Vm.a.data(), V, sizeof(double) * Vm.a.size());
  QuadraticProblem prob(&a.poseGraph(), r);
  prob.precondition(Xm, Vm, Z);
  std::memcpy(out, Z.a.data(), sizeof(double) * Z.a.size());
  ORC_CATCH
}

int orc_dense_q(void *t, int agent, double *Q /* 4n x 4n col-major */, double *G /* r x 4n */) {
  ORC_TRY Team *team = (Team *)t;
  Agent &a = team->agent(agent);
  if (!buildProblem(team, agent, false)) return -2;
  Mat Qd = a.poseGraph().denseQ();
  if (Q) std::memcpy(Q, Qd.a.data(), sizeof(double) * Qd.a.size());
  if (G) std::memcpy(G, a.poseGraph().G().a.data(), sizeof(double) * a.poseGraph().G().a.size());
  ORC_CATCH
}

int orc_set_measurement_weight(void *t, int agent, int r1, int p1, int r2, int p2, double w, int fixed) {
  ORC_TRY
  Agent &a = ((Team *)t)->agent(agent);
  if (!a.setMeasurementWeight(r1, p1, r2, p2, w, fixed != 0)) return -4;
  a.poseGraph().clearDataMatrices();
  ORC_CATCH
}
int orc_compute_measurement_residual(void *t, int agent, int r1, int p1, int r2, int p2, double *res) {
  ORC_TRY Agent &a = ((Team *)t)->agent(agent);
  Measurement *m = a.poseGraph().findMeasurement(r1, p1, r2, p2);
  if (!m || !a.computeMeasurementResidual(*m, res)) return -4;
  ORC_CATCH
}

// manifold ops on raw r x 4n arrays
int orc_manifold_project(int r, int n, const double *M, double *out) {
  ORC_TRY Mat A(r, 4 * n), B;
  std::memcpy(A.a.data(), M, sizeof(double) * A.a.size());
  manifoldProject(A, B);
  std::memcpy(out, B.a.data(), sizeof(double) * B.a.size());
  ORC_CATCH
}
int orc_tangent_project(int r, int n, const double *X, const double *Z, double *out) {
  ORC_TRY Mat A(r, 4 * n), B(r, 4 * n), C;
  std::memcpy(A.a.data(), X, sizeof(double) * A.a.size());
  std::memcpy(B.a.data(), Z, sizeof(double) * B.a.size());
  tangentProject(A, B, C);
  std::memcpy(out, C.a.data(), sizeof(double) * C.a.size());
  ORC_CATCH
}
int orc_retract(int r, int n, const double *X, const double *xi, double *out) {
  ORC_TRY Mat A(r, 4 * n), B(r, 4 * n), C;
  std::memcpy(A.a.data(), X, sizeof(double) * A.a.size());
  std::memcpy(B.a.data(), xi, sizeof(double) * B.a.size());
  retract(A, B, C);
  std::memcpy(out, C.a.data(), sizeof(double) * C.a.size());
  ORC_CATCH
}

double orc_robust_weight(int cost_type, double barc, double mu, double residual) {
  RobustCost rc;
  rc.type = (CostType)cost_type;
  rc.barcSq = barc * barc;
  rc.mu = mu;
  try {
    return rc.weight(residual);
  } catch (...) {
    return -1.0;
  }
}

}  // extern "C"
