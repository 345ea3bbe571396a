// =============================================================================
// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A dependency-free C++17 restatement of the RBCD hot path that dpgo_ros drives
// through DPGO::PGOAgent.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, link or call this.  The CUDA
// product under dpgo_ros_b200/csrc never includes or links anything from here.
//
// PARITY UNPINNED: the arithmetic the reference runs lives in the un-vendored,
// un-pinned dependency mit-acl/dpgo (find_package(DPGO), CMakeLists.txt:6;
// package.xml:68; README.md:9-13 "use the default branch") and, under it,
// ROPTLIB (RTRNewton / tCG, Stiefel QF retraction) and CHOLMOD.  None of that
// source is in /root/reference and the reference's only test
// (tests/testUtils.cpp) holds no optimiser golden vectors.  This file restates
// the published algorithms (RBCD / RBCD++: Tian et al. T-RO 2021; GNC: Yang et
// al. RA-L 2020; RTR-tCG: Absil et al. 2007) and is anchored on the wrapper's
// call sites, cited per function as src/PGOAgentROS.cpp:<line>.  It is
// cross-checked by an independent numpy restatement (oracle/np_oracle.py),
// finite differences and SE-Sync's published optima (tests/test_oracle.py,
// tests/test_known_optima.py).  What IS pinned against the reference itself is
// the SCHEDULE (Team::run, the GNC stages, termination): the reference's wrapper
// sources compile unmodified against the DPGO:: shim (oracle/Makefile.ref ->
// oracle/_ref/) and their real control flow reproduces Team::run's iteration
// counts and final cost exactly (tests/test_zz_reference_wrapper.py).  The
// ARITHMETIC stays unpinned.
// =============================================================================
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace dpgo_oracle {

// ----- small dense column-major matrix ---------------------------------------
struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)j * rows + i]; }
  double operator()(int i, int j) const { return a[(size_t)j * rows + i]; }
  double *col(int j) { return a.data() + (size_t)j * rows; }
  const double *col(int j) const { return a.data() + (size_t)j * rows; }
  size_t size() const { return a.size(); }
  void setZero() { std::fill(a.begin(), a.end(), 0.0); }
};
double dot(const Mat &A, const Mat &B);
double squaredNorm(const Mat &A);
void axpy(double alpha, const Mat &X, Mat &Y);  // Y += alpha X

// ----- value types crossing the boundary ----------------------
// RelativeSEMeasurement(r1,r2,p1,p2,R,t,kappa,tau) + weight, fixedWeight
// (src/utils.cpp:144-149).
struct Measurement {
  int r1 = 0, p1 = 0, r2 = 0, p2 = 0;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // column-major 3x3
  double t[3] = {0, 0, 0};
  double kappa = 1, tau = 1;
  double weight = 1;
  bool fixedWeight = false;
};

enum class OptMethod : int { RTR = 0, RGD = 1 };  // src/PGOAgentROSNode.cpp:85,90
enum class CostType : int { L2 = 0, L1 = 1, Huber = 2, TLS = 3, GM = 4, GNC_TLS = 5 };  // :178-188
enum class AgentState : int { WAIT_FOR_DATA = 0, WAIT_FOR_INITIALIZATION = 1, INITIALIZED = 2 };  // tests/testUtils.cpp:67-69

struct Params {
  int d = 3, r = 5, numRobots = 1;
  // ROptParameters (src/PGOAgentROSNode.cpp:96-100)
  OptMethod method = OptMethod::RTR;
  double RGD_stepsize = 1e-3;
  bool RGD_use_preconditioner = true;
  int RTR_iterations = 3;
  int RTR_tCG_iterations = 50;
  double RTR_initial_radius = 100;
  double gradnorm_tol = 1e-2;
  // acceleration (src/PGOAgentROSNode.cpp:126-130)
  bool acceleration = false;
  int restartInterval = 30;
  // robust cost (src/PGOAgentROSNode.cpp:174-221)
  CostType costType = CostType::L2;
  double GNCBarc = 5.0, GNCMuStep = 2.0, GNCInitMu = 1e-5;
  int robustOptNumWeightUpdates = 4;
  int robustOptNumResets = 0;
  int robustOptInnerIters = 30;
  double robustOptMinConvergenceRatio = 0.0;
  // termination (src/PGOAgentROSNode.cpp:145,226-231)
  int maxNumIters = 1000;
  double relChangeTol = 0.1;
  // preconditioner regularisation (Q + lambda I); [UPSTREAM-RECALL] 1e-1
  double precondLambda = 1e-1;
};

struct OptResult {  // mLocalOptResult, read at src/PGOAgentROS.cpp:169-172
  bool success = false;
  double fInit = 0, fOpt = 0, gradNormInit = 0, gradNormOpt = 0;
  double relativeChange = 0;
  int rtrOuterIters = 0, tcgIters = 0, rtrRejections = 0;
};

struct Status {  // PGOAgentStatus 6-arg ctor, src/utils.cpp:274-279
  int agentID = 0;
  AgentState state = AgentState::WAIT_FOR_DATA;
  int instanceNumber = 0;
  int iterationNumber = 0;
  bool readyToTerminate = false;
  double relativeChange = 0;
};

// ----- robust cost (a8) -------------------------------------------------------
struct RobustCost {
  CostType type = CostType::L2;
  double barcSq = 25.0, muStep = 2.0, initMu = 1e-5, mu = 1e-5;
  void configure(const Params &p);
  void reset() { mu = initMu; }
  void update();                    // mu <- mu * step  (GNC_TLS)
  double weight(double r) const;    // called at src/PGOAgentROS.cpp:1050
};

// ----- block-sparse SPD Cholesky at pose granularity (4x4 blocks) -------------
// Stand-in for the CHOLMOD factorisation of Q + lambda I (a6).
class BlockCholesky {
 public:
  // A given as lower+upper block map: rows[j] sorted block-row ids of column j
  // (full symmetric pattern, including diagonal), vals 16 doubles per block,
  // column-major inside the block.
  void factor(int n, const std::vector<std::vector<int>> &rows, const std::vector<std::vector<double>> &vals);
  // In-place solve of (A) Z^T = V^T for an r x 4n column-major V (each row of V
  // is one right-hand side).
  void solveRows(Mat &V) const;
  size_t nnzBlocks() const;

 private:
  int n_ = 0;
  std::vector<int> perm_, iperm_;            // perm_[k] = original pose eliminated k-th
  std::vector<std::vector<int>> lrows_;      // strictly-lower structure per (permuted) column
  std::vector<std::vector<double>> lvals_;   // 16 doubles per entry
  std::vector<double> ldiag_;                // 16 doubles per column: lower-triangular L_kk
};

// ----- Lifted SE manifold ops (a5) ---------------------------------------------
// X is r x 4n column-major; pose i occupies columns [4i, 4i+3) (Y_i, r x 3) and
// 4i+3 (p_i).
void projectToStiefel(const double *M, int r, double *out);       // U V^T via one-sided Jacobi SVD
void manifoldProject(const Mat &M, Mat &out);                      // per-pose projectToStiefel; translations copied
void tangentProject(const Mat &X, const Mat &Z, Mat &out);         // Z_Y - Y sym(Y^T Z_Y); translations copied
void retract(const Mat &X, const Mat &xi, Mat &out);               // qf(Y + xi_Y) (diag R > 0); p + xi_p

// ----- pose graph + data matrices (a4) -----------------------------------------
struct PoseKey {
  int robot, frame;
  bool operator<(const PoseKey &o) const { return robot < o.robot || (robot == o.robot && frame < o.frame); }
  bool operator==(const PoseKey &o) const { return robot == o.robot && frame == o.frame; }
};
using PoseDict = std::map<PoseKey, std::vector<double>>;  // r*4 doubles, column-major r x 4

class PoseGraph {
 public:
  PoseGraph(int id, int r, int d) : id_(id), r_(r), d_(d) {}
  void addMeasurement(const Measurement &m);   // src/PGOAgentROS.cpp:277,1307
  bool hasMeasurement(int r1, int p1, int r2, int p2) const;
  int n() const { return n_; }
  int numOdometry() const { return (int)odom_.size(); }
  int numPrivateLoopClosures() const { return (int)privateLC_.size(); }
  int numSharedLoopClosures() const { return (int)sharedLC_.size(); }
  int numMeasurements() const { return numOdometry() + numPrivateLoopClosures() + numSharedLoopClosures(); }
  std::vector<Measurement> &odometry() { return odom_; }
  std::vector<Measurement> &privateLoopClosures() { return privateLC_; }
  std::vector<Measurement> &sharedLoopClosures() { return sharedLC_; }
  const std::vector<Measurement> &sharedLoopClosures() const { return sharedLC_; }
  const std::set<int> &neighbors() const { return nbrs_; }
  // my frames that appear in a shared loop closure with `nbr` (sorted)
  std::vector<int> myPublicPoseIDs(int nbr) const;
  // neighbour frames this agent needs (sorted by PoseKey)
  std::vector<PoseKey> neighborPublicPoseIDs() const;
  Measurement *findMeasurement(int r1, int p1, int r2, int p2);

  void clearDataMatrices() { haveQ_ = false; havePrecon_ = false; }  // src/PGOAgentROS.cpp:1351
  // setRobotActive(id, false): the shared loop closures with that neighbour leave the problem (upstream's default,
  // useInactiveNeighbors(false); the alternative is commented out in the wrapper, src/PGOAgentROS.cpp:151-156)
  void setNeighborActive(int robot, bool active) {
    if (active ? inactive_.erase(robot) != 0 : inactive_.insert(robot).second) clearDataMatrices();
  }
  bool neighborActive(int robot) const { return inactive_.count(robot) == 0; }
  // Builds Q if stale; always rebuilds G from `nbrPoses`.  Returns false if a
  // needed neighbour pose is missing.
  bool constructDataMatrices(const PoseDict &nbrPoses, bool needPreconditioner, double lambda);
  // out = X Q  (r x 4n)
  void applyQ(const Mat &X, Mat &out) const;
  const Mat &G() const { return G_; }
  const BlockCholesky &preconditioner() const { return chol_; }
  // dense copy of Q (4n x 4n) -- test hook
  Mat denseQ() const;

 private:
  void buildQ();
  int id_, r_, d_;
  int n_ = 0;
  std::vector<Measurement> odom_, privateLC_, sharedLC_;
  std::set<int> nbrs_, inactive_;
  std::set<std::pair<std::pair<int, int>, std::pair<int, int>>> have_;
  bool haveQ_ = false, havePrecon_ = false;
  // block-CSC of Q: for column block j: sorted row blocks + values
  std::vector<std::vector<int>> qrows_;
  std::vector<std::vector<double>> qvals_;
  Mat G_;
  BlockCholesky chol_;
};

// ----- QuadraticProblem / QuadraticOptimizer (a2, a3) ---------------------------
class QuadraticProblem {
 public:
  QuadraticProblem(const PoseGraph *pg, int r) : pg_(pg), r_(r) {}
  double f(const Mat &X) const;                       // 0.5 <XQ, X> + <G, X>
  void eucGrad(const Mat &X, Mat &g) const;           // XQ + G
  void rieGrad(const Mat &X, Mat &g) const;           // Proj_X(XQ + G)
  double rieGradNorm(const Mat &X) const;
  void rieHess(const Mat &X, const Mat &egrad, const Mat &V, Mat &out) const;  // Proj(VQ - V sym(Y^T egrad_Y))
  void precondition(const Mat &X, const Mat &V, Mat &out) const;               // Proj_X(V (Q + lambda I)^-1)
  int n() const { return pg_->n(); }

 private:
  const PoseGraph *pg_;
  int r_;
};

class QuadraticOptimizer {
 public:
  QuadraticOptimizer(const QuadraticProblem *p, const Params &params) : prob_(p), params_(params) {}
  Mat optimize(const Mat &Y);
  const OptResult &result() const { return result_; }

 private:
  Mat trustRegion(const Mat &Yinit);
  Mat gradientDescent(const Mat &Yinit);
  // one ROPTLIB-style RTRNewton run; returns final iterate; sets `accepted` of the last step
  Mat rtrRun(const Mat &x0, int maxOuter, double initialDelta, double maxDelta, bool *lastAccepted);
  const QuadraticProblem *prob_;
  Params params_;
  OptResult result_;
};

// ----- the agent (a1, a7-a10) ---------------------------------------------------
class Agent {
 public:
  Agent(int id, const Params &p);
  int id() const { return id_; }
  const Params &params() const { return params_; }
  int numPoses() const { return pg_->n(); }
  int iterationNumber() const { return iter_; }
  void setIterationNumber(int it) { iter_ = it; }  // what the RECOVER handler does to mIterationNumber (:1196)
  AgentState state() const { return state_; }
  PoseGraph &poseGraph() { return *pg_; }

  void addMeasurement(const Measurement &m);                    // src/PGOAgentROS.cpp:277
  void setLiftingMatrix(const double *Y /* r x d col-major */);   // :928
  // Local initialisation (:348).  T = d x (d+1) per pose, column-major, n poses,
  // in the robot-local frame; null => chain odometry from identity.
  void initialize(const double *T_local);
  // local_initialization_method "Chordal" (src/PGOAgentROSNode.cpp:106-112; the demos' default,
  // launch/dpgo_demo.launch:9): two linear solves over the robot's own odometry + private loop closures with pose 0
  // fixed to the identity -- rotations from the chordal relaxation (then the polar factor of every 3x3 block),
  // translations from the resulting linear least squares.  [UPSTREAM-RECALL of the method, not of its code.]
  void initializeChordal();
  const Mat &localTrajectory() const { return Tlocal_; }
  // :353,358 -- T_world_robot is d x (d+1) column-major.
  void initializeInGlobalFrame(const double *T_world_robot);
  bool iterate(bool doOptimization);                            // :160 (true), :1185 (false)
  void reset();                                                 // :223

  // a9 -- :666-668 / :1276-1278
  bool getSharedPoseDictWithNeighbor(PoseDict &out, int nbr) const;
  bool getAuxSharedPoseDictWithNeighbor(PoseDict &out, int nbr) const;
  void updateNeighborPoses(int nbr, const PoseDict &poses);
  void updateAuxNeighborPoses(int nbr, const PoseDict &poses);
  std::vector<int> getNeighbors() const;

  // a10
  Status getStatus() const;                                     // :616
  void setNeighborStatus(const Status &s) { teamStatus_[s.agentID] = s; }  // :965
  // setRobotActive (:382 ... :1582): deactivated robots are left out of the leader's termination / re-weighting tests
  void setRobotActive(int robot, bool active) {
    if (active) inactive_.erase(robot); else inactive_.insert(robot);
    pg_->setNeighborActive(robot, active);
  }
  bool shouldTerminate() const;                                 // :208
  bool shouldUpdateMeasurementWeights() const;                  // :210
  // a8
  void updateMeasurementWeights();                              // :1218
  bool setMeasurementWeight(int r1, int p1, int r2, int p2, double w, bool fixed);  // :1341
  bool computeMeasurementResidual(const Measurement &m, double *residual) const;    // :1049
  RobustCost &robustCost() { return robust_; }

  const Mat &X() const { return X_; }
  const OptResult &localOptResult() const { return optResult_; }
  int weightUpdateCount() const { return weightUpdateCount_; }
  int robustOptInnerIter() const { return robustInnerIter_; }
  double gamma() const { return gamma_; }
  // full local objective incl. shared edges evaluated with current neighbour poses: 2f reporting helper
  const Mat &Yaux() const { return Y_; }
  const Mat &V() const { return V_; }

 private:
  bool updateX(bool doOptimization, bool acceleration);
  void initializeAcceleration();
  void updateGamma();
  void updateAlpha();
  void updateY();
  void updateV();
  bool shouldRestart() const;
  void restartNesterov(bool doOptimization);

  int id_;
  Params params_;
  std::shared_ptr<PoseGraph> pg_;
  AgentState state_ = AgentState::WAIT_FOR_DATA;
  int instance_ = 0, iter_ = 0;
  Mat X_, Xprev_, Xinit_, Y_, V_;
  Mat Tlocal_;  // d x (d+1)n local-frame initial trajectory
  bool haveLift_ = false;
  Mat YLift_;
  double gamma_ = 0, alpha_ = 0;
  PoseDict nbrPoses_, nbrAuxPoses_;
  Status status_;
  std::map<int, Status> teamStatus_;
  std::set<int> inactive_;
  OptResult optResult_;
  RobustCost robust_;
  int weightUpdateCount_ = 0, robustInnerIter_ = 0;
};

// ----- in-process replay of the wrapper's synchronous schedule ------------------
// One UPDATE token per global iteration, RoundRobin (src/PGOAgentROS.cpp:464-472);
// non-selected robots iterate(false) at once (:1185) and publish; the selected
// robot waits for its neighbours' poses of this iteration when accelerated
// (:136-149), then iterate(true) (:160); the leader (robot 0) evaluates
// shouldTerminate / shouldUpdateMeasurementWeights after its own turn (:207-217).
struct TeamRunResult {
  int iterations = 0;
  bool terminated = false;      // shouldTerminate() fired
  int weightUpdates = 0;
  double wallSeconds = 0;
};

class Team {
 public:
  Team(const Params &p);
  Agent &agent(int i) { return *agents_[i]; }
  int size() const { return (int)agents_.size(); }
  // exchange all public (+aux) poses between every neighbour pair
  void exchangeAll();
  // run up to maxIters more global iterations (or until termination)
  TeamRunResult run(int maxIters, int numThreads, bool stopOnTerminate);
  // The asynchronous mode (ASAPP: runOnceAsynchronous, src/PGOAgentROS.cpp:119-127; every robot optimises on its
  // own clock against the latest neighbour poses it has received, no acceleration, no termination test --
  // SURVEY 3.3) restated as its deterministic equal-rate / unit-delay schedule: in every tick ALL robots run
  // iterate(true) against the neighbour poses published in the previous tick, then all publish.
  TeamRunResult runParallel(int ticks, int numThreads);
  int nextSelected() const { return selected_; }
  // total cost 2 f over the whole team graph at the current X (each edge once)
  double globalCost() const;

 private:
  void deliver(int from);  // publish `from`'s public (+aux) poses to its neighbours
  Params params_;
  std::vector<std::unique_ptr<Agent>> agents_;
  int selected_ = 0;
};

}  // namespace dpgo_oracle
