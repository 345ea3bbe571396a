/*
 * DPGO/PGOAgent.h -- source-compatible stand-in for mit-acl/dpgo's PGOAgent, the base class of
 * dpgo_ros's wrapper: `class PGOAgentROS : public PGOAgent` (include/dpgo_ros/PGOAgentROS.h:121),
 * constructed as PGOAgent(ID, params) (src/PGOAgentROS.cpp:26).  Every method the wrapper calls and
 * every protected member it touches is here under the same name; the hot methods
 * forward to libdpgo_b200.so (include/dpgo_b200.h), where the arithmetic runs on the GPU:
 *
 *   iterate(bool)                       -> dpgo_b200_iterate                  src/PGOAgentROS.cpp:160, 1185
 *   getSharedPoseDictWithNeighbor / Aux -> dpgo_b200_get_shared_pose_dict     :666-668
 *   updateNeighborPoses / Aux           -> dpgo_b200_update_neighbor_poses    :1276-1278
 *   updateMeasurementWeights            -> dpgo_b200_update_measurement_weights  :1218
 *   computeMeasurementResidual          -> dpgo_b200_compute_measurement_residual :1049
 *
 * What stays on the host is bookkeeping (status maps, active-robot flags) and the once-per-round
 * global-frame read-out (rounding).  Header-only; link with -ldpgo_b200.
 *
 * Threading: in synchronous mode the wrapper is single-threaded.  With `asynchronous = true`
 * (src/PGOAgentROSNode.cpp:80-93) this class owns the optimisation thread, as upstream does: it is started by
 * initializeInGlobalFrame, runs iterate(true) on a Poisson clock of rate asynchronousOptimizationRate and raises
 * mPublishAsynchronousRequested, which runOnceAsynchronous polls (src/PGOAgentROS.cpp:119-127).  Every public method
 * takes the agent's mutex, so the wrapper's callbacks may run next to that thread.
 */
#ifndef DPGO_SHIM_PGOAGENT_H
#define DPGO_SHIM_PGOAGENT_H

#include <atomic>
#include <chrono>
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "DPGO/DPGO_robust.h"
#include "DPGO/DPGO_types.h"
#include "DPGO/DPGO_utils.h"
#include "DPGO/PGOLogger.h"
#include "DPGO/PoseGraph.h"
#include "DPGO/RelativeSEMeasurement.h"
#include "dpgo_b200.h"

namespace DPGO {

// fields set in src/PGOAgentROSNode.cpp:72-232; defaults of launch/PGOAgent.launch:9-50
class PGOAgentParameters {
 public:
  unsigned d, r, numRobots;
  bool asynchronous = false;                       // :80
  double asynchronousOptimizationRate = 1.0;       // :92
  ROptParameters localOptimizationParams;          // :82-100
  InitializationMethod localInitializationMethod = InitializationMethod::Odometry;  // :106-112
  bool multirobotInitialization = true;            // :120
  bool acceleration = false;                       // :126
  unsigned restartInterval = 30;                   // :129
  RobustCostParameters robustCostParams;           // :178-211
  int robustOptNumWeightUpdates = 4;               // :212 (int: read with ros::param::get)
  int robustOptNumResets = 0;                      // :213
  unsigned robustOptInnerIters = 30;               // :217
  double robustOptMinConvergenceRatio = 0.8;       // :214
  unsigned robustInitMinInliers = 2;               // :220
  unsigned maxNumIters = 1000;                     // :226-231
  double relChangeTol = 5e-3;                      // :145
  bool verbose = false;                            // :148
  bool logData = false;                            // :169
  std::string logDirectory;                        // :170-172
  int device = 0;                                  // (shim) CUDA device this agent lives on
  double preconditionerShift = 0.1;                // (shim) lambda of (Q + lambda I)^-1

  PGOAgentParameters(unsigned dIn, unsigned rIn, unsigned numRobotsIn = 1) : d(dIn), r(rIn), numRobots(numRobotsIn) {}

  friend std::ostream &operator<<(std::ostream &os, const PGOAgentParameters &p) {   // PGOAgentROS.h:91
    os << "PGOAgent parameters: d " << p.d << ", r " << p.r << ", robots " << p.numRobots << ", "
       << (p.asynchronous ? "asynchronous" : "synchronous") << ", local solver "
       << (p.localOptimizationParams.method == ROptParameters::ROptMethod::RTR ? "RTR" : "RGD") << ", acceleration "
       << p.acceleration << " (restart " << p.restartInterval << "), rel. change tol " << p.relChangeTol
       << ", max iterations " << p.maxNumIters << ", robust cost " << (int)p.robustCostParams.costType << "\n";
    return os;
  }
};

#define DPGO_SHIM_LOCK std::lock_guard<std::recursive_mutex> dpgo_shim_lock_(mMutex)

class PGOAgent {
 public:
  PGOAgent(unsigned ID, const PGOAgentParameters &params)
      : mID(ID), d(params.d), r(params.r), mParams(params), mStatus(ID, PGOAgentState::WAIT_FOR_DATA, 0, 0, false, 0),
        mRobustCost(params.robustCostParams), mTeamRobotActive(params.numRobots, true) {
    mPoseGraph = std::make_shared<PoseGraph>(mID, r, d);
    createHandle();
    // Robot 0 owns the lifting matrix of the team: the wrapper only ever READS it from the leader (getLiftingMatrix,
    // src/PGOAgentROS.cpp:404) and broadcasts it to the others (liftingMatrixCallback -> setLiftingMatrix, :928), so
    // the base class has to create it.  Upstream draws a random point of St(d, r); a fixed one keeps runs reproducible.
    if (mID == 0) setLiftingMatrix(fixedStiefelVariable(d, r));
  }
  virtual ~PGOAgent() {
    endOptimizationLoop();
    if (h_) dpgo_b200_agent_destroy(h_);
  }

  // ---- asynchronous mode: the library-owned optimisation thread
  void startOptimizationLoop(double freq) {
    if (mOptimizationThread) return;
    mEndLoopRequested = false;
    mOptimizationThread.reset(new std::thread([this, freq] {
      std::mt19937 rng(mID + 1);
      std::exponential_distribution<double> gap(freq > 0 ? freq : 1.0);
      while (!mEndLoopRequested) {
        std::this_thread::sleep_for(std::chrono::duration<double>(gap(rng)));
        if (mEndLoopRequested) break;
        DPGO_SHIM_LOCK;
        if (mState != PGOAgentState::INITIALIZED) continue;
        iterate(true);
        mPublishAsynchronousRequested = true;   // polled at src/PGOAgentROS.cpp:120
      }
    }));
  }
  void endOptimizationLoop() {
    if (!mOptimizationThread) return;
    mEndLoopRequested = true;
    mOptimizationThread->join();
    mOptimizationThread.reset();
  }
  bool isOptimizationRunning() const { return (bool)mOptimizationThread; }
  PGOAgent(const PGOAgent &) = delete;
  PGOAgent &operator=(const PGOAgent &) = delete;

  // ---- identity / counters
  unsigned getID() const { return mID; }
  unsigned dimension() const { return d; }
  unsigned relaxation_rank() const { return r; }
  unsigned num_poses() const { return mPoseGraph->n(); }                       // :285
  unsigned instance_number() const { return mInstanceNumber; }                 // :433
  unsigned iteration_number() const { return mIterationNumber; }               // :139
  std::vector<unsigned> getNeighbors() const {                                 // :663
    const auto &s = mPoseGraph->neighborIDs();
    return std::vector<unsigned>(s.begin(), s.end());
  }

  // ---- pose graph
  void addMeasurement(const RelativeSEMeasurement &m) {                        // :277, :1307
    DPGO_SHIM_LOCK;
    rebindGraphIfReplaced();
    if (mState != PGOAgentState::WAIT_FOR_DATA) return;  // measurements are fixed once a round has started
    if (!mPoseGraph->addMeasurement(m)) return;
    uploadMeasurement(m);
  }

  // ---- lifecycle
  void setLiftingMatrix(const Matrix &M) {                                     // :928
    DPGO_SHIM_LOCK;
    if (M.rows() != r || M.cols() != d) throw std::invalid_argument("setLiftingMatrix: expected r x d");
    YLift.emplace(M);
    check(dpgo_b200_set_lifting_matrix(h_, M.data()), "setLiftingMatrix");
  }
  bool getLiftingMatrix(Matrix &M) const {                                     // :404
    if (!YLift.has_value()) return false;
    M = YLift.value();
    return true;
  }
  // local initialisation: odometry chain, or the trajectory estimate handed over by the front end (:285-303)
  void initialize(const PoseArray *TInitPtr = nullptr) {                       // :348
    DPGO_SHIM_LOCK;
    rebindGraphIfReplaced();
    if (mPoseGraph->n() == 0) return;
    if (TInitPtr && TInitPtr->n() == num_poses()) {
      std::vector<double> T((size_t)12 * num_poses());
      for (unsigned i = 0; i < num_poses(); ++i) {
        const Matrix P = TInitPtr->pose(i);
        for (int a = 0; a < 3; ++a)
          for (int c = 0; c < 4; ++c) T[(size_t)i * 12 + a * 4 + c] = P(a, c);
      }
      check(dpgo_b200_initialize(h_, T.data()), "initialize");
    } else if (mParams.localInitializationMethod == InitializationMethod::Chordal) {
      check(dpgo_b200_initialize_chordal(h_), "initialize (Chordal)");               // PGOAgentROSNode.cpp:108-109
    } else if (mParams.localInitializationMethod == InitializationMethod::GNC_TLS) {
      std::vector<double> T;                                                            // PGOAgentROSNode.cpp:110-111
      if (robustLocalInitialization(T))
        check(dpgo_b200_initialize(h_, T.data()), "initialize (GNC_TLS)");
      else
        check(dpgo_b200_initialize_chordal(h_), "initialize (GNC_TLS -> Chordal)");
    } else {
      check(dpgo_b200_initialize(h_, nullptr), "initialize");                          // Odometry
    }
    mState = PGOAgentState::WAIT_FOR_INITIALIZATION;
    mStatus.state = mState;
  }
  void initializeInGlobalFrame(const Pose &T_world_robot) {                    // :353, :358
    DPGO_SHIM_LOCK;
    if (mState == PGOAgentState::WAIT_FOR_DATA) return;
    if (!YLift.has_value()) throw std::runtime_error("initializeInGlobalFrame: lifting matrix not set");
    double T[12];
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 4; ++c) T[a * 4 + c] = T_world_robot.getData()(a, c);
    check(dpgo_b200_initialize_in_global_frame(h_, T), "initializeInGlobalFrame");
    mState = PGOAgentState::INITIALIZED;
    mStatus.state = mState;
    if (mID == 0) anchorFirstPose();
    if (mParams.asynchronous) startOptimizationLoop(mParams.asynchronousOptimizationRate);   // (upstream starts it here too)
  }
  // use this robot's first pose as the global anchor (:360)
  void anchorFirstPose() {
    Matrix X0;
    if (getSharedPose(0, X0)) setGlobalAnchor(X0);
  }
  void setGlobalAnchor(const Matrix &M) {                                      // :939, :1466
    DPGO_SHIM_LOCK;
    if (M.rows() != r || M.cols() != d + 1) throw std::invalid_argument("setGlobalAnchor: expected r x (d+1)");
    globalAnchor.emplace(LiftedPose(M));
  }
  virtual void reset() {                                                       // :223
    endOptimizationLoop();
    DPGO_SHIM_LOCK;
    check(dpgo_b200_reset(h_), "reset");
    mInstanceNumber++;
    mIterationNumber = mLastIterationSeen = 0;
    mWeightUpdateCount = 0;
    mRobustOptInnerIter = 0;
    mState = PGOAgentState::WAIT_FOR_DATA;
    mStatus = PGOAgentStatus(mID, mState, mInstanceNumber, 0, false, 0);
    mTeamStatus.clear();
    mLocalOptResult = ROPTResult();
    mPublishPublicPosesRequested = false;
    mPublishAsynchronousRequested = false;
    globalAnchor.reset();
    neighborPoseDict.clear();
    neighborAuxPoseDict.clear();
    for (unsigned id = 0; id < mTeamRobotActive.size(); ++id) setRobotActive(id, true);
    // The lifting matrix and the MEASUREMENTS survive a reset: the wrapper's next round only adds what
    // hasMeasurement() does not know yet (src/PGOAgentROS.cpp:268-280) and relies on the weights it fixed in the
    // TERMINATE handler (m->weight = 0, m->fixedWeight = true, :1051-1054) still being there; it swaps in a fresh
    // PoseGraph itself when it wants a clean slate (completeReset, :237).  The device agent is rebuilt from the mirror.
    recreateHandle();
  }

  // ---- the hot call
  bool iterate(bool doOptimization = true) {                                   // :160 (true), :1185 (false)
    DPGO_SHIM_LOCK;
    if (mIterationNumber != mLastIterationSeen)   // the wrapper rewound mIterationNumber (RECOVER, :1196)
      dpgo_b200_set_iteration_number(h_, (int)mIterationNumber);
    const int rc = dpgo_b200_iterate(h_, doOptimization ? 1 : 0);
    // DPGO_B200_ERR_MISSING: the iteration was counted but the local solve had to be skipped (a neighbour's poses have
    // not arrived); upstream returns the success of the solve and the wrapper logs it (:160-175)
    if (rc != 0 && rc != DPGO_B200_ERR_MISSING) return false;
    const bool solved = rc == 0;
    dpgo_b200_status s;
    dpgo_b200_get_status(h_, &s);
    mIterationNumber = mLastIterationSeen = (unsigned)s.iteration_number;
    mStatus.iterationNumber = mIterationNumber;
    mStatus.instanceNumber = mInstanceNumber;
    mStatus.state = mState;
    mStatus.relativeChange = s.relative_change;                                // :891
    mStatus.readyToTerminate = s.ready_to_terminate != 0;
    mTeamStatus[mID] = mStatus;
    if (mParams.robustCostParams.costType != RobustCostParameters::Type::L2) mRobustOptInnerIter++;
    if (mState == PGOAgentState::INITIALIZED) {
      if (doOptimization && solved) {
        dpgo_b200_opt_result o;
        // fOpt / gradNormOpt cost a second gradient pass on the GPU: only fetched when somebody prints them (:169-172)
        (mParams.verbose ? dpgo_b200_get_opt_result : dpgo_b200_get_opt_result_lazy)(h_, &o);
        mLocalOptResult.success = o.success != 0;
        mLocalOptResult.fInit = o.f_init;
        mLocalOptResult.fOpt = o.f_opt;
        mLocalOptResult.gradNormInit = o.gradnorm_init;
        mLocalOptResult.gradNormOpt = o.gradnorm_opt;
        mLocalOptResult.relativeChange = o.relative_change;
      }
      mPublishPublicPosesRequested = mParams.acceleration || doOptimization;   // :109
    }
    return solved;
  }

  // ---- public poses (a9)
  bool getSharedPose(unsigned index, Matrix &Mout) {                           // :424
    DPGO_SHIM_LOCK;
    if (mState != PGOAgentState::INITIALIZED || index >= num_poses()) return false;
    Mout = Matrix(r, d + 1);
    return dpgo_b200_get_pose(h_, 0, (int)index, Mout.data()) == 0;
  }
  bool getSharedPoseDictWithNeighbor(PoseDict &map, unsigned neighborID) { return getDict(map, neighborID, 0); }      // :668
  bool getAuxSharedPoseDictWithNeighbor(PoseDict &map, unsigned neighborID) { return getDict(map, neighborID, 1); }   // :666
  void updateNeighborPoses(unsigned neighborID, const PoseDict &poseDict) { putDict(neighborID, poseDict, 0); }       // :1276
  void updateAuxNeighborPoses(unsigned neighborID, const PoseDict &poseDict) { putDict(neighborID, poseDict, 1); }    // :1278
  void setNeighborPoses(unsigned neighborID, const PoseDict &poseDict) { updateNeighborPoses(neighborID, poseDict); } // north-star alias
  Matrix getX() {                                                              // north-star alias: r x (d+1) n
    DPGO_SHIM_LOCK;
    Matrix X(r, (size_t)(d + 1) * num_poses());
    check(dpgo_b200_get_x(h_, 0, X.data()), "getX");
    return X;
  }

  // ---- global-frame read-out (a11): T_i = (proj_SO(3)(Ya^T Yi), Ya^T (pi - pa)) with the anchor [Ya | pa]
  bool getTrajectoryInGlobalFrame(PoseArray &Trajectory) {                     // :624, :657
    DPGO_SHIM_LOCK;
    if (!globalAnchor.has_value() || mState != PGOAgentState::INITIALIZED) return false;
    const Matrix X = getX();
    PoseArray T(d, num_poses());
    for (unsigned i = 0; i < num_poses(); ++i) T.pose(i) = roundPose(X.block(0, (size_t)i * (d + 1), r, d + 1));
    Trajectory = T;
    return true;
  }
  bool getPoseInGlobalFrame(unsigned poseID, Matrix &T) {                      // :774-775, :807, :811
    DPGO_SHIM_LOCK;
    Matrix Xi;
    if (!globalAnchor.has_value() || !getSharedPose(poseID, Xi)) return false;
    T = roundPose(Xi);
    return true;
  }
  bool getNeighborPoseInGlobalFrame(unsigned neighborID, unsigned poseID, Matrix &T) {   // :808, :812, :1395
    DPGO_SHIM_LOCK;
    if (!globalAnchor.has_value()) return false;
    auto it = neighborPoseDict.find(PoseID(neighborID, poseID));
    if (it == neighborPoseDict.end()) return false;
    T = roundPose(it->second.getData());
    return true;
  }

  // ---- status / termination (a10)
  PGOAgentStatus getStatus() {                                                 // :616
    DPGO_SHIM_LOCK;
    mStatus.agentID = mID;
    mStatus.state = mState;
    mStatus.instanceNumber = mInstanceNumber;
    mStatus.iterationNumber = mIterationNumber;
    return mStatus;
  }
  void setNeighborStatus(const PGOAgentStatus &status) {                       // :965
    DPGO_SHIM_LOCK;
    mTeamStatus[status.agentID] = status;
    dpgo_b200_status s{(int)status.agentID, (int)status.state, (int)status.instanceNumber, (int)status.iterationNumber,
                       status.readyToTerminate ? 1 : 0, status.relativeChange};
    dpgo_b200_set_neighbor_status(h_, &s);
  }
  bool hasNeighborStatus(unsigned id) const {                                              // :1116
    DPGO_SHIM_LOCK;
    return mTeamStatus.count(id) != 0;
  }
  PGOAgentStatus getNeighborStatus(unsigned id) const {                                    // :1121
    DPGO_SHIM_LOCK;
    return mTeamStatus.at(id);
  }
  void setRobotActive(unsigned id, bool active) {                                          // :382 ... :1582
    DPGO_SHIM_LOCK;
    if (id < mTeamRobotActive.size()) mTeamRobotActive[id] = active;
    mPoseGraph->setNeighborActive(id, active);
    if (h_ && id < mParams.numRobots) dpgo_b200_set_robot_active(h_, (int)id, active ? 1 : 0);
  }
  bool isRobotActive(unsigned id) const { return id < mTeamRobotActive.size() && mTeamRobotActive[id]; }   // :195
  bool isRobotInitialized(unsigned id) const {                                             // :451, :468, :1144
    DPGO_SHIM_LOCK;
    if (id == mID) return mState == PGOAgentState::INITIALIZED;
    auto it = mTeamStatus.find(id);
    return it != mTeamStatus.end() && it->second.state == PGOAgentState::INITIALIZED;
  }
  size_t numActiveRobots() const {                                                       // :554, :885
    size_t k = 0;
    for (bool b : mTeamRobotActive) k += b;
    return k;
  }
  bool shouldTerminate() {                                                                 // :208
    DPGO_SHIM_LOCK;
    mStatus.iterationNumber = mIterationNumber;
    return dpgo_b200_should_terminate(h_) == 1;
  }

  // ---- GNC (a8)
  bool shouldUpdateMeasurementWeights() {                                                  // :210
    DPGO_SHIM_LOCK;
    return dpgo_b200_should_update_measurement_weights(h_) == 1;
  }
  void updateMeasurementWeights() {                                                        // :1218
    DPGO_SHIM_LOCK;
    check(dpgo_b200_update_measurement_weights(h_), "updateMeasurementWeights");
    pullWeights();
    mWeightUpdateCount = (unsigned)dpgo_b200_weight_update_count(h_);
    mRobustOptInnerIter = 0;
  }
  bool setMeasurementWeight(const PoseID &src_ID, const PoseID &dst_ID, double weight, bool fixed_weight = false) {  // :1341
    DPGO_SHIM_LOCK;
    RelativeSEMeasurement *m = mPoseGraph->findMeasurement(src_ID, dst_ID);
    if (!m) return false;
    m->weight = weight;
    m->fixedWeight = fixed_weight;
    return dpgo_b200_set_measurement_weight(h_, (int)src_ID.robot_id, (int)src_ID.frame_id, (int)dst_ID.robot_id,
                                            (int)dst_ID.frame_id, weight, fixed_weight ? 1 : 0) == 0;
  }
  bool computeMeasurementResidual(const RelativeSEMeasurement &measurement, double *residual) {   // :1049
    DPGO_SHIM_LOCK;
    return dpgo_b200_compute_measurement_residual(h_, (int)measurement.r1, (int)measurement.p1, (int)measurement.r2,
                                                  (int)measurement.p2, residual) == 0;
  }

 protected:
  // ---- members PGOAgentROS reads / writes directly ("Protected data members")
  unsigned mID;
  unsigned d, r;
  PGOAgentParameters mParams;
  PGOAgentState mState = PGOAgentState::WAIT_FOR_DATA;
  PGOAgentStatus mStatus;
  std::shared_ptr<PoseGraph> mPoseGraph;            // reassigned by the wrapper at :237
  ROPTResult mLocalOptResult;                       // :169-172
  unsigned mInstanceNumber = 0;
  // written by the RECOVER handler (:1196) and, in asynchronous mode, by the optimisation thread inside iterate() while
  // the wrapper reads it through iteration_number(): atomic, assignable and readable like the plain unsigned upstream has
  std::atomic<unsigned> mIterationNumber{0};
  unsigned mWeightUpdateCount = 0;                  // :193
  unsigned mRobustOptInnerIter = 0;                 // :193, :545
  std::map<unsigned, PGOAgentStatus> mTeamStatus;   // :196-199
  RobustCost mRobustCost;                           // :1050
  std::atomic<bool> mPublishPublicPosesRequested{false};    // :109, :112 (raised inside iterate(), possibly by the optimisation thread)
  std::atomic<bool> mPublishAsynchronousRequested{false};   // :120, :125 (raised by the optimisation thread)
  std::optional<Matrix> YLift;                      // :1408, :1419, :1459
  std::optional<LiftedPose> globalAnchor;           // :426-429
  PoseDict neighborPoseDict, neighborAuxPoseDict;   // :1422
  std::vector<bool> mTeamRobotActive;

 private:
  static void check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string("DPGO::PGOAgent::") + what + ": " + dpgo_b200_last_error());
  }
  dpgo_b200_params toC() const {
    dpgo_b200_params q{};
    const PGOAgentParameters &p = mParams;
    q.d = (int)p.d;
    q.r = (int)p.r;
    q.num_robots = (int)p.numRobots;
    q.method = p.localOptimizationParams.method == ROptParameters::ROptMethod::RTR ? 0 : 1;
    q.rgd_stepsize = p.localOptimizationParams.RGD_stepsize;
    q.rgd_use_preconditioner = p.localOptimizationParams.RGD_use_preconditioner;
    q.rtr_iterations = (int)p.localOptimizationParams.RTR_iterations;
    q.rtr_tcg_iterations = (int)p.localOptimizationParams.RTR_tCG_iterations;
    q.rtr_initial_radius = p.localOptimizationParams.RTR_initial_radius;
    q.gradnorm_tol = p.localOptimizationParams.gradnorm_tol;
    q.acceleration = p.acceleration;
    q.restart_interval = (int)p.restartInterval;
    q.cost_type = (int)p.robustCostParams.costType;
    q.gnc_barc = p.robustCostParams.GNCBarc;
    q.gnc_mu_step = p.robustCostParams.GNCMuStep;
    q.gnc_init_mu = p.robustCostParams.GNCInitMu;
    q.robust_opt_num_weight_updates = (int)p.robustOptNumWeightUpdates;
    q.robust_opt_num_resets = (int)p.robustOptNumResets;
    q.robust_opt_inner_iters = (int)p.robustOptInnerIters;
    q.robust_opt_min_convergence_ratio = p.robustOptMinConvergenceRatio;
    q.max_num_iters = (int)p.maxNumIters;
    q.rel_change_tol = p.relChangeTol;
    q.precond_lambda = p.preconditionerShift;
    return q;
  }
  void createHandle() {
    const dpgo_b200_params q = toC();
    check(dpgo_b200_agent_create((int)mID, &q, mParams.device, &h_), "PGOAgent");
    mBoundGraph = mPoseGraph.get();
    dpgo_b200_agent_t h = h_;
    // both are reached from the wrapper WITHOUT passing through a PGOAgent method (mPoseGraph->clearDataMatrices(),
    // :1351; mRobustCost.weight(), :1050) -- possibly while the optimisation thread is inside iterate()
    mPoseGraph->bindClear([this, h] {
      std::lock_guard<std::recursive_mutex> lock(mMutex);
      dpgo_b200_clear_data_matrices(h);
    });
    mRobustCost.bind([this, h](double res) {
      std::lock_guard<std::recursive_mutex> lock(mMutex);
      return dpgo_b200_robust_weight(h, res);
    });
  }
  void recreateHandle() {
    if (h_) dpgo_b200_agent_destroy(h_);
    h_ = nullptr;
    createHandle();
    if (YLift.has_value()) dpgo_b200_set_lifting_matrix(h_, YLift.value().data());
    for (unsigned id = 0; id < mTeamRobotActive.size() && id < mParams.numRobots; ++id)
      if (!mTeamRobotActive[id]) dpgo_b200_set_robot_active(h_, (int)id, 0);
    // measurements already in the host mirror (kept across reset(), or pre-loaded into a graph the wrapper swapped in)
    for (auto *vec : {&mPoseGraph->odometry(), &mPoseGraph->privateLoopClosures(), &mPoseGraph->sharedLoopClosures()})
      for (const auto &m : *vec) uploadMeasurement(m);
  }
  void uploadMeasurement(const RelativeSEMeasurement &m) {
    const int r1 = (int)m.r1, p1 = (int)m.p1, r2 = (int)m.r2, p2 = (int)m.p2;
    double Rrm[9], tv[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Rrm[i * 3 + j] = m.R(i, j);
      tv[i] = m.t(i, 0);
    }
    const unsigned char fixed = m.fixedWeight ? 1 : 0;
    check(dpgo_b200_add_measurements(h_, 1, &r1, &p1, &r2, &p2, Rrm, tv, &m.kappa, &m.tau, &m.weight, &fixed),
          "addMeasurement");
  }
  // the wrapper swaps in a fresh PoseGraph on a complete reset (:237): follow it with a fresh device agent
  void rebindGraphIfReplaced() {
    if (mPoseGraph.get() == mBoundGraph) return;
    mState = PGOAgentState::WAIT_FOR_DATA;
    recreateHandle();
  }
  void pullWeights() {
    const unsigned np = mPoseGraph->numPrivateLoopClosures(), ns = mPoseGraph->numSharedLoopClosures();
    std::vector<double> w(np + ns);
    const int k = dpgo_b200_get_lc_weights(h_, w.data(), (int)w.size());
    if (k != (int)(np + ns)) return;
    for (unsigned e = 0; e < np; ++e) mPoseGraph->privateLoopClosures()[e].weight = w[e];
    for (unsigned e = 0; e < ns; ++e) mPoseGraph->sharedLoopClosures()[e].weight = w[np + e];
  }
  bool getDict(PoseDict &map, unsigned nbr, int aux) {
    DPGO_SHIM_LOCK;
    if (mState != PGOAgentState::INITIALIZED) return false;
    const int cap = dpgo_b200_num_shared_poses(h_, (int)nbr);
    if (cap < 0) return false;
    map.clear();
    if (cap == 0) return true;
    std::vector<int> ids(cap);
    std::vector<double> buf((size_t)cap * r * 4);  // each pose r x 4, column-major
    int n = 0;
    if (dpgo_b200_get_shared_pose_dict(h_, (int)nbr, aux, ids.data(), buf.data(), cap, &n) != 0) return false;
    for (int k = 0; k < n; ++k) {
      Matrix M(r, 4);
      std::copy(buf.begin() + (size_t)k * r * 4, buf.begin() + (size_t)(k + 1) * r * 4, M.data());
      map.emplace(PoseID(mID, (unsigned)ids[k]), LiftedPose(M));
    }
    return true;
  }
  void putDict(unsigned nbr, const PoseDict &dict, int aux) {
    DPGO_SHIM_LOCK;
    std::vector<int> ids;
    std::vector<double> buf;
    ids.reserve(dict.size());
    buf.reserve(dict.size() * r * 4);
    PoseDict &mirror = aux ? neighborAuxPoseDict : neighborPoseDict;
    for (const auto &kv : dict) {
      if (kv.first.robot_id != nbr) continue;
      ids.push_back((int)kv.first.frame_id);
      const Matrix &M = kv.second.getData();
      buf.insert(buf.end(), M.data(), M.data() + (size_t)r * 4);
      mirror[kv.first] = kv.second;
    }
    if (!ids.empty())
      check(dpgo_b200_update_neighbor_poses(h_, (int)nbr, aux, ids.data(), buf.data(), (int)ids.size()),
            "updateNeighborPoses");
    // cross-robot initialisation: a robot that only has its local trajectory places itself in the global frame from
    // the first initialised neighbour whose public poses it hears (INITIALIZE round, src/PGOAgentROS.cpp:1091-1159)
    if (!aux && mState == PGOAgentState::WAIT_FOR_INITIALIZATION && mParams.multirobotInitialization)
      tryInitializeFromNeighbor(nbr);
  }
  // local_initialization_method "GNC_TLS": robust pose graph optimisation over the robot's OWN odometry and private loop
  // closures before it meets the team -- graduated non-convexity with the truncated-least-squares cost (Yang et al., RA-L
  // 2020) around the same RTR solver: solve, re-weight every loop closure from its residual, tighten mu, until every
  // weight has settled at 0 or 1.  Composed from the library's own entry points on a scratch single-robot agent
  // (odometry keeps weight 1: fixedWeight, src/utils.cpp:147-149).  Returns the local trajectory (pose 0 = identity),
  // n x 3 x 4 row-major.  [UPSTREAM-RECALL of the method; the constants below are this repo's.]
  bool robustLocalInitialization(std::vector<double> &T) {
    const unsigned n = num_poses();
    if (n == 0 || mPoseGraph->numPrivateLoopClosures() == 0) return false;   // nothing to reject: Chordal is the same guess
    dpgo_b200_params q = toC();
    q.method = 0;                    // RTR
    q.rtr_iterations = 20;
    q.rtr_tcg_iterations = 100;
    q.gradnorm_tol = 1.0;
    q.acceleration = 0;
    q.cost_type = (int)RobustCostParameters::Type::GNC_TLS;
    q.robust_opt_num_weight_updates = 1 << 30;
    q.robust_opt_num_resets = 0;
    q.robust_opt_inner_iters = 1;
    q.max_num_iters = 1 << 30;
    q.rel_change_tol = 0;
    dpgo_b200_agent_t tmp = nullptr;
    if (dpgo_b200_agent_create((int)mID, &q, mParams.device, &tmp) != 0) return false;
    bool ok = true;
    auto upload = [&](const RelativeSEMeasurement &m) {
      const int r1 = (int)m.r1, p1 = (int)m.p1, r2 = (int)m.r2, p2 = (int)m.p2;
      double Rrm[9], tv[3];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rrm[i * 3 + j] = m.R(i, j);
        tv[i] = m.t(i, 0);
      }
      const unsigned char fixed = m.fixedWeight ? 1 : 0;
      ok = ok && dpgo_b200_add_measurements(tmp, 1, &r1, &p1, &r2, &p2, Rrm, tv, &m.kappa, &m.tau, &m.weight, &fixed) == 0;
    };
    for (const auto &m : mPoseGraph->odometry()) upload(m);
    for (const auto &m : mPoseGraph->privateLoopClosures()) upload(m);
    const Matrix Y0 = fixedStiefelVariable(d, r);   // any lifting matrix serves a single-robot problem
    const double eye[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    ok = ok && dpgo_b200_set_lifting_matrix(tmp, Y0.data()) == 0 && dpgo_b200_initialize_chordal(tmp) == 0 &&
         dpgo_b200_initialize_in_global_frame(tmp, eye) == 0;
    const unsigned np = mPoseGraph->numPrivateLoopClosures();
    std::vector<double> w(np, 1.0);
    // For small mu every loop closure is weak (w ~ barc sqrt(mu) / residual): the surrogate is nearly convex and close
    // to the odometry solution.  mu grows by GNCMuStep per round; the weights only mean "inlier / outlier" once the
    // transition band [mu / (mu + 1), (mu + 1) / mu] barc^2 has closed in on barc^2, so never stop before mu >= 1.
    double mu = q.gnc_init_mu;
    for (unsigned it = 0; ok && it < 100; ++it) {
      ok = dpgo_b200_iterate(tmp, 1) == 0 && dpgo_b200_update_measurement_weights(tmp) == 0 &&
           dpgo_b200_get_lc_weights(tmp, w.data(), (int)np) == (int)np;
      mu *= q.gnc_mu_step;
      bool settled = mu >= 1.0;
      for (double v : w) settled = settled && (v < 1e-3 || v > 1.0 - 1e-3);
      if (settled) break;
    }
    if (ok) {
      ok = dpgo_b200_iterate(tmp, 1) == 0;          // one more solve under the final weights
      Matrix X(r, (size_t)(d + 1) * n);
      ok = ok && dpgo_b200_get_x(tmp, 0, X.data()) == 0;
      if (ok) {
        const Matrix Ya = X.block(0, 0, r, d), pa = X.block(0, d, r, 1), YaT = Ya.transpose();
        T.assign((size_t)12 * n, 0.0);
        for (unsigned i = 0; i < n; ++i) {
          const Matrix Ri = projectToRotationGroup(YaT * X.block(0, (size_t)i * (d + 1), r, d));
          const Matrix ti = YaT * (X.block(0, (size_t)i * (d + 1) + d, r, 1) - pa);
          for (int a = 0; a < 3; ++a) {
            for (int c = 0; c < 3; ++c) T[(size_t)i * 12 + a * 4 + c] = Ri(a, c);
            T[(size_t)i * 12 + a * 4 + 3] = ti(a, 0);
          }
        }
      }
    }
    dpgo_b200_agent_destroy(tmp);
    return ok;
  }
  void tryInitializeFromNeighbor(unsigned nbr) {
    if (!YLift.has_value() || num_poses() == 0) return;
    std::vector<double> Tl((size_t)12 * num_poses());
    if (dpgo_b200_get_local_trajectory(h_, Tl.data()) != 0) return;
    auto localPose = [&](size_t i) {
      Matrix T(3, 4);
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 4; ++c) T(a, c) = Tl[i * 12 + a * 4 + c];
      return T;
    };
    const Matrix YT = YLift.value().transpose();
    std::vector<Matrix> candidates;
    for (const auto &m : mPoseGraph->sharedLoopClosures()) {
      Matrix meas(3, 4);
      meas.block(0, 0, 3, 3) = m.R;
      meas.block(0, 3, 3, 1) = m.t;
      if (m.r1 == nbr && m.r2 == mID) {          // neighbour pose -> my pose
        auto it = neighborPoseDict.find(PoseID(nbr, (unsigned)m.p1));
        if (it == neighborPoseDict.end()) continue;
        Matrix Tw = YT * it->second.getData();    // unlift: X = YLift T
        Tw.block(0, 0, 3, 3) = projectToRotationGroup(Tw.block(0, 0, 3, 3));
        candidates.push_back(se3Compose(se3Compose(Tw, meas), se3Inverse(localPose(m.p2))));
      } else if (m.r2 == nbr && m.r1 == mID) {   // my pose -> neighbour pose
        auto it = neighborPoseDict.find(PoseID(nbr, (unsigned)m.p2));
        if (it == neighborPoseDict.end()) continue;
        Matrix Tw = YT * it->second.getData();
        Tw.block(0, 0, 3, 3) = projectToRotationGroup(Tw.block(0, 0, 3, 3));
        candidates.push_back(se3Compose(se3Compose(Tw, se3Inverse(meas)), se3Inverse(localPose(m.p1))));
      }
    }
    Matrix Tworld;
    if (!robustTransformAverage(candidates, 0.2, 1.0, mParams.robustInitMinInliers, Tworld)) return;
    initializeInGlobalFrame(Pose(Tworld));
  }
  Matrix roundPose(const Matrix &Xi) const {
    const Matrix Ya = globalAnchor.value().rotation(), pa = globalAnchor.value().translation();
    const Matrix YaT = Ya.transpose();
    Matrix T(d, d + 1);
    T.block(0, 0, d, d) = projectToRotationGroup(YaT * Xi.block(0, 0, r, d));
    T.block(0, d, d, 1) = YaT * (Xi.block(0, d, r, 1) - pa);
    return T;
  }

  dpgo_b200_agent_t h_ = nullptr;
  const PoseGraph *mBoundGraph = nullptr;
  unsigned mLastIterationSeen = 0;
  mutable std::recursive_mutex mMutex;
  std::unique_ptr<std::thread> mOptimizationThread;
  std::atomic<bool> mEndLoopRequested{false};
};

#undef DPGO_SHIM_LOCK

}  // namespace DPGO
#endif
