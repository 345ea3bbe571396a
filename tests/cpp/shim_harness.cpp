// ROS-free stand-in for PGOAgentROS on top of the DPGO:: shim (include/DPGO/*.h -> libdpgo_b200.so).
//
// It does what the wrapper does with its base class and nothing else: derives from DPGO::PGOAgent
// (include/dpgo_ros/PGOAgentROS.h:121), touches the same protected members, splits a g2o file
// with the dataset publisher's rule (src/PGODatasetPublisherNode.cpp:84-134) and replays the synchronous
// protocol (src/PGOAgentROS.cpp:102-220, 1161-1189, 662-690, 1255-1284) for N robots in one process.
// Prints one line per fact the Python test checks and dumps every robot's X as raw doubles.
//
//   shim_harness <file.g2o> <num_robots> <max_iters> <rgd|rtr> <out_prefix> [device]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "DPGO/DPGO_solver.h"
#include "DPGO/DPGO_utils.h"
#include "DPGO/PGOAgent.h"

using namespace DPGO;

class WrapperAgent : public PGOAgent {  // plays PGOAgentROS
 public:
  WrapperAgent(unsigned ID, const PGOAgentParameters &p) : PGOAgent(ID, p) {}
  // members the wrapper reads directly
  bool publishRequested() const { return mPublishPublicPosesRequested; }       // :109
  void clearPublishRequest() { mPublishPublicPosesRequested = false; }          // :112
  double relativeChange() const { return mStatus.relativeChange; }              // :891
  const ROPTResult &localOptResult() const { return mLocalOptResult; }          // :169-172
  PGOAgentState state() const { return mState; }                                // :418
  unsigned numSharedLCs() const { return mPoseGraph->numSharedLoopClosures(); } // :345
  unsigned numOdom() const { return mPoseGraph->numOdometry(); }
  unsigned numPrivateLCs() const { return mPoseGraph->numPrivateLoopClosures(); }
  bool accelerated() const { return mParams.acceleration; }
  bool publishAsyncRequested() const { return mPublishAsynchronousRequested; }  // :120
  void clearPublishAsyncRequest() { mPublishAsynchronousRequested = false; }    // :125
  std::set<unsigned> activeNeighbors() const { return mPoseGraph->activeNeighborIDs(); }   // :137
};

int main(int argc, char **argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: shim_harness <file.g2o> <num_robots> <max_iters> <rgd|rtr> <out_prefix> [device]\n");
    return 2;
  }
  const std::string file = argv[1], prefix = argv[5];
  std::string mode = argv[4];
  // "<mode>_multi": no global-frame guess is handed over -- every robot initialises in its own frame and robots
  // other than 0 place themselves in the global frame from their neighbours' public poses (multirobot_initialization)
  const bool multi_init = mode.size() > 6 && mode.substr(mode.size() - 6) == "_multi";
  if (multi_init) mode = mode.substr(0, mode.size() - 6);
  const unsigned N = (unsigned)std::atoi(argv[2]);
  const int max_iters = std::atoi(argv[3]);
  const int device = argc > 6 ? std::atoi(argv[6]) : 0;
  const unsigned d = 3, r = 5;
  try {
    size_t num_poses = 0;
    std::vector<RelativeSEMeasurement> dataset = read_g2o_file(file, num_poses);       // PGODatasetPublisherNode.cpp:80
    // contiguous split, the last robot takes the remainder (:84-103)
    const size_t per = num_poses / N;
    auto robot_of = [&](size_t g) { return std::min<size_t>(g / per, N - 1); };
    auto local_of = [&](size_t g) { return g - robot_of(g) * per; };

    PGOAgentParameters params(d, r, N);
    params.device = device;
    params.relChangeTol = 0.1;
    params.maxNumIters = 100000;
    if (multi_init) params.robustInitMinInliers = 1;  // (the Python harness' alignment has no inlier floor either)
    if (mode == "rgd") {
      params.localOptimizationParams.method = ROptParameters::ROptMethod::RGD;
      params.localOptimizationParams.RGD_stepsize = 0.2;
      params.localOptimizationParams.RGD_use_preconditioner = true;
      params.acceleration = true;
      params.restartInterval = 50;
    } else if (mode == "async") {  // launch/asapp_demo.launch:7-9: asynchronous, RGD 0.2 + preconditioner, no acceleration
      params.asynchronous = true;
      params.asynchronousOptimizationRate = 2000;
      params.localOptimizationParams.method = ROptParameters::ROptMethod::RGD;
      params.localOptimizationParams.RGD_stepsize = 0.2;
      params.localOptimizationParams.RGD_use_preconditioner = true;
      params.acceleration = false;
    } else {
      params.localOptimizationParams.method = ROptParameters::ROptMethod::RTR;
      params.localOptimizationParams.gradnorm_tol = 0.5;
      params.acceleration = false;
    }
    std::vector<std::unique_ptr<WrapperAgent>> agents;
    for (unsigned a = 0; a < N; ++a) agents.emplace_back(new WrapperAgent(a, params));

    // requestPoseGraph (:246-320): odometry, private and shared loop closures of every robot; shared loop closures
    // reach both ends (publishPublicMeasurements, :692-719 / :1286-1313)
    std::vector<Matrix> Rg(num_poses, Matrix::Identity(3, 3)), tg(num_poses, Matrix(3, 1));
    std::vector<const RelativeSEMeasurement *> odo_global(num_poses, nullptr);
    for (const auto &mg : dataset) {
      RelativeSEMeasurement m = mg;
      m.r1 = robot_of(mg.p1);
      m.r2 = robot_of(mg.p2);
      m.p1 = local_of(mg.p1);
      m.p2 = local_of(mg.p2);
      if (mg.p1 + 1 == mg.p2) odo_global[mg.p1] = &mg;
      if (m.r1 == m.r2 && m.p1 + 1 == m.p2) m.fixedWeight = true;                      // src/utils.cpp:147-149
      agents[m.r1]->addMeasurement(m);
      if (m.r2 != m.r1) agents[m.r2]->addMeasurement(m);
    }
    // trajectory estimate handed over with the pose graph (:285-303): the global odometry chain
    for (size_t g = 0; g + 1 < num_poses; ++g) {
      if (!odo_global[g]) throw std::runtime_error("dataset without an odometry chain");
      Rg[g + 1] = Rg[g] * odo_global[g]->R;
      tg[g + 1] = Rg[g] * odo_global[g]->t + tg[g];
    }
    const Matrix YLift = fixedStiefelVariable(d, r);
    for (unsigned a = 0; a < N; ++a) {
      const unsigned n = agents[a]->num_poses();
      PoseArray TInit(d, n);
      for (unsigned i = 0; i < n; ++i) {
        TInit.rotation(i) = Rg[a * per + i];
        TInit.translation(i) = tg[a * per + i];
      }
      agents[a]->setLiftingMatrix(YLift);                                             // :928
      if (multi_init) {
        agents[a]->initialize();                                                      // :348 (local frame, odometry)
        if (a == 0) agents[a]->initializeInGlobalFrame(Pose(d));                      // :353 (robot 0 only)
      } else {
        agents[a]->initialize(&TInit);                                                // :348
        agents[a]->initializeInGlobalFrame(Pose(d));                                  // :353
      }
      std::printf("robot %u: %u poses, %u odometry, %u private LC, %u shared LC, state %d\n", a, n, agents[a]->numOdom(),
                  agents[a]->numPrivateLCs(), agents[a]->numSharedLCs(), (int)agents[a]->state());
    }
    auto publish = [&](unsigned a) {                                                  // publishPublicPoses, :662-690
      for (unsigned b : agents[a]->getNeighbors()) {
        PoseDict dict;
        if (!agents[a]->getSharedPoseDictWithNeighbor(dict, b)) throw std::runtime_error("getSharedPoseDictWithNeighbor");
        agents[b]->updateNeighborPoses(a, dict);                                      // publicPosesCallback, :1276
        if (agents[a]->accelerated()) {
          PoseDict aux;
          if (!agents[a]->getAuxSharedPoseDictWithNeighbor(aux, b)) throw std::runtime_error("getAuxSharedPoseDict");
          agents[b]->updateAuxNeighborPoses(a, aux);                                  // :1278
        }
      }
      agents[a]->clearPublishRequest();
    };
    if (multi_init) {
      // INITIALIZE rounds (:1091-1159): initialised robots publish; the others initialise themselves inside
      // updateNeighborPoses; the leader re-issues INITIALIZE until everybody is in (:1133-1136)
      for (unsigned round = 0; round < N; ++round)
        for (unsigned a = 0; a < N; ++a)
          if (agents[a]->state() == PGOAgentState::INITIALIZED) publish(a);
      for (unsigned a = 0; a < N; ++a)
        if (agents[a]->state() != PGOAgentState::INITIALIZED) throw std::runtime_error("cross-robot initialisation failed");
      std::printf("multi-robot initialisation: all %u robots initialized\n", N);
    }
    for (unsigned a = 0; a < N; ++a) publish(a);                                      // INITIALIZE, :1100
    // the leader's anchor reaches everyone (publishAnchor :412-441 -> setGlobalAnchor :939)
    Matrix anchor;
    agents[0]->getSharedPose(0, anchor);
    for (unsigned a = 1; a < N; ++a) agents[a]->setGlobalAnchor(anchor);

    int it = 0, terminated_at = -1;
    if (mode == "async") {
      // runOnceAsynchronous (:119-127): every robot's own thread optimises; the wrapper publishes when asked to.
      // max_iters is the wall time in milliseconds here.
      const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(max_iters);
      long published = 0;
      while (std::chrono::steady_clock::now() < t_end)
        for (unsigned a = 0; a < N; ++a)
          if (agents[a]->publishAsyncRequested()) {
            agents[a]->clearPublishAsyncRequest();
            publish(a);
            ++published;
          }
      for (unsigned a = 0; a < N; ++a) agents[a]->endOptimizationLoop();
      std::printf("async: %ld publications\n", published);
      it = max_iters;
    }
    for (; it < max_iters; ++it) {
      const unsigned sel = (unsigned)it % N;                                          // RoundRobin, :464-472
      for (unsigned a = 0; a < N; ++a)
        if (a != sel) {
          agents[a]->iterate(false);                                                  // UPDATE handler, :1185
          if (agents[a]->publishRequested()) publish(a);
        }
      agents[sel]->iterate(true);                                                     // runOnceSynchronous, :160
      if (agents[sel]->publishRequested()) publish(sel);
      for (unsigned a = 0; a < N; ++a)                                                // publishStatus, :183 / :1186
        for (unsigned b = 0; b < N; ++b)
          if (a != b) agents[b]->setNeighborStatus(agents[a]->getStatus());
      if (sel == 0 && agents[0]->shouldTerminate()) {                                 // :207-208
        terminated_at = it + 1;
        break;
      }
    }
    std::printf("iterations %d terminated_at %d\n", terminated_at > 0 ? terminated_at : it, terminated_at);
    // the leader re-publishes its anchor with every iteration (publishAnchor, :178-180, :412-441)
    agents[0]->anchorFirstPose();
    agents[0]->getSharedPose(0, anchor);
    for (unsigned a = 1; a < N; ++a) agents[a]->setGlobalAnchor(anchor);
    for (unsigned a = 0; a < N; ++a) {
      const Matrix X = agents[a]->getX();
      std::ofstream out(prefix + "_X" + std::to_string(a) + ".bin", std::ios::binary);
      out.write(reinterpret_cast<const char *>(X.data()), (std::streamsize)(X.size() * sizeof(double)));
      PoseArray T(d, 1);
      const bool ok = agents[a]->getTrajectoryInGlobalFrame(T);                       // :624
      std::printf("robot %u: iteration %u rel_change %.17g |X| %.17g trajectory %d first_t %.12g %.12g %.12g\n", a,
                  agents[a]->iteration_number(), agents[a]->relativeChange(), X.norm(), (int)ok,
                  ok ? T.translation(0)(0) : 0.0, ok ? T.translation(0)(1) : 0.0, ok ? T.translation(0)(2) : 0.0);
      std::ofstream tout(prefix + "_T" + std::to_string(a) + ".bin", std::ios::binary);
      if (ok) tout.write(reinterpret_cast<const char *>(T.getData().data()), (std::streamsize)(T.getData().size() * sizeof(double)));
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "shim_harness: %s\n", e.what());
    return 3;
  }
  return 0;
}
