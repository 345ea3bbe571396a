// local_initialization_method = GNC_TLS (src/PGOAgentROSNode.cpp:110-111) through the DPGO:: shim: a single robot holding
// smallGrid3D with a handful of its loop closures replaced by garbage.  The robust initialisation has to end up close to
// the trajectory of the CLEAN problem; the Chordal initialisation on the same corrupted data does not.
// Links the shim against whichever back end the test chooses (oracle/abi_on_oracle.cpp on the CPU, libdpgo_b200.so on a GPU).
// usage: robust_init_check <smallGrid3D.g2o>      prints "chordal_error gnc_error" (RMS position error, metres)
#include <cmath>
#include <cstdio>
#include <random>

#include "DPGO/PGOAgent.h"

using namespace DPGO;

// the local guess as the wrapper would see it: initialize() + initializeInGlobalFrame(identity) + getTrajectoryInGlobalFrame()
static std::vector<double> initialTrajectory(const std::vector<RelativeSEMeasurement> &meas, InitializationMethod method) {
  PGOAgentParameters params(3, 5, 1);
  params.localInitializationMethod = method;
  params.robustCostParams.GNCBarc = RobustCost::computeErrorThresholdAtQuantile(0.9, 3);   // PGOAgentROSNode.cpp:198-203
  params.robustCostParams.GNCMuStep = 2.0;
  params.robustCostParams.GNCInitMu = 1e-5;
  PGOAgent agent(0, params);
  for (const auto &m : meas) agent.addMeasurement(m);
  agent.initialize();
  agent.initializeInGlobalFrame(Pose(3));
  PoseArray T(3, agent.num_poses());
  if (!agent.getTrajectoryInGlobalFrame(T)) throw std::runtime_error("no trajectory");
  std::vector<double> t;
  for (unsigned i = 0; i < agent.num_poses(); ++i)
    for (int a = 0; a < 3; ++a) t.push_back(T.translation(i)(a));
  return t;
}

static double rms(const std::vector<double> &a, const std::vector<double> &b) {
  double s = 0;
  for (size_t k = 0; k < a.size(); ++k) s += (a[k] - b[k]) * (a[k] - b[k]);
  return std::sqrt(s / (a.size() / 3));
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  size_t n = 0;
  std::vector<RelativeSEMeasurement> clean = read_g2o_file(argv[1], n);
  for (auto &m : clean) {
    m.r1 = m.r2 = 0;
    if (m.p1 + 1 == m.p2) m.fixedWeight = true;   // odometry is trusted (src/utils.cpp:147-149)
  }
  // reference: the robust initialisation of the clean problem (every loop closure an inlier)
  const std::vector<double> truth = initialTrajectory(clean, InitializationMethod::GNC_TLS);
  std::vector<RelativeSEMeasurement> dirty = clean;
  std::mt19937 rng(7);
  std::uniform_real_distribution<double> u(-4.0, 4.0);
  int corrupted = 0;
  for (auto &m : dirty) {
    if (m.fixedWeight || corrupted >= 8 || (rng() % 5) != 0) continue;
    m.t << u(rng), u(rng), u(rng);
    Matrix R(3, 3);   // a rotation by 180 degrees about z composed with the measurement
    R << -1, 0, 0, 0, -1, 0, 0, 0, 1;
    m.R = R * m.R;
    ++corrupted;
  }
  const std::vector<double> chordal = initialTrajectory(dirty, InitializationMethod::Chordal);
  const std::vector<double> robust = initialTrajectory(dirty, InitializationMethod::GNC_TLS);
  std::printf("%d %.6f %.6f\n", corrupted, rms(chordal, truth), rms(robust, truth));
  return 0;
}
