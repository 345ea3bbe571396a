// Host-only checks of the DPGO:: shim (no GPU): the value-type operations dpgo_ros relies on
// (tests/testUtils.cpp:16-70, src/utils.cpp:20-71,154-164), PoseGraph bookkeeping, the g2o / CSV loaders and
// RobustCost::computeErrorThresholdAtQuantile (src/PGOAgentROSNode.cpp:201).
#include <cmath>
#include <cstdio>
#include <string>

#include "DPGO/DPGO_robust.h"
#include "DPGO/DPGO_utils.h"
#include "DPGO/PGOLogger.h"
#include "DPGO/PoseGraph.h"

using namespace DPGO;

#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char **argv) {
  const std::string data = argc > 1 ? argv[1] : "data";
  // Matrix: comma initialiser is row-major, storage column-major (tests/testUtils.cpp:17-25)
  Matrix M(2, 3);
  M << 1, 2, 3, 4, 5, 6;
  CHECK(M(0, 2) == 3 && M(1, 0) == 4 && M.data()[1] == 4 && M.rows() == 2 && M.cols() == 3);
  CHECK(std::fabs((M - M).norm()) == 0.0);
  const Matrix P = M * M.transpose();
  CHECK(P.rows() == 2 && P(0, 0) == 14 && P(0, 1) == 32 && P(1, 1) == 77);
  {
    // storage: <= 32 coefficients live inside the object, larger matrices on the heap; copies, moves, swaps and
    // assignments across the two regimes keep values and shapes (round 2's small-buffer Matrix)
    Matrix small(5, 4), big(9, 9);
    for (size_t i = 0; i < small.size(); ++i) small.data()[i] = 1.0 + i;
    for (size_t i = 0; i < big.size(); ++i) big.data()[i] = 100.0 + i;
    Matrix c1 = small, c2 = big;
    CHECK(c1.data() != small.data() && c2.data() != big.data() && c1(4, 3) == 20 && c2(8, 8) == 180);
    Matrix m1 = std::move(c1), m2 = std::move(c2);
    CHECK(m1.rows() == 5 && m1(0, 1) == 6 && m2.cols() == 9 && m2(1, 0) == 101);
    m1 = big;      // inline -> heap
    m2 = small;    // heap -> inline
    CHECK(m1.rows() == 9 && m1(8, 8) == 180 && m2.rows() == 5 && m2(4, 3) == 20);
    m1 = m1;       // self-assignment
    CHECK(m1(8, 8) == 180 && (m1 - big).norm() == 0.0 && (m2 - small).norm() == 0.0);
    std::vector<Matrix> v(3, small);
    v.push_back(big);
    v.insert(v.begin(), big);
    CHECK(v.front()(8, 8) == 180 && v[1](4, 3) == 20 && v.back()(0, 0) == 100);
    Matrix edge(8, 4);   // exactly 32 coefficients: the largest inline shape (r = 8)
    edge(7, 3) = 3.5;
    const Matrix ecopy = edge;
    CHECK(ecopy(7, 3) == 3.5 && ecopy(0, 0) == 0.0);
  }
  Matrix B = Matrix::Zero(3, 4);
  B.block(0, 0, 3, 3) = Matrix::Identity(3, 3);
  Matrix tcol(3, 1);
  tcol << 7, 8, 9;
  B.block(0, 3, 3, 1) = tcol;
  CHECK(B(2, 3) == 9 && B(1, 1) == 1 && static_cast<const Matrix &>(B).block(0, 3, 3, 1).norm() == std::sqrt(194.0));
  CHECK(Vector::Zero(5).rows() == 5);
  // poses (src/PGOAgentROS.cpp:353-357, 1420-1421; src/utils.cpp:154-164)
  Pose T(3);
  CHECK(T.getData().rows() == 3 && T.getData().cols() == 4 && T.rotation()(1, 1) == 1.0);
  PoseArray traj(3, 4);
  traj.translation(2) = tcol;
  CHECK(traj.getData().cols() == 16 && traj.pose(2)(1, 3) == 8 && traj.d() == 3 && traj.n() == 4);
  LiftedPose X(5, 3);
  X.translation() = Matrix(Vector::Zero(5));
  PoseDict dict;
  dict.emplace(PoseID(1, 7), Matrix::Identity(5, 4));   // LiftedPose implicitly from Matrix (:1273)
  CHECK(dict.begin()->first.frame_id == 7 && dict.begin()->second.getData()(2, 2) == 1.0);
  // status / enums (tests/testUtils.cpp:56-69, msg/Status.msg:1-3)
  PGOAgentStatus st(3, PGOAgentState::INITIALIZED, 1, 42, true, 0.5);
  CHECK(st.agentID == 3 && st.iterationNumber == 42 && st.readyToTerminate && (int)PGOAgentState::WAIT_FOR_DATA == 0 &&
        (int)PGOAgentState::WAIT_FOR_INITIALIZATION == 1 && (int)PGOAgentState::INITIALIZED == 2);
  CHECK(EdgeID(PoseID(0, 1), PoseID(1, 4)).isSharedLoopClosure() && EdgeID(PoseID(2, 1), PoseID(2, 2)).isOdometry() &&
        EdgeID(PoseID(2, 1), PoseID(2, 5)).isPrivateLoopClosure());
  // chi-square threshold: sqrt(chi2inv(0.9; 3)) = 2.50028
  CHECK(std::fabs(RobustCost::computeErrorThresholdAtQuantile(0.9, 3) - 2.500277) < 1e-5);
  CHECK(std::fabs(chi2inv(0.5, 2) - 2.0 * std::log(2.0)) < 1e-10);
  // g2o loader: smallGrid3D has 125 poses / 297 edges, info diag 100 / 25 -> tau 100, kappa 12.5
  size_t n = 0;
  const auto meas = read_g2o_file(data + "/smallGrid3D.g2o", n);
  CHECK(n == 125 && meas.size() == 297);
  CHECK(std::fabs(meas[0].tau - 100.0) < 1e-9 && std::fabs(meas[0].kappa - 12.5) < 1e-9);
  CHECK(std::fabs((meas[0].R * meas[0].R.transpose() - Matrix::Identity(3, 3)).norm()) < 1e-12);
  // pose graph: contiguous 2-robot split (62 / 63 poses, 30 shared loop closures)
  PoseGraph g0(0, 5, 3), g1(1, 5, 3);
  for (auto m : meas) {
    const size_t per = n / 2, a = std::min<size_t>(m.p1 / per, 1), b = std::min<size_t>(m.p2 / per, 1);
    m.p1 -= a * per; m.p2 -= b * per; m.r1 = a; m.r2 = b;
    g0.addMeasurement(m);
    g1.addMeasurement(m);
    CHECK(!g0.addMeasurement(m));   // duplicates refused (hasMeasurement, :276)
  }
  CHECK(g0.n() == 62 && g1.n() == 63 && g0.numOdometry() == 61 && g1.numOdometry() == 62);
  CHECK(g0.numSharedLoopClosures() == 30 && g1.numSharedLoopClosures() == 30);
  CHECK(g0.numPrivateLoopClosures() == 71 && g1.numPrivateLoopClosures() == 73);
  CHECK(g0.activeNeighborIDs().count(1) == 1 && g0.activeNeighborPublicPoseIDs().size() == 25);
  const auto &lc = g0.sharedLoopClosures()[0];
  CHECK(g0.hasMeasurement(PoseID(lc.r1, lc.p1), PoseID(lc.r2, lc.p2)));
  g0.sharedLoopClosures()[0].weight = 0.0;
  g0.privateLoopClosures()[0].weight = 0.5;
  const auto stat = g0.statistics();
  CHECK(stat.reject_loop_closures == 1 && stat.undecided_loop_closures == 1 && stat.total_loop_closures == 101);
  // tunnels CSV (data/tunnels/robot0/measurements.csv: 828 rows, kappa 10000, tau 100)
  PGOLogger logger("");
  const auto csv = logger.loadMeasurements(data + "/tunnels/robot0/measurements.csv", false);
  CHECK(csv.size() == 828 && csv[0].kappa == 10000 && csv[0].tau == 100 && csv[0].weight == 1.0);
  // rounding helper
  Matrix Rn = meas[3].R;
  Rn(0, 1) += 1e-3;
  const Matrix Rp = projectToRotationGroup(Rn);
  CHECK(std::fabs((Rp * Rp.transpose() - Matrix::Identity(3, 3)).norm()) < 1e-12 && (Rp - meas[3].R).norm() < 2e-3);
  const Matrix Y = fixedStiefelVariable(3, 5);
  CHECK(std::fabs((Y.transpose() * Y - Matrix::Identity(3, 3)).norm()) < 1e-14);
  // robust transform averaging (cross-robot initialisation): 7 consistent votes + 3 outliers
  {
    Matrix T0(3, 4);
    T0.block(0, 0, 3, 3) = meas[5].R;
    T0.block(0, 3, 3, 1) = tcol;
    std::vector<Matrix> cands;
    for (int k = 0; k < 7; ++k) {
      Matrix C = T0;
      C(0, 3) += 0.01 * (k - 3);
      C(1, 0) += 1e-3 * (k - 3);
      cands.push_back(C);
    }
    for (int k = 0; k < 3; ++k) {
      Matrix C = T0;
      C.block(0, 0, 3, 3) = meas[20 + k].R;
      C(2, 3) += 5.0 + k;
      cands.push_back(C);
    }
    Matrix avg;
    unsigned inl = 0;
    CHECK(robustTransformAverage(cands, 0.2, 1.0, 2, avg, &inl) && inl == 7);
    CHECK((avg - T0).norm() < 5e-3);
    const Matrix Ra = static_cast<const Matrix &>(avg).block(0, 0, 3, 3);
    CHECK(std::fabs((Ra * Ra.transpose() - Matrix::Identity(3, 3)).norm()) < 1e-12);
    CHECK(!robustTransformAverage(cands, 0.2, 1.0, 8, avg));   // not enough inliers
    CHECK((se3Compose(T0, se3Inverse(T0)) - Matrix::Identity(3, 4)).norm() < 1e-12);
  }
  std::printf("shim host checks ok\n");
  return 0;
}
