// The three cases of the reference's own unit test (tests/testUtils.cpp:16-70: MatrixMsg round trip, status round
// trip + enum equality) against the ROS-free wire layer, plus PublicPoses / weights / Command byte streams and the
// CSV iteration log columns (src/PGOAgentROS.cpp:863-864).
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

#include "dpgo_ros_wire/wire.h"

using namespace DPGO;
using namespace dpgo_ros;

#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char **argv) {
  const std::string tmp = argc > 1 ? argv[1] : "/tmp";
  // testUtils.cpp:16-26 -- matrix -> message -> matrix, tolerance 1e-6
  Matrix M(3, 4);
  M << 1.5, -2, 3, 4, 5, 6.25, 7, 8, 9, 10, 11, -12.125;
  const MatrixMsg mm = MatrixToMsg(M);
  CHECK(mm.rows == 3 && mm.cols == 4 && mm.values.size() == 12);
  CHECK(mm.values[1] == -2 && mm.values[4] == 5);   // row-major on the wire (src/utils.cpp:27-32)
  CHECK((MatrixFromMsg(mm) - M).norm() < 1e-6);
  // device pose buffer (r x 4 column-major) <-> message
  double buf[20], back[20];
  for (int k = 0; k < 20; ++k) buf[k] = 0.25 * k - 1;
  const MatrixMsg pm = PoseBufferToMsg(buf, 5);
  CHECK(pm.rows == 5 && pm.cols == 4 && pm.values[1] == buf[5] && pm.values[4] == buf[1]);
  PoseBufferFromMsg(pm, back);
  for (int k = 0; k < 20; ++k) CHECK(back[k] == buf[k]);
  // testUtils.cpp:55-70 -- status round trip, enum values shared with Status.msg
  PGOAgentStatus st(3, PGOAgentState::INITIALIZED, 2, 117, true, 0.0625);
  const Status sm = statusToMsg(st);
  const PGOAgentStatus st2 = statusFromMsg(decodeStatus(encode(sm)));
  CHECK(st2.agentID == 3 && st2.state == PGOAgentState::INITIALIZED && st2.instanceNumber == 2 &&
        st2.iterationNumber == 117 && st2.readyToTerminate && st2.relativeChange == 0.0625);
  CHECK((int)PGOAgentState::WAIT_FOR_DATA == Status::WAIT_FOR_DATA &&
        (int)PGOAgentState::WAIT_FOR_INITIALIZATION == Status::WAIT_FOR_INITIALIZATION &&
        (int)PGOAgentState::INITIALIZED == Status::INITIALIZED);
  CHECK(encode(sm).size() == 2 * 4 + 1 + 1 + 4);
  PGOAgentStatus fine(0, PGOAgentState::INITIALIZED, 0, 1, false, 0.1);   // 0.1 is not a float32
  CHECK(statusFromMsg(statusToMsg(fine)).relativeChange != 0.1 &&
        std::fabs(statusFromMsg(statusToMsg(fine)).relativeChange - 0.1) < 1e-8);
  // PublicPoses: dictionary -> message -> bytes -> message -> dictionary
  PoseDict dict;
  for (unsigned f : {4u, 9u, 31u}) {
    Matrix X(5, 4);
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 4; ++j) X(i, j) = f + 0.1 * i - 0.01 * j;
    dict.emplace(PoseID(2, f), X);
  }
  const PublicPoses pp = PublicPosesToMsg(dict, 2, 0, 5, 1, 42, true);
  const std::vector<uint8_t> bytes = encode(pp);
  CHECK(bytes.size() == 5 * 2 + 1 + (4 + 3 * 4) + 4 + 3 * (2 + 2 + 4 + 20 * 8));
  const PublicPoses pp2 = decodePublicPoses(bytes);
  CHECK(pp2.robot_id == 2 && pp2.destination_robot_id == 5 && pp2.iteration_number == 42 && pp2.is_auxiliary &&
        pp2.pose_ids.size() == 3 && pp2.pose_ids[2] == 31);
  const PoseDict dict2 = PublicPosesFromMsg(pp2);
  CHECK(dict2.size() == 3);
  for (const auto &kv : dict) CHECK((dict2.at(kv.first).getData() - kv.second.getData()).norm() == 0.0);
  std::vector<uint8_t> cut(bytes.begin(), bytes.end() - 3);
  bool threw = false;
  try {
    decodePublicPoses(cut);
  } catch (const std::exception &) {
    threw = true;
  }
  CHECK(threw);
  // weights: the lower ID owns a shared edge (:732)
  PoseGraph g(1, 5, 3);
  Matrix R = Matrix::Identity(3, 3), t(3, 1);
  RelativeSEMeasurement a(1, 3, 7, 2, R, t, 10, 1), b(0, 1, 5, 6, R, t, 10, 1), c(1, 3, 8, 4, R, t, 10, 1);
  a.weight = 0.25;
  c.weight = 1.0;
  c.fixedWeight = true;
  g.addMeasurement(a);
  g.addMeasurement(b);
  g.addMeasurement(c);
  const RelativeMeasurementWeights w3 = decodeRelativeMeasurementWeights(encode(MeasurementWeightsToMsg(g, 1, 0, 3)));
  CHECK(w3.weights.size() == 2 && w3.weights[0] == 0.25f && w3.src_pose_ids[1] == 8 && w3.fixed_weights[1] == 1 &&
        w3.destination_robot_id == 3);
  CHECK(MeasurementWeightsToMsg(g, 1, 0, 0).weights.empty());   // robot 0 owns the edge it shares with robot 1
  // Command + RoundRobin token
  Command cmd;
  cmd.command = Command::UPDATE;
  cmd.publishing_robot = 1;
  cmd.executing_robot = 2;
  cmd.executing_iteration = 77;
  cmd.active_robots = {0, 1, 2, 4};
  const Command cmd2 = decodeCommand(encode(cmd));
  CHECK(cmd2.command == 1 && cmd2.executing_robot == 2 && cmd2.executing_iteration == 77 && cmd2.active_robots.size() == 4);
  CHECK(Command::REQUEST_POSE_GRAPH == 0 && Command::UPDATE_WEIGHT == 5 && Command::NOOP == 8);
  CHECK(nextRobotRoundRobin(2, {true, true, true, false, true}) == 4 && nextRobotRoundRobin(4, {true, true, true, false, true}) == 0);
  // CSV log: header and one row in the reference's column order
  IterationLog log;
  const std::string path = tmp + "/dpgo_log_test.csv";
  CHECK(log.open(path));
  CHECK(log.logIteration(3, 0, 8, 12, 312, 4096, 0.001, 1.5, 0.25));
  CHECK(log.logString("TERMINATE"));
  std::ifstream in(path);
  std::string l1, l2, l3;
  std::getline(in, l1);
  std::getline(in, l2);
  std::getline(in, l3);
  CHECK(l1 == "robot_id, cluster_id, num_active_robots, iteration, num_poses, bytes_received, iter_time_sec, total_time_sec, rel_change ");
  CHECK(l2 == "3,0,8,12,312,4096,0.001,1.5,0.25" && l3 == "TERMINATE");
  std::printf("wire checks ok\n");
  return 0;
}
