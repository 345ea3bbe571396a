/*
 * dpgo_ros_wire/wire.h -- ROS-free mirror of dpgo_ros's wire formats, so that a
 * PGOAgentROS sitting on the B200 path -- or a bridge that copies device buffers to a socket -- produces and
 * consumes exactly the payloads of the reference:
 *
 *   msg/MatrixMsg.msg:1-3            rows, cols, float64[] values ROW-MAJOR (src/utils.cpp:20-49)
 *   msg/PublicPoses.msg:1-8          robot / cluster / destination ids, instance + iteration number, is_auxiliary,
 *                                    uint32[] pose_ids, MatrixMsg[] poses      (src/PGOAgentROS.cpp:662-690)
 *   msg/Status.msg:5-11              relative_change crosses the wire as float32 (src/utils.cpp:262-281)
 *   msg/Command.msg:1-17             opcodes 0-8 + scheduling fields          (src/PGOAgentROS.cpp:481-504)
 *   msg/RelativeMeasurementWeights.msg:1-9   float32 weights, lower-ID owner rule (src/PGOAgentROS.cpp:721-754)
 *
 * The structs carry the .msg field lists (without std_msgs/Header); encode()/decode() use the ROS 1 serialisation
 * rules (little endian, arrays prefixed by a uint32 length, bool = 1 byte) so the byte streams are interchangeable.
 * Also here: the per-round CSV iteration log in the reference's column layout (src/PGOAgentROS.cpp:853-907).
 * Host-side, not on the hot path.
 */
#ifndef DPGO_ROS_WIRE_H
#define DPGO_ROS_WIRE_H

#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "DPGO/DPGO_types.h"
#include "DPGO/PoseGraph.h"

namespace dpgo_ros {

struct MatrixMsg {
  uint16_t rows = 0, cols = 0;
  std::vector<double> values;  // row-major
};
struct PublicPoses {
  uint16_t robot_id = 0, cluster_id = 0, destination_robot_id = 0, instance_number = 0, iteration_number = 0;
  bool is_auxiliary = false;
  std::vector<uint32_t> pose_ids;
  std::vector<MatrixMsg> poses;
};
struct Status {
  enum : uint8_t { WAIT_FOR_DATA = 0, WAIT_FOR_INITIALIZATION = 1, INITIALIZED = 2 };
  uint16_t instance_number = 0, iteration_number = 0, robot_id = 0, cluster_id = 0;
  uint8_t state = 0;
  bool ready_to_terminate = false;
  float relative_change = 0;
};
struct Command {
  enum : uint8_t { REQUEST_POSE_GRAPH = 0, UPDATE = 1, TERMINATE = 2, HARD_TERMINATE = 3, INITIALIZE = 4,
                   UPDATE_WEIGHT = 5, RECOVER = 6, SET_ACTIVE_ROBOTS = 7, NOOP = 8 };
  uint8_t command = NOOP;
  uint16_t cluster_id = 0, publishing_robot = 0, executing_robot = 0, executing_iteration = 0;
  std::vector<uint16_t> active_robots;
};
struct RelativeMeasurementWeights {
  uint16_t robot_id = 0, cluster_id = 0, destination_robot_id = 0;
  std::vector<uint16_t> src_robot_ids, dst_robot_ids;
  std::vector<uint32_t> src_pose_ids, dst_pose_ids;
  std::vector<float> weights;
  std::vector<uint8_t> fixed_weights;
};

// ---- codecs between DPGO:: values and messages ---------------------------------------------------------------
inline MatrixMsg MatrixToMsg(const DPGO::Matrix &M) {
  MatrixMsg msg;
  msg.rows = (uint16_t)M.rows();
  msg.cols = (uint16_t)M.cols();
  msg.values.resize(M.size());
  size_t k = 0;
  for (size_t i = 0; i < M.rows(); ++i)
    for (size_t j = 0; j < M.cols(); ++j) msg.values[k++] = M(i, j);
  return msg;
}
inline DPGO::Matrix MatrixFromMsg(const MatrixMsg &msg) {
  if (msg.values.size() != (size_t)msg.rows * msg.cols) throw std::invalid_argument("MatrixMsg: size mismatch");
  DPGO::Matrix M(msg.rows, msg.cols);
  for (size_t k = 0; k < msg.values.size(); ++k) M(k / msg.cols, k % msg.cols) = msg.values[k];
  return M;
}
// the device layout (r x 4 column-major per pose, include/dpgo_b200.h) <-> MatrixMsg without a DPGO::Matrix in between
inline MatrixMsg PoseBufferToMsg(const double *pose_colmajor, unsigned r, unsigned cols = 4) {
  MatrixMsg msg;
  msg.rows = (uint16_t)r;
  msg.cols = (uint16_t)cols;
  msg.values.resize((size_t)r * cols);
  for (unsigned i = 0; i < r; ++i)
    for (unsigned j = 0; j < cols; ++j) msg.values[(size_t)i * cols + j] = pose_colmajor[(size_t)j * r + i];
  return msg;
}
inline void PoseBufferFromMsg(const MatrixMsg &msg, double *pose_colmajor) {
  for (unsigned i = 0; i < msg.rows; ++i)
    for (unsigned j = 0; j < msg.cols; ++j) pose_colmajor[(size_t)j * msg.rows + i] = msg.values[(size_t)i * msg.cols + j];
}

inline Status statusToMsg(const DPGO::PGOAgentStatus &s) {
  Status msg;
  msg.robot_id = (uint16_t)s.agentID;
  msg.state = (uint8_t)s.state;
  msg.instance_number = (uint16_t)s.instanceNumber;
  msg.iteration_number = (uint16_t)s.iterationNumber;
  msg.ready_to_terminate = s.readyToTerminate;
  msg.relative_change = (float)s.relativeChange;  // float32 on the wire (msg/Status.msg:11)
  return msg;
}
inline DPGO::PGOAgentStatus statusFromMsg(const Status &msg) {
  return DPGO::PGOAgentStatus(msg.robot_id, static_cast<DPGO::PGOAgentState>(msg.state), msg.instance_number,
                              msg.iteration_number, msg.ready_to_terminate, msg.relative_change);
}

// publishPublicPoses (src/PGOAgentROS.cpp:662-690) / publicPosesCallback (:1255-1284)
inline PublicPoses PublicPosesToMsg(const DPGO::PoseDict &dict, unsigned robot, unsigned cluster, unsigned destination,
                                    unsigned instance, unsigned iteration, bool auxiliary) {
  PublicPoses msg;
  msg.robot_id = (uint16_t)robot;
  msg.cluster_id = (uint16_t)cluster;
  msg.destination_robot_id = (uint16_t)destination;
  msg.instance_number = (uint16_t)instance;
  msg.iteration_number = (uint16_t)iteration;
  msg.is_auxiliary = auxiliary;
  for (const auto &kv : dict) {
    if (kv.first.robot_id != robot) throw std::invalid_argument("PublicPoses: pose of another robot");
    msg.pose_ids.push_back(kv.first.frame_id);
    msg.poses.push_back(MatrixToMsg(kv.second.getData()));
  }
  return msg;
}
inline DPGO::PoseDict PublicPosesFromMsg(const PublicPoses &msg) {
  if (msg.pose_ids.size() != msg.poses.size()) throw std::invalid_argument("PublicPoses: ids / poses mismatch");
  DPGO::PoseDict dict;
  for (size_t k = 0; k < msg.pose_ids.size(); ++k)
    dict.emplace(DPGO::PoseID(msg.robot_id, msg.pose_ids[k]), DPGO::LiftedPose(MatrixFromMsg(msg.poses[k])));
  return dict;
}

// publishMeasurementWeights (:721-754): the weights of the shared loop closures with `neighbor` that THIS robot owns
// (the lower ID owns an edge, :732); measurementWeightsCallback (:1315-1353) applies them on the other end.
inline RelativeMeasurementWeights MeasurementWeightsToMsg(DPGO::PoseGraph &graph, unsigned robot, unsigned cluster,
                                                          unsigned neighbor) {
  RelativeMeasurementWeights msg;
  msg.robot_id = (uint16_t)robot;
  msg.cluster_id = (uint16_t)cluster;
  msg.destination_robot_id = (uint16_t)neighbor;
  if (neighbor <= robot) return msg;  // the other robot owns every edge we share
  for (const auto &m : graph.sharedLoopClosures()) {
    const unsigned other = m.r1 == robot ? (unsigned)m.r2 : (unsigned)m.r1;
    if (other != neighbor) continue;
    msg.src_robot_ids.push_back((uint16_t)m.r1);
    msg.dst_robot_ids.push_back((uint16_t)m.r2);
    msg.src_pose_ids.push_back((uint32_t)m.p1);
    msg.dst_pose_ids.push_back((uint32_t)m.p2);
    msg.weights.push_back((float)m.weight);  // float32 on the wire (msg/RelativeMeasurementWeights.msg:8)
    msg.fixed_weights.push_back(m.fixedWeight ? 1 : 0);
  }
  return msg;
}

// ---- ROS 1 byte streams ----------------------------------------------------------------------------------------
class Writer {
 public:
  template <class T>
  void put(const T &v) {
    const size_t o = buf.size();
    buf.resize(o + sizeof(T));
    std::memcpy(buf.data() + o, &v, sizeof(T));
  }
  template <class T>
  void put_array(const std::vector<T> &v) {
    put<uint32_t>((uint32_t)v.size());
    const size_t o = buf.size();
    buf.resize(o + v.size() * sizeof(T));
    if (!v.empty()) std::memcpy(buf.data() + o, v.data(), v.size() * sizeof(T));
  }
  std::vector<uint8_t> buf;
};
class Reader {
 public:
  Reader(const uint8_t *p, size_t n) : p_(p), n_(n) {}
  template <class T>
  T get() {
    need(sizeof(T));
    T v;
    std::memcpy(&v, p_ + o_, sizeof(T));
    o_ += sizeof(T);
    return v;
  }
  template <class T>
  std::vector<T> get_array() {
    const uint32_t len = get<uint32_t>();
    need((size_t)len * sizeof(T));
    std::vector<T> v(len);
    if (len) std::memcpy(v.data(), p_ + o_, (size_t)len * sizeof(T));
    o_ += (size_t)len * sizeof(T);
    return v;
  }
  bool done() const { return o_ == n_; }

 private:
  void need(size_t k) const {
    if (o_ + k > n_) throw std::out_of_range("dpgo_ros wire: truncated message");
  }
  const uint8_t *p_;
  size_t n_, o_ = 0;
};

inline void encode(Writer &w, const MatrixMsg &m) {
  w.put(m.rows);
  w.put(m.cols);
  w.put_array(m.values);
}
inline MatrixMsg decodeMatrixMsg(Reader &r) {
  MatrixMsg m;
  m.rows = r.get<uint16_t>();
  m.cols = r.get<uint16_t>();
  m.values = r.get_array<double>();
  return m;
}
inline std::vector<uint8_t> encode(const PublicPoses &m) {
  Writer w;
  w.put(m.robot_id);
  w.put(m.cluster_id);
  w.put(m.destination_robot_id);
  w.put(m.instance_number);
  w.put(m.iteration_number);
  w.put<uint8_t>(m.is_auxiliary ? 1 : 0);
  w.put_array(m.pose_ids);
  w.put<uint32_t>((uint32_t)m.poses.size());
  for (const auto &p : m.poses) encode(w, p);
  return w.buf;
}
inline PublicPoses decodePublicPoses(const std::vector<uint8_t> &bytes) {
  Reader r(bytes.data(), bytes.size());
  PublicPoses m;
  m.robot_id = r.get<uint16_t>();
  m.cluster_id = r.get<uint16_t>();
  m.destination_robot_id = r.get<uint16_t>();
  m.instance_number = r.get<uint16_t>();
  m.iteration_number = r.get<uint16_t>();
  m.is_auxiliary = r.get<uint8_t>() != 0;
  m.pose_ids = r.get_array<uint32_t>();
  const uint32_t n = r.get<uint32_t>();
  for (uint32_t k = 0; k < n; ++k) m.poses.push_back(decodeMatrixMsg(r));
  if (!r.done()) throw std::invalid_argument("PublicPoses: trailing bytes");
  return m;
}
inline std::vector<uint8_t> encode(const Status &m) {
  Writer w;
  w.put(m.instance_number);
  w.put(m.iteration_number);
  w.put(m.robot_id);
  w.put(m.cluster_id);
  w.put(m.state);
  w.put<uint8_t>(m.ready_to_terminate ? 1 : 0);
  w.put(m.relative_change);
  return w.buf;
}
inline Status decodeStatus(const std::vector<uint8_t> &bytes) {
  Reader r(bytes.data(), bytes.size());
  Status m;
  m.instance_number = r.get<uint16_t>();
  m.iteration_number = r.get<uint16_t>();
  m.robot_id = r.get<uint16_t>();
  m.cluster_id = r.get<uint16_t>();
  m.state = r.get<uint8_t>();
  m.ready_to_terminate = r.get<uint8_t>() != 0;
  m.relative_change = r.get<float>();
  return m;
}
inline std::vector<uint8_t> encode(const Command &m) {
  Writer w;
  w.put(m.command);
  w.put(m.cluster_id);
  w.put(m.publishing_robot);
  w.put(m.executing_robot);
  w.put(m.executing_iteration);
  w.put_array(m.active_robots);
  return w.buf;
}
inline Command decodeCommand(const std::vector<uint8_t> &bytes) {
  Reader r(bytes.data(), bytes.size());
  Command m;
  m.command = r.get<uint8_t>();
  m.cluster_id = r.get<uint16_t>();
  m.publishing_robot = r.get<uint16_t>();
  m.executing_robot = r.get<uint16_t>();
  m.executing_iteration = r.get<uint16_t>();
  m.active_robots = r.get_array<uint16_t>();
  return m;
}
inline std::vector<uint8_t> encode(const RelativeMeasurementWeights &m) {
  Writer w;
  w.put(m.robot_id);
  w.put(m.cluster_id);
  w.put(m.destination_robot_id);
  w.put_array(m.src_robot_ids);
  w.put_array(m.dst_robot_ids);
  w.put_array(m.src_pose_ids);
  w.put_array(m.dst_pose_ids);
  w.put_array(m.weights);
  w.put_array(m.fixed_weights);
  return w.buf;
}
inline RelativeMeasurementWeights decodeRelativeMeasurementWeights(const std::vector<uint8_t> &bytes) {
  Reader r(bytes.data(), bytes.size());
  RelativeMeasurementWeights m;
  m.robot_id = r.get<uint16_t>();
  m.cluster_id = r.get<uint16_t>();
  m.destination_robot_id = r.get<uint16_t>();
  m.src_robot_ids = r.get_array<uint16_t>();
  m.dst_robot_ids = r.get_array<uint16_t>();
  m.src_pose_ids = r.get_array<uint32_t>();
  m.dst_pose_ids = r.get_array<uint32_t>();
  m.weights = r.get_array<float>();
  m.fixed_weights = r.get_array<uint8_t>();
  return m;
}

// the synchronous schedule's token passing (publishUpdateCommand, :443-479): RoundRobin over the active robots
inline unsigned nextRobotRoundRobin(unsigned current, const std::vector<bool> &active) {
  const unsigned n = (unsigned)active.size();
  for (unsigned k = 1; k <= n; ++k) {
    const unsigned cand = (current + k) % n;
    if (active[cand]) return cand;
  }
  return current;
}

// ---- per-round CSV log, the reference's columns (src/PGOAgentROS.cpp:853-907) ---------------------------------
class IterationLog {
 public:
  bool open(const std::string &filename) {
    if (f_.is_open()) f_.close();
    f_.open(filename);
    if (!f_.is_open()) return false;
    f_ << "robot_id, cluster_id, num_active_robots, iteration, num_poses, bytes_received, "
          "iter_time_sec, total_time_sec, rel_change \n";
    f_.flush();
    return true;
  }
  bool logIteration(unsigned robot, unsigned cluster, unsigned active, unsigned iteration, unsigned poses,
                    size_t bytes_received, double iter_sec, double total_sec, double rel_change) {
    if (!f_.is_open()) return false;
    f_ << robot << "," << cluster << "," << active << "," << iteration << "," << poses << "," << bytes_received << ","
       << iter_sec << "," << total_sec << "," << rel_change << "\n";
    f_.flush();
    return true;
  }
  bool logString(const std::string &s) {  // TERMINATE / HARD_TERMINATE / UPDATE_WEIGHT / TIMEOUT markers
    if (!f_.is_open()) return false;
    f_ << s << "\n";
    f_.flush();
    return true;
  }

 private:
  std::ofstream f_;
};

}  // namespace dpgo_ros
#endif
