/* pose_graph_tools_msgs of the ROS stand-in: the fields dpgo_ros touches (src/utils.cpp:108-152, 237-260;
 * src/PGOAgentROS.cpp:241-330; src/PGODatasetPublisherNode.cpp:62-72).  TEST INFRASTRUCTURE. */
#ifndef ROS_STUB_POSE_GRAPH_TOOLS_MSGS_POSEGRAPH_H
#define ROS_STUB_POSE_GRAPH_TOOLS_MSGS_POSEGRAPH_H
#include <array>

#include "geometry_msgs/Pose.h"
namespace pose_graph_tools_msgs {
struct PoseGraphEdge {
  enum { ODOM = 0, LOOPCLOSE = 1, LANDMARK = 2, REJECTED_LOOPCLOSE = 3, PRIOR = 4 };
  std_msgs::Header header;
  uint64_t key_from = 0, key_to = 0;
  int32_t robot_from = 0, robot_to = 0;
  int32_t type = 0;
  geometry_msgs::Pose pose;
  std::array<double, 36> covariance{};
};
struct PoseGraphNode {
  std_msgs::Header header;
  int32_t robot_id = 0;
  uint64_t key = 0;
  geometry_msgs::Pose pose;
};
struct PoseGraph {
  std_msgs::Header header;
  std::vector<PoseGraphNode> nodes;
  std::vector<PoseGraphEdge> edges;
  typedef std::shared_ptr<const PoseGraph> ConstPtr;
};
typedef std::shared_ptr<const PoseGraph> PoseGraphConstPtr;
}  // namespace pose_graph_tools_msgs
#endif
