#ifndef ROS_STUB_GEOMETRY_MSGS_POSE_H
#define ROS_STUB_GEOMETRY_MSGS_POSE_H
#include <memory>
#include <vector>

#include "geometry_msgs/Point.h"
#include "std_msgs/Header.h"
namespace geometry_msgs {
struct Pose {
  Point position;
  Quaternion orientation;
};
struct PoseStamped {
  std_msgs::Header header;
  Pose pose;
};
struct PoseArray {
  std_msgs::Header header;
  std::vector<Pose> poses;
  typedef std::shared_ptr<const PoseArray> ConstPtr;
};
typedef std::shared_ptr<const PoseArray> PoseArrayConstPtr;
}  // namespace geometry_msgs
#endif
