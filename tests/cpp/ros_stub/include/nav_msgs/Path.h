#ifndef ROS_STUB_NAV_MSGS_PATH_H
#define ROS_STUB_NAV_MSGS_PATH_H
#include "geometry_msgs/Pose.h"
namespace nav_msgs {
struct Path {
  std_msgs::Header header;
  std::vector<geometry_msgs::PoseStamped> poses;
  typedef std::shared_ptr<const Path> ConstPtr;
};
typedef std::shared_ptr<const Path> PathConstPtr;
}  // namespace nav_msgs
#endif
