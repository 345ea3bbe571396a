#ifndef ROS_STUB_SENSOR_MSGS_POINTCLOUD_H
#define ROS_STUB_SENSOR_MSGS_POINTCLOUD_H
#include "geometry_msgs/Pose.h"
namespace sensor_msgs {
struct PointCloud {
  std_msgs::Header header;
  std::vector<geometry_msgs::Point32> points;
};
}  // namespace sensor_msgs
#endif
