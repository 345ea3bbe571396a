/* std_msgs/Header of the ROS stand-in (tests/cpp/ros_stub).  TEST INFRASTRUCTURE. */
#ifndef ROS_STUB_STD_MSGS_HEADER_H
#define ROS_STUB_STD_MSGS_HEADER_H
#include <cstdint>
#include <memory>
#include <string>

#include "ros/ros.h"
namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs
#endif
