/* ros/console.h -- rosconsole macros of the ROS stand-in (see ros/ros.h).  TEST INFRASTRUCTURE. */
#ifndef ROS_STUB_CONSOLE_H
#define ROS_STUB_CONSOLE_H
#include <sstream>

#include "ros/ros.h"

#define ROS_STUB_LOG_(level, tag, ...) ::ros::sim::log(level, tag, __VA_ARGS__)
#define ROS_STUB_STREAM_(level, tag, args)                      \
  do {                                                          \
    if ((level) <= ::ros::sim::world().log_level) {             \
      std::ostringstream ros_stub_ss_;                          \
      ros_stub_ss_ << args;                                     \
      ::ros::sim::log(level, tag, "%s", ros_stub_ss_.str().c_str()); \
    }                                                           \
  } while (0)
#define ROS_STUB_THROTTLE_(period, level, tag, ...)                                  \
  do {                                                                               \
    static double ros_stub_last_ = -1e300;                                           \
    if (::ros::sim::world().now - ros_stub_last_ >= (period)) {                      \
      ros_stub_last_ = ::ros::sim::world().now;                                      \
      ::ros::sim::log(level, tag, __VA_ARGS__);                                      \
    }                                                                                \
  } while (0)

#define ROS_DEBUG(...) ROS_STUB_LOG_(3, "DEBUG", __VA_ARGS__)
#define ROS_INFO(...) ROS_STUB_LOG_(2, "INFO", __VA_ARGS__)
#define ROS_WARN(...) ROS_STUB_LOG_(1, "WARN", __VA_ARGS__)
#define ROS_ERROR(...) ROS_STUB_LOG_(1, "ERROR", __VA_ARGS__)
#define ROS_DEBUG_STREAM(args) ROS_STUB_STREAM_(3, "DEBUG", args)
#define ROS_INFO_STREAM(args) ROS_STUB_STREAM_(2, "INFO", args)
#define ROS_WARN_STREAM(args) ROS_STUB_STREAM_(1, "WARN", args)
#define ROS_ERROR_STREAM(args) ROS_STUB_STREAM_(1, "ERROR", args)
#define ROS_INFO_THROTTLE(period, ...) ROS_STUB_THROTTLE_(period, 2, "INFO", __VA_ARGS__)
#define ROS_WARN_THROTTLE(period, ...) ROS_STUB_THROTTLE_(period, 1, "WARN", __VA_ARGS__)
#define ROS_ERROR_THROTTLE(period, ...) ROS_STUB_THROTTLE_(period, 1, "ERROR", __VA_ARGS__)
#endif
