#ifndef ROS_STUB_GEOMETRY_MSGS_POINT_H
#define ROS_STUB_GEOMETRY_MSGS_POINT_H
namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
struct Point32 {
  float x = 0, y = 0, z = 0;
};
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 0;
};
}  // namespace geometry_msgs
#endif
