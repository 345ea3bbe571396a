#ifndef ROS_STUB_STD_MSGS_COLORRGBA_H
#define ROS_STUB_STD_MSGS_COLORRGBA_H
namespace std_msgs {
struct ColorRGBA {
  float r = 0, g = 0, b = 0, a = 0;
};
}  // namespace std_msgs
#endif
