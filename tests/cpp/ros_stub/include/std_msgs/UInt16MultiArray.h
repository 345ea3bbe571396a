#ifndef ROS_STUB_STD_MSGS_UINT16MULTIARRAY_H
#define ROS_STUB_STD_MSGS_UINT16MULTIARRAY_H
#include <cstdint>
#include <memory>
#include <vector>
namespace std_msgs {
struct UInt16MultiArray {
  std::vector<uint16_t> data;
  typedef std::shared_ptr<UInt16MultiArray> Ptr;
  typedef std::shared_ptr<const UInt16MultiArray> ConstPtr;
};
typedef std::shared_ptr<UInt16MultiArray> UInt16MultiArrayPtr;
typedef std::shared_ptr<const UInt16MultiArray> UInt16MultiArrayConstPtr;
}  // namespace std_msgs
#endif
