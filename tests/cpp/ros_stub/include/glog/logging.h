/* glog/logging.h of the ROS stand-in: CHECK / CHECK_EQ / LOG as dpgo_ros uses them
 * (src/PGOAgentROS.cpp:130, 684, 713; src/utils.cpp:283-284).  TEST INFRASTRUCTURE. */
#ifndef ROS_STUB_GLOG_H
#define ROS_STUB_GLOG_H
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace glog_stub {
class Fatal {
 public:
  Fatal(const char *file, int line, const char *what) { ss_ << file << ":" << line << " Check failed: " << what << " "; }
  [[noreturn]] ~Fatal() {
    std::cerr << ss_.str() << std::endl;
    std::abort();
  }
  template <class T>
  Fatal &operator<<(const T &v) {
    ss_ << v;
    return *this;
  }

 private:
  std::ostringstream ss_;
};
class Sink {
 public:
  template <class T>
  Sink &operator<<(const T &) { return *this; }
};
struct Voidify {
  void operator&(const Fatal &) {}
  void operator&(const Sink &) {}
};
}  // namespace glog_stub
#define CHECK(cond) (cond) ? (void)0 : ::glog_stub::Voidify() & ::glog_stub::Fatal(__FILE__, __LINE__, #cond)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define LOG(severity) ::glog_stub::Sink()
#endif
