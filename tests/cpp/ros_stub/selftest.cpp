// Self-test of the single-process ROS stand-in (tests/cpp/ros_stub/include/ros/ros.h): the properties the wrapper runs
// rely on -- one node at a time, simulated clock, publish-order delivery (to the publisher itself too), timers, services
// in the caller's thread, typed private parameters, partitions.  Prints a trace that must be identical on every run.
#include <ros/ros.h>
#include <std_msgs/UInt16MultiArray.h>

#include <cstdio>
#include <sstream>
#include <thread>

struct Req { int x = 0; };
struct Res { int y = 0; };
struct Srv { Req request; Res response; };

static std::ostringstream trace;

struct Pinger {
  ros::NodeHandle nh;
  ros::Publisher pub;
  ros::Subscriber sub, self_sub;
  ros::Timer timer;
  int id, received = 0, own = 0, ticks = 0;
  Pinger(int id_, const std::string &out, const std::string &in) : id(id_) {
    pub = nh.advertise<std_msgs::UInt16MultiArray>(out, 10);
    sub = nh.subscribe(in, 10, &Pinger::cb, this);
    self_sub = nh.subscribe(out, 10, &Pinger::ownCb, this);
    timer = nh.createTimer(ros::Duration(0.25), &Pinger::tick, this);
  }
  void cb(const std_msgs::UInt16MultiArrayConstPtr &m) {
    received++;
    trace << "t=" << ros::Time::now().toSec() - 1.0e9 << " node" << id << " got " << m->data[0] << "\n";
    if (m->data[0] < 6) {
      std_msgs::UInt16MultiArray r;
      r.data.push_back(m->data[0] + 1);
      pub.publish(r);
    }
  }
  void ownCb(const std_msgs::UInt16MultiArrayConstPtr &) { own++; }
  void tick(const ros::TimerEvent &) { ticks++; }
  bool serve(Req &q, Res &r) {
    r.y = 2 * q.x;
    return true;
  }
};

static int check(bool c, const char *what) {
  if (!c) std::printf("FAILED: %s\n", what);
  return c ? 0 : 1;
}

int main() {
  ros::sim::Node *a = ros::sim::add_node("/a", "n"), *b = ros::sim::add_node("/b", "n");
  a->params["rate"] = 5;               // int, read as int and as double
  a->params["name"] = std::string("alpha");
  int failures = 0;
  Pinger *pa = nullptr, *pb = nullptr;
  auto body = [&](ros::sim::Node *n, int id) {
    ros::sim::enter(n);
    Pinger p(id, id == 0 ? "/ping" : "/pong", id == 0 ? "/pong" : "/ping");
    (id == 0 ? pa : pb) = &p;
    ros::ServiceServer srv;
    if (id == 1) srv = p.nh.advertiseService("/double", &Pinger::serve, &p);
    if (id == 0) {
      int i = 0;
      double d = 0;
      std::string s;
      bool flag = false;
      failures += check(ros::param::get("~rate", i) && i == 5, "int parameter");
      failures += check(ros::param::get("~rate", d) && d == 5.0, "int parameter read as double");
      failures += check(ros::param::get("~name", s) && s == "alpha", "string parameter");
      failures += check(!ros::param::get("~rate", flag) && !ros::param::get("~missing", i), "absent / mistyped parameter");
      failures += check(ros::service::waitForService("/double", ros::Duration(1.0)), "waitForService");
      Srv call;
      call.request.x = 21;
      failures += check(ros::service::call("/double", call) && call.response.y == 42, "service call");
      std_msgs::UInt16MultiArray first;
      first.data.push_back(1);
      p.pub.publish(first);
    }
    ros::Rate rate(100);
    while (ros::ok() && ros::sim::world().now < 1.0) {
      ros::spinOnce();
      if (id == 0 && ros::sim::world().now >= 0.5 && !b->partitioned) {   // cut node b off, then talk to it
        b->partitioned = true;
        std_msgs::UInt16MultiArray lost;
        lost.data.push_back(1);
        p.pub.publish(lost);
      }
      rate.sleep();
    }
    if (id == 0) ros::shutdown();
    while (id == 1 && ros::ok()) rate.sleep();
    trace << "node" << id << " received " << p.received << " own " << p.own << " ticks " << p.ticks << "\n";
    ros::sim::leave();
  };
  std::thread ta(body, a, 0), tb(body, b, 1);
  ta.join();
  tb.join();
  std::printf("%s", trace.str().c_str());
  const std::string t = trace.str();
  // 1 -> b, 2 -> a, ... 6 -> a: three messages each way, then the partitioned message is lost; the publisher hears itself
  failures += check(t.find("node0 received 3 own 4 ticks") != std::string::npos, "node0 counts");
  failures += check(t.find("node1 received 3 own 3 ticks") != std::string::npos, "node1 counts (partition dropped the 4th)");
  // node0 waits 0.02 s for node1's service, then: a reply published by node1 reaches node0 at node0's next 100 Hz spin;
  // node0 runs before node1 within a tick (lower id), so node0's reply is picked up by node1 in the SAME tick
  failures += check(t.find("t=0.02 node1 got 1\nt=0.03 node0 got 2\nt=0.03 node1 got 3\nt=0.04 node0 got 4\n"
                           "t=0.04 node1 got 5\nt=0.05 node0 got 6\n") == 0, "hop timing on the simulated clock");
  failures += check(t.find("ticks 3") != std::string::npos || t.find("ticks 4") != std::string::npos, "0.25 s timer over 1 s");
  std::printf(failures ? "selftest FAILED\n" : "selftest ok\n");
  return failures ? 1 : 0;
}
