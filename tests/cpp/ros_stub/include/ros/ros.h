/*
 * ros/ros.h -- single-process stand-in for the slice of roscpp that dpgo_ros uses, so that the UNMODIFIED wrapper
 * sources of the reference (src/PGOAgentROS.cpp, src/PGOAgentROSNode.cpp, src/PGODatasetPublisherNode.cpp,
 * src/utils.cpp, tests/testUtils.cpp) compile and run against the DPGO:: shim in include/DPGO -- TEST INFRASTRUCTURE.
 *
 * One OS thread per ROS node ("process" of the launch file).  The threads are scheduled COOPERATIVELY on a simulated
 * clock: exactly one node runs at a time; a node gives up the processor only inside ros::Duration::sleep /
 * ros::Rate::sleep, and the clock then jumps to the earliest wake-up time.  Computation takes no simulated time, so a
 * run is deterministic (same event order on every machine and for every DPGO back end) and the wrapper's wall-clock
 * logic (0.5 s start-up sleeps, 3 s timers, 10 s idle before REQUEST_POSE_GRAPH, 15 s time-outs,
 * src/PGOAgentROS.cpp:84-99, 1369-1391, 1499-1563) costs nothing.
 *
 * Topics deliver to every subscriber's callback queue at publish time (also to the publishing node itself, as roscpp
 * does); queues are drained by ros::spinOnce in publish order.  Services run in the caller's thread.
 */
#ifndef ROS_STUB_ROS_H
#define ROS_STUB_ROS_H
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <typeindex>
#include <variant>
#include <vector>

namespace ros {

namespace sim {
typedef std::variant<bool, int, double, std::string> ParamValue;

struct TimerRec {
  double period = 0, next = 0;
  bool active = true;
  std::function<void()> fire;
};
struct Node {
  int id = 0;
  std::string ns, name;                       // "/kimera0/dpgo_ros_node", "agent"
  std::map<std::string, ParamValue> params;   // private ("~") parameters
  std::deque<std::function<void()>> queue;    // callback queue
  std::vector<std::shared_ptr<TimerRec>> timers;
  double wake = 0;
  bool alive = false;
  bool partitioned = false;                   // network partition: topics neither reach nor leave this node
  std::condition_variable cv;                 // signalled when this node gets the processor
};
struct SubRec {
  Node *node;
  std::type_index type;
  std::function<void(const std::shared_ptr<const void> &)> deliver;
  std::shared_ptr<bool> active;
};
struct ServiceRec {
  std::type_index req_type;
  std::function<bool(void *, void *)> call;
};
struct World {
  std::mutex mu;
  double now = 0;
  int current = -1;          // id of the node that holds the processor
  bool shutdown = false;
  std::vector<std::unique_ptr<Node>> nodes;
  std::map<std::string, std::vector<SubRec>> subs;
  std::map<std::string, ServiceRec> services;
  std::map<std::string, unsigned long> published;   // messages per topic (statistics)
  int log_level = 1;         // 0 silent, 1 warnings + errors, 2 everything
  // > 0: pace the simulated clock against the wall clock (simulated seconds per real second).  Needed when something
  // outside the cooperative schedule runs in real time -- the optimisation thread PGOAgent owns in asynchronous mode.
  double realtime_factor = 0;
  // wall-clock instant at which the message now being delivered was published (for monitors that time a run)
  std::chrono::steady_clock::time_point delivering_published_at;
};
inline World &world() {
  static World w;
  return w;
}
inline Node *&self() {
  static thread_local Node *n = nullptr;
  return n;
}

// ---- cooperative scheduler -------------------------------------------------------------------------------------------
inline Node *add_node(const std::string &ns, const std::string &name) {
  World &w = world();
  std::lock_guard<std::mutex> g(w.mu);
  auto n = std::make_unique<Node>();
  n->id = (int)w.nodes.size();
  n->ns = ns;
  n->name = name;
  n->alive = true;
  n->wake = w.now;
  w.nodes.push_back(std::move(n));
  return w.nodes.back().get();
}
// hand the processor to the node with the earliest wake-up time (ties: lowest id); caller holds w.mu
inline void dispatch_locked(World &w) {
  Node *best = nullptr;
  for (auto &n : w.nodes)
    if (n->alive && (!best || n->wake < best->wake)) best = n.get();
  if (!best) {
    w.current = -1;
  } else {
    if (best->wake > w.now && w.realtime_factor > 0)   // every cooperative thread is parked here or on its cv
      std::this_thread::sleep_for(std::chrono::duration<double>((best->wake - w.now) / w.realtime_factor));
    if (best->wake > w.now) w.now = best->wake;
    w.current = best->id;
    best->cv.notify_one();
  }
}
// called first thing by a node's thread: wait for the processor
inline void enter(Node *n) {
  self() = n;
  World &w = world();
  std::unique_lock<std::mutex> lk(w.mu);
  if (w.current < 0) dispatch_locked(w);
  n->cv.wait(lk, [&] { return w.current == n->id; });
}
// called last thing by a node's thread
inline void leave() {
  World &w = world();
  std::unique_lock<std::mutex> lk(w.mu);
  self()->alive = false;
  dispatch_locked(w);
  self() = nullptr;
}
inline void sleep_for(double seconds) {
  Node *n = self();
  World &w = world();
  if (!n) throw std::logic_error("ros stub: sleep outside a node thread");
  std::unique_lock<std::mutex> lk(w.mu);
  n->wake = w.now + (seconds > 0 ? seconds : 0);
  dispatch_locked(w);
  if (w.current != n->id) n->cv.wait(lk, [&] { return w.current == n->id; });
}
inline std::string resolve(const std::string &ns, const std::string &name) {
  if (!name.empty() && name[0] == '/') return name;
  return ns + "/" + name;
}

inline void vlog(int level, const char *tag, const char *fmt, va_list ap) {
  World &w = world();
  if (level > w.log_level) return;
  char buf[2048];
  vsnprintf(buf, sizeof buf, fmt, ap);
  Node *n = self();
  fprintf(stderr, "[%9.3f] [%s] [%s] %s\n", w.now, tag, n ? (n->ns + "/" + n->name).c_str() : "-", buf);
}
inline void log(int level, const char *tag, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
inline void log(int level, const char *tag, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vlog(level, tag, fmt, ap);
  va_end(ap);
}
}  // namespace sim

// ---- time -------------------------------------------------------------------------------------------------------------
class Duration {
 public:
  Duration() = default;
  explicit Duration(double s) : s_(s) {}
  double toSec() const { return s_; }
  bool sleep() const {
    sim::sleep_for(s_);
    return true;
  }
  bool operator>(const Duration &o) const { return s_ > o.s_; }
  bool operator<(const Duration &o) const { return s_ < o.s_; }

 private:
  double s_ = 0;
};
class Time {
 public:
  Time() = default;
  explicit Time(double s) : s_(s) {}
  static Time now() { return Time(1.0e9 + sim::world().now); }   // any fixed epoch
  double toSec() const { return s_; }
  Duration operator-(const Time &o) const { return Duration(s_ - o.s_); }
  Time operator+(const Duration &d) const { return Time(s_ + d.toSec()); }
  bool operator>(const Time &o) const { return s_ > o.s_; }
  bool operator<(const Time &o) const { return s_ < o.s_; }
  bool operator>=(const Time &o) const { return s_ >= o.s_; }
  bool operator<=(const Time &o) const { return s_ <= o.s_; }
  bool operator==(const Time &o) const { return s_ == o.s_; }

 private:
  double s_ = 0;
};
class Rate {
 public:
  explicit Rate(double hz) : period_(1.0 / hz) {}
  bool sleep() {
    sim::sleep_for(period_);
    return true;
  }

 private:
  double period_;
};
struct TimerEvent {
  Time last_expected, last_real, current_expected, current_real;
};

// ---- handles ------------------------------------------------------------------------------------------------------------
class Publisher {
 public:
  Publisher() = default;
  Publisher(std::string topic, std::type_index type) : topic_(std::move(topic)), type_(type) {}
  template <class M>
  void publish(const M &msg) const {
    if (topic_.empty()) return;
    if (std::type_index(typeid(M)) != type_) throw std::logic_error("ros stub: publish() with the wrong message type on " + topic_);
    sim::World &w = sim::world();
    w.published[topic_]++;
    auto it = w.subs.find(topic_);
    if (it == w.subs.end()) return;
    std::shared_ptr<const void> copy = std::make_shared<const M>(msg);
    const auto published_at = std::chrono::steady_clock::now();
    sim::Node *from = sim::self();
    for (auto &s : it->second) {
      if (!*s.active || !s.node->alive) continue;
      if (s.node != from && (s.node->partitioned || (from && from->partitioned))) continue;   // lost on the air
      if (s.type != type_) throw std::logic_error("ros stub: subscriber / publisher type mismatch on " + topic_);
      auto deliver = s.deliver;
      auto active = s.active;
      s.node->queue.push_back([deliver, active, copy, published_at] {
        sim::world().delivering_published_at = published_at;
        if (*active) deliver(copy);
      });
    }
  }
  std::string getTopic() const { return topic_; }

 private:
  std::string topic_;
  std::type_index type_ = std::type_index(typeid(void));
};
class Subscriber {
 public:
  Subscriber() = default;
  explicit Subscriber(std::shared_ptr<bool> active) : active_(std::move(active)) {}
  void shutdown() {
    if (active_) *active_ = false;
  }

 private:
  std::shared_ptr<bool> active_;
};
class Timer {
 public:
  Timer() = default;
  explicit Timer(std::shared_ptr<sim::TimerRec> rec) : rec_(std::move(rec)) {}
  void stop() {
    if (rec_) rec_->active = false;
  }
  void start() {
    if (rec_) rec_->active = true;
  }

 private:
  std::shared_ptr<sim::TimerRec> rec_;
};
class ServiceServer {
 public:
  ServiceServer() = default;
  explicit ServiceServer(std::string name) : name_(std::move(name)) {}

 private:
  std::string name_;
};

class NodeHandle {
 public:
  NodeHandle() : ns_(sim::self() ? sim::self()->ns : std::string()) {}
  explicit NodeHandle(const std::string &ns) {
    sim::Node *n = sim::self();
    if (ns == "~")
      ns_ = n->ns + "/" + n->name;
    else
      ns_ = sim::resolve(n ? n->ns : std::string(), ns);
  }
  const std::string &getNamespace() const { return ns_; }

  template <class M>
  Publisher advertise(const std::string &topic, uint32_t /*queue_size*/, bool /*latch*/ = false) {
    return Publisher(sim::resolve(ns_, topic), std::type_index(typeid(M)));
  }
  template <class M, class T>
  Subscriber subscribe(const std::string &topic, uint32_t /*queue_size*/, void (T::*fp)(const std::shared_ptr<const M> &), T *obj) {
    auto active = std::make_shared<bool>(true);
    sim::SubRec rec{sim::self(), std::type_index(typeid(M)),
                    [fp, obj](const std::shared_ptr<const void> &p) { (obj->*fp)(std::static_pointer_cast<const M>(p)); }, active};
    sim::world().subs[sim::resolve(ns_, topic)].push_back(std::move(rec));
    return Subscriber(active);
  }
  template <class M>
  Subscriber subscribe(const std::string &topic, uint32_t /*queue_size*/, std::function<void(const std::shared_ptr<const M> &)> fn) {
    auto active = std::make_shared<bool>(true);
    sim::SubRec rec{sim::self(), std::type_index(typeid(M)),
                    [fn](const std::shared_ptr<const void> &p) { fn(std::static_pointer_cast<const M>(p)); }, active};
    sim::world().subs[sim::resolve(ns_, topic)].push_back(std::move(rec));
    return Subscriber(active);
  }
  template <class T>
  Timer createTimer(Duration period, void (T::*fp)(const TimerEvent &), T *obj, bool oneshot = false) {
    auto rec = std::make_shared<sim::TimerRec>();
    rec->period = period.toSec();
    rec->next = sim::world().now + rec->period;
    sim::TimerRec *raw = rec.get();
    rec->fire = [fp, obj, raw, oneshot] {
      TimerEvent ev;
      ev.current_real = ev.current_expected = Time::now();
      if (oneshot) raw->active = false;
      (obj->*fp)(ev);
    };
    sim::self()->timers.push_back(rec);
    return Timer(rec);
  }
  template <class T, class Req, class Res>
  ServiceServer advertiseService(const std::string &service, bool (T::*fp)(Req &, Res &), T *obj) {
    const std::string name = sim::resolve(ns_, service);
    sim::ServiceRec rec{std::type_index(typeid(Req)),
                        [fp, obj](void *req, void *res) { return (obj->*fp)(*static_cast<Req *>(req), *static_cast<Res *>(res)); }};
    sim::world().services.erase(name);
    sim::world().services.emplace(name, std::move(rec));
    return ServiceServer(name);
  }
  template <class T>
  bool getParam(const std::string &key, T &out) const;

 private:
  std::string ns_;
};

// ---- parameters (only private "~name" keys are used by dpgo_ros) --------------------------------------------------------
namespace param {
inline const sim::ParamValue *find(const std::string &key) {
  sim::Node *n = sim::self();
  if (!n) return nullptr;
  std::string k = key;
  if (!k.empty() && k[0] == '~') k = k.substr(1);
  auto it = n->params.find(k);
  return it == n->params.end() ? nullptr : &it->second;
}
inline bool get(const std::string &key, int &out) {
  const sim::ParamValue *v = find(key);
  if (!v || !std::holds_alternative<int>(*v)) return false;
  out = std::get<int>(*v);
  return true;
}
inline bool get(const std::string &key, bool &out) {
  const sim::ParamValue *v = find(key);
  if (!v || !std::holds_alternative<bool>(*v)) return false;
  out = std::get<bool>(*v);
  return true;
}
inline bool get(const std::string &key, double &out) {
  const sim::ParamValue *v = find(key);
  if (!v) return false;
  if (std::holds_alternative<double>(*v)) {
    out = std::get<double>(*v);
    return true;
  }
  if (std::holds_alternative<int>(*v)) {   // the parameter server converts int -> double
    out = std::get<int>(*v);
    return true;
  }
  return false;
}
inline bool get(const std::string &key, std::string &out) {
  const sim::ParamValue *v = find(key);
  if (!v || !std::holds_alternative<std::string>(*v)) return false;
  out = std::get<std::string>(*v);
  return true;
}
}  // namespace param
template <class T>
bool NodeHandle::getParam(const std::string &key, T &out) const {
  return param::get(key, out);
}

// ---- services --------------------------------------------------------------------------------------------------------------
namespace service {
inline bool exists(const std::string &name, bool /*print_failure_reason*/ = false) { return sim::world().services.count(name) != 0; }
inline bool waitForService(const std::string &name, Duration timeout = Duration(-1)) {
  double waited = 0;
  while (!exists(name)) {
    if (sim::world().shutdown) return false;
    if (timeout.toSec() >= 0 && waited >= timeout.toSec()) return false;
    sim::sleep_for(0.02);
    waited += 0.02;
  }
  return true;
}
template <class Srv>
bool call(const std::string &name, Srv &srv) {
  auto it = sim::world().services.find(name);
  if (it == sim::world().services.end()) return false;
  if (it->second.req_type != std::type_index(typeid(srv.request))) throw std::logic_error("ros stub: service type mismatch on " + name);
  return it->second.call(&srv.request, &srv.response);
}
}  // namespace service

// ---- process-level API -------------------------------------------------------------------------------------------------------
inline void init(int & /*argc*/, char ** /*argv*/, const std::string &name) {
  // a thread started by a launcher already is a node (sim::add_node + sim::enter); a plain process becomes one here
  if (!sim::self()) sim::enter(sim::add_node("", name));
}
inline bool ok() { return !sim::world().shutdown; }
inline void shutdown() { sim::world().shutdown = true; }
inline void spinOnce() {
  sim::Node *n = sim::self();
  sim::World &w = sim::world();
  for (size_t k = 0; k < n->timers.size(); ++k) {
    auto t = n->timers[k];
    if (!t->active || t->next > w.now) continue;
    t->next = std::max(t->next + t->period, w.now);
    t->fire();
  }
  size_t budget = n->queue.size();   // callbacks available now; what they enqueue is served by the next spin
  while (budget-- > 0 && !n->queue.empty()) {
    auto cb = std::move(n->queue.front());
    n->queue.pop_front();
    cb();
  }
}
inline void spin() {
  while (ok()) {
    spinOnce();
    sim::sleep_for(0.01);
  }
}
}  // namespace ros

#include "ros/console.h"
#endif
