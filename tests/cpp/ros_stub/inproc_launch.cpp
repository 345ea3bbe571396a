/*
 * inproc_launch.cpp -- runs the reference's demo launch files inside ONE process, on the ROS stand-in of this directory:
 * one thread per node of launch/dpgo_demo.launch / launch/dpgo_gnc_demo.launch (the dataset publisher + one
 * dpgo_ros_node per robot), each executing the reference's own, unmodified `main` (src/PGOAgentROSNode.cpp:19,
 * src/PGODatasetPublisherNode.cpp:178; renamed at compile time with -Dmain=...), plus a monitor node that listens to the
 * robots' topics and writes what happened to a JSON file.  TEST INFRASTRUCTURE.
 *
 * The DPGO:: classes under the wrapper are the shim of include/DPGO; which back end they reach is a link-time choice
 * (oracle/Makefile.ref): libdpgo_b200.so (the product, needs a B200) or oracle/abi_on_oracle.cpp (CPU checker).
 *
 * usage: dpgo_ros_inproc_<backend> --robots N (--g2o FILE | --measurements DIR) --out FILE.json
 *            [--preset dpgo_demo|gnc_demo] [--param key=value ...] [--rounds R] [--max-sim-seconds T] [--log 0|1|2]
 *
 * --preset asapp_demo runs the asynchronous demo (launch/asapp_demo.launch): there is no TERMINATE in that mode, so the run
 * is stopped after --run-sim-seconds T and the first / last trajectory every robot published (publish_iterate) are kept;
 * --realtime F paces the simulated clock at F simulated seconds per real second, because the optimisation threads that
 * DPGO::PGOAgent owns in this mode run in real time.
 * --disconnect K@T: at simulated time T robot K drops off the network (its topics neither arrive nor leave) and the other
 * robots' connectivity topics (/kimeraJ/connected_peer_ids, src/PGOAgentROS.cpp:61-63, 910-923) stop listing it.
 * --rounds R: keep the nodes alive until every robot has published R optimised trajectories (the leader starts a new
 * round 10 s after a reset, src/PGOAgentROS.cpp:1381-1385).
 */
#include <ros/ros.h>

#include <dpgo_ros/Command.h>
#include <dpgo_ros/PublicPoses.h>
#include <dpgo_ros/RelativeMeasurementWeights.h>
#include <dpgo_ros/Status.h>
#include <geometry_msgs/PoseArray.h>
#include <std_msgs/UInt16MultiArray.h>

#include "dpgo_b200.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

int dpgo_ros_agent_main(int argc, char **argv);               // src/PGOAgentROSNode.cpp:19
int dpgo_ros_dataset_publisher_main(int argc, char **argv);   // src/PGODatasetPublisherNode.cpp:178

namespace {
using ros::sim::ParamValue;
typedef std::map<std::string, ParamValue> ParamMap;

ParamValue parseValue(const std::string &v) {
  if (v == "true") return true;
  if (v == "false") return false;
  char *end = nullptr;
  const long i = std::strtol(v.c_str(), &end, 10);
  if (end && *end == 0 && !v.empty()) return (int)i;
  const double d = std::strtod(v.c_str(), &end);
  if (end && *end == 0 && !v.empty()) return d;
  return v;
}

// <arg default=...> of launch/PGOAgent.launch:9-50 (what every dpgo_ros_node gets unless the demo overrides it)
ParamMap agentDefaults() {
  return ParamMap{{"dimension", 3}, {"relaxation_rank", 5}, {"asynchronous", false}, {"asynchronous_rate", 10.0},
                  {"verbose", false}, {"RGD_stepsize", 1e-3}, {"RGD_use_preconditioner", true}, {"RTR_iterations", 3},
                  {"RTR_tCG_iterations", 50}, {"RTR_gradnorm_tol", 1e-2}, {"local_initialization_method", std::string("Odometry")},
                  {"update_rule", std::string("Uniform")}, {"multirobot_initialization", true}, {"acceleration", false},
                  {"restart_interval", 50}, {"robust_cost_type", std::string("L2")}, {"GNC_use_probability", true},
                  {"GNC_quantile", 0.9}, {"GNC_barc", 5.0}, {"GNC_mu_step", 2.0}, {"GNC_init_mu", 1e-5},
                  {"robust_opt_num_weight_updates", 4}, {"robust_opt_num_resets", 0}, {"robust_opt_min_convergence_ratio", 0.0},
                  {"robust_opt_inner_iters_per_robot", 10}, {"robust_init_min_inliers", 2}, {"max_iteration_number", 1000},
                  {"relative_change_tolerance", 0.1}, {"publish_iterate", false}, {"visualize_loop_closures", false},
                  {"complete_reset", false}, {"enable_recovery", false}, {"synchronize_measurements", true},
                  {"max_distributed_init_steps", 30}, {"inter_update_sleep_time", 0.0}, {"weight_convergence_threshold", -1.0},
                  {"max_delayed_iterations", 0}, {"timeout_threshold", 15.0}, {"log_output_path", std::string("")}};
}
void applyPreset(const std::string &preset, ParamMap &p) {
  if (preset == "dpgo_demo") {   // launch/dpgo_demo.launch:5-45 (publish_iterate is rviz-only and stays off here)
    p["relative_change_tolerance"] = 0.2;
    p["local_initialization_method"] = std::string("Chordal");
    p["update_rule"] = std::string("RoundRobin");
    p["RTR_iterations"] = 3;
    p["RTR_tCG_iterations"] = 50;
    p["RTR_gradnorm_tol"] = 0.5;
    p["synchronize_measurements"] = true;
  } else if (preset == "gnc_demo") {   // launch/dpgo_gnc_demo.launch:3-47
    p["robust_cost_type"] = std::string("GNC_TLS");
    p["verbose"] = true;
    p["relative_change_tolerance"] = 0.2;
    p["update_rule"] = std::string("RoundRobin");
    p["local_initialization_method"] = std::string("Odometry");
    p["RTR_iterations"] = 3;
    p["RTR_tCG_iterations"] = 50;
    p["RTR_gradnorm_tol"] = 0.5;
    p["GNC_use_probability"] = false;
    p["GNC_barc"] = 3.0;
    p["GNC_mu_step"] = 2.0;
    p["GNC_init_mu"] = 1e-5;
    p["robust_init_min_inliers"] = 3;
    p["robust_opt_num_weight_updates"] = 3;
    p["robust_opt_num_resets"] = 3;
    p["robust_opt_inner_iters_per_robot"] = 50;
    p["synchronize_measurements"] = false;
  } else if (preset == "asapp_demo") {   // launch/asapp_demo.launch:2-37
    p["asynchronous"] = true;
    p["asynchronous_rate"] = 100.0;
    p["verbose"] = true;
    p["publish_iterate"] = true;
    p["RGD_stepsize"] = 0.2;
    p["RGD_use_preconditioner"] = true;
    p["local_initialization_method"] = std::string("Chordal");
    p["synchronize_measurements"] = true;
  } else if (!preset.empty()) {
    std::fprintf(stderr, "unknown preset %s\n", preset.c_str());
    std::exit(2);
  }
}

struct RobotRecord {
  bool have_trajectory = false;
  int trajectories = 0;             // optimised trajectories published so far (one per round)
  bool awaiting = false;            // a TERMINATE command was seen and this robot's result has not arrived yet
  std::vector<double> trajectory;   // latest, n x 7: x y z qx qy qz qw
  std::vector<double> first_trajectory;
  unsigned long trajectory_msgs = 0;
  unsigned max_iteration = 0;       // largest iteration_number this robot reported while INITIALIZED
  unsigned status_msgs = 0;
  double last_relative_change = 0;
  unsigned weights_msgs = 0;
};

class Monitor {
 public:
  Monitor(int robots, int rounds) : rounds_(rounds), rec_(robots) {
    ros::NodeHandle nh;
    for (int k = 0; k < robots; ++k) {
      const std::string prefix = "/kimera" + std::to_string(k) + "/dpgo_ros_node/";
      subs_.push_back(nh.subscribe<geometry_msgs::PoseArray>(prefix + "trajectory", 10, [this, k](const geometry_msgs::PoseArrayConstPtr &m) {
        RobotRecord &r = rec_[k];
        // the 30 s visualisation timer re-publishes the cached result (src/PGOAgentROS.cpp:1387-1390): count a robot's
        // trajectory once per TERMINATE command
        r.have_trajectory = true;
        if (r.awaiting) r.trajectories++;
        r.awaiting = false;
        r.trajectory_msgs++;
        if (r.trajectory_msgs == 1) first_only_ = true;
        r.trajectory.clear();
        for (const auto &p : m->poses)
          for (double v : {p.position.x, p.position.y, p.position.z, p.orientation.x, p.orientation.y, p.orientation.z, p.orientation.w})
            r.trajectory.push_back(v);
        if (first_only_) r.first_trajectory = r.trajectory;
        first_only_ = false;
      }));
      subs_.push_back(nh.subscribe<dpgo_ros::Status>(prefix + "status", 100, [this, k](const dpgo_ros::StatusConstPtr &m) {
        RobotRecord &r = rec_[k];
        r.status_msgs++;
        if (m->state == dpgo_ros::Status::INITIALIZED) {
          if (m->iteration_number > r.max_iteration) r.max_iteration = m->iteration_number;
          r.last_relative_change = m->relative_change;
        }
      }));
      subs_.push_back(nh.subscribe<dpgo_ros::Command>(prefix + "command", 100, [this](const dpgo_ros::CommandConstPtr &m) {
        commands_[m->command]++;
        if (m->command == dpgo_ros::Command::UPDATE && m->executing_iteration > last_update_iteration_) last_update_iteration_ = m->executing_iteration;
        if (m->command == dpgo_ros::Command::UPDATE && m->executing_iteration == 1) {
          round_start_ = ros::sim::world().delivering_published_at;
        }
        if (m->command == dpgo_ros::Command::TERMINATE) {
          round_wall_seconds_.push_back(std::chrono::duration<double>(ros::sim::world().delivering_published_at - round_start_).count());
          // library clock over the same window (publish-time stamps: this node may hear both commands late)
          const double tb = std::chrono::duration<double>(round_start_.time_since_epoch()).count();
          const double te = std::chrono::duration<double>(ros::sim::world().delivering_published_at.time_since_epoch()).count();
          std::string prof((size_t)dpgo_b200_debug_api_profile(tb, te, nullptr, 0, 0), '\0');
          dpgo_b200_debug_api_profile(tb, te, &prof[0], (int)prof.size(), 0);
          round_library_profile_.push_back(prof.c_str());
        }
        if (m->command == dpgo_ros::Command::TERMINATE) {
          terminate_time_ = ros::sim::world().now;
          round_iterations_.push_back(last_update_iteration_);
          last_update_iteration_ = 0;
          for (auto &r : rec_) r.awaiting = true;
        }
      }));
      subs_.push_back(nh.subscribe<dpgo_ros::RelativeMeasurementWeights>(
          prefix + "measurement_weights", 100, [this, k](const dpgo_ros::RelativeMeasurementWeightsConstPtr &) { rec_[k].weights_msgs++; }));
    }
  }
  void expectNothingFrom(int robot) { ignored_ = robot; }
  bool done() const {
    for (size_t k = 0; k < rec_.size(); ++k)
      if ((int)k != ignored_ && rec_[k].trajectories < rounds_) return false;
    return true;
  }
  void write(const std::string &path, bool timed_out) const {
    std::ofstream f(path);
    f.precision(17);
    f << "{\n  \"backend\": \"" << dpgo_b200_version() << "\",\n  \"kernel_launches\": " << dpgo_b200_kernel_launch_count()
      << ",\n  \"timed_out\": " << (timed_out ? "true" : "false") << ",\n  \"sim_seconds\": " << ros::sim::world().now
      << ",\n  \"terminate_sim_seconds\": " << terminate_time_ << ",\n  \"round_iterations\": [";
    for (size_t k = 0; k < round_iterations_.size(); ++k) f << (k ? ", " : "") << round_iterations_[k];
    // real (wall-clock) time between the publication of the first UPDATE command and that of TERMINATE: simulated time
    // costs nothing, so this is the compute + host protocol time of the optimisation itself
    f << "],\n  \"round_wall_seconds\": [";
    for (size_t k = 0; k < round_wall_seconds_.size(); ++k) f << (k ? ", " : "") << round_wall_seconds_[k];
    // seconds inside the DPGO library (every dpgo_b200_* entry point) between the same two events, total and per entry
    // point: round_wall_seconds minus this is the wrapper's own host code + the ROS stand-in
    f << "],\n  \"round_library_seconds\": [";
    for (size_t k = 0; k < round_library_profile_.size(); ++k) {
      double total = 0;
      std::istringstream is(round_library_profile_[k]);
      std::string name;
      double sec;
      long long calls;
      while (is >> name >> sec >> calls)
        if (name[0] != '.') total += sec;  // '.name': a nested section of an entry point already counted
      f << (k ? ", " : "") << total;
    }
    f << "],\n  \"round_library_profile\": [";
    for (size_t k = 0; k < round_library_profile_.size(); ++k) {
      f << (k ? ", {" : "{");
      std::istringstream is(round_library_profile_[k]);
      std::string name;
      double sec;
      long long calls;
      bool firstp = true;
      while (is >> name >> sec >> calls) {
        f << (firstp ? "" : ", ") << "\"" << name << "\": [" << sec << ", " << calls << "]";
        firstp = false;
      }
      f << "}";
    }
    f << "],\n  \"commands\": {";
    bool first = true;
    for (const auto &kv : commands_) {
      f << (first ? "" : ", ") << "\"" << kv.first << "\": " << kv.second;
      first = false;
    }
    f << "},\n  \"messages\": {";
    first = true;
    for (const auto &kv : ros::sim::world().published) {
      f << (first ? "" : ", ") << "\"" << kv.first << "\": " << kv.second;
      first = false;
    }
    f << "},\n  \"robots\": [\n";
    for (size_t k = 0; k < rec_.size(); ++k) {
      const RobotRecord &r = rec_[k];
      f << "    {\"id\": " << k << ", \"max_iteration\": " << r.max_iteration << ", \"status_msgs\": " << r.status_msgs
        << ", \"weights_msgs\": " << r.weights_msgs << ", \"last_relative_change\": " << r.last_relative_change
        << ", \"trajectories\": " << r.trajectories << ", \"trajectory_msgs\": " << r.trajectory_msgs << ", \"first_trajectory\": [";
      for (size_t i = 0; i < r.first_trajectory.size(); ++i) f << (i ? ", " : "") << r.first_trajectory[i];
      f << "], \"trajectory\": [";
      for (size_t i = 0; i < r.trajectory.size(); ++i) f << (i ? ", " : "") << r.trajectory[i];
      f << "]}" << (k + 1 < rec_.size() ? "," : "") << "\n";
    }
    f << "  ]\n}\n";
  }

 private:
  int rounds_;
  int ignored_ = -1;
  bool first_only_ = false;
  std::vector<RobotRecord> rec_;
  std::vector<ros::Subscriber> subs_;
  std::vector<unsigned> round_iterations_;   // iteration number of the last UPDATE command of every finished round
  std::vector<double> round_wall_seconds_;
  std::vector<std::string> round_library_profile_;
  std::chrono::steady_clock::time_point round_start_ = std::chrono::steady_clock::now();
  std::map<int, unsigned long> commands_;
  unsigned last_update_iteration_ = 0;
  double terminate_time_ = -1;
};
}  // namespace

int main(int argc, char **argv) {
  dpgo_b200_debug_api_profile(0, 0, nullptr, 0, 2);  // switch the library's per-entry-point clock on (off by default)
  int robots = 0, rounds = 1;
  std::string g2o, measurements_dir, out = "inproc_result.json", preset;
  double max_sim_seconds = 3600, run_sim_seconds = -1, disconnect_at = -1;
  int disconnect_robot = -1;
  std::vector<std::pair<std::string, std::string>> overrides;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "missing value after %s\n", a.c_str());
        std::exit(2);
      }
      return argv[++i];
    };
    if (a == "--robots") robots = std::atoi(next().c_str());
    else if (a == "--g2o") g2o = next();
    else if (a == "--measurements") measurements_dir = next();
    else if (a == "--out") out = next();
    else if (a == "--preset") preset = next();
    else if (a == "--rounds") rounds = std::atoi(next().c_str());
    else if (a == "--run-sim-seconds") run_sim_seconds = std::atof(next().c_str());
    else if (a == "--disconnect") {
      const std::string v = next();
      const size_t at = v.find('@');
      if (at == std::string::npos) {
        std::fprintf(stderr, "--disconnect expects ROBOT@SIM_SECONDS\n");
        return 2;
      }
      disconnect_robot = std::atoi(v.substr(0, at).c_str());
      disconnect_at = std::atof(v.substr(at + 1).c_str());
    }
    else if (a == "--realtime") ros::sim::world().realtime_factor = std::atof(next().c_str());
    else if (a == "--max-sim-seconds") max_sim_seconds = std::atof(next().c_str());
    else if (a == "--log") ros::sim::world().log_level = std::atoi(next().c_str());
    else if (a == "--param") {
      const std::string kv = next();
      const size_t eq = kv.find('=');
      if (eq == std::string::npos) {
        std::fprintf(stderr, "--param expects key=value\n");
        return 2;
      }
      overrides.emplace_back(kv.substr(0, eq), kv.substr(eq + 1));
    } else {
      std::fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
  }
  if (disconnect_robot >= robots) {
    std::fprintf(stderr, "--disconnect: robot %d does not exist\n", disconnect_robot);
    return 2;
  }
  if (robots <= 0 || (g2o.empty() == measurements_dir.empty())) {
    std::fprintf(stderr, "usage: %s --robots N (--g2o FILE | --measurements DIR) --out FILE.json [--preset dpgo_demo|gnc_demo] "
                         "[--param key=value ...] [--rounds R] [--max-sim-seconds T] [--log 0|1|2]\n", argv[0]);
    return 2;
  }

  // ---- the launch file: nodes and their private parameters
  ros::sim::Node *publisher = ros::sim::add_node("", "dataset_publisher");   // launch/dpgo_demo.launch:13-17
  publisher->params["num_robots"] = robots;
  if (!g2o.empty()) publisher->params["g2o_file"] = g2o;
  else
    for (int k = 0; k < robots; ++k)   // params/robot_measurements.yaml
      publisher->params["robot" + std::to_string(k) + "_measurements"] = measurements_dir + "/robot" + std::to_string(k) + "/measurements.csv";

  ParamMap common = agentDefaults();
  applyPreset(preset, common);
  for (const auto &kv : overrides) common[kv.first] = parseValue(kv.second);
  std::vector<ros::sim::Node *> agents;
  for (int k = 0; k < robots; ++k) {   // <group ns="kimeraK"> + <node ns="dpgo_ros_node" name="agent">, launch/PGOAgent.launch:52
    ros::sim::Node *n = ros::sim::add_node("/kimera" + std::to_string(k) + "/dpgo_ros_node", "agent");
    n->params = common;
    n->params["agent_id"] = k;
    n->params["num_robots"] = robots;
    agents.push_back(n);
  }
  ros::sim::Node *monitor = ros::sim::add_node("", "monitor");

  // ---- one thread per node, each running the reference's own main()
  std::vector<std::thread> threads;
  static bool node_failed = false;   // only ever touched by the thread that holds the processor
  auto run = [](ros::sim::Node *n, int (*entry)(int, char **)) {
    ros::sim::enter(n);
    char name[] = "node";
    char *av[] = {name, nullptr};
    int rc = -100;
    try {
      rc = entry(1, av);
    } catch (const std::exception &e) {
      std::fprintf(stderr, "[%s/%s] uncaught exception: %s\n", n->ns.c_str(), n->name.c_str(), e.what());
      node_failed = true;
      ros::shutdown();
    }
    if (rc != 0 && ros::ok()) {
      std::fprintf(stderr, "[%s/%s] main returned %d\n", n->ns.c_str(), n->name.c_str(), rc);
      node_failed = true;
      ros::shutdown();
    }
    ros::sim::leave();
  };
  threads.emplace_back(run, publisher, &dpgo_ros_dataset_publisher_main);
  for (auto *n : agents) threads.emplace_back(run, n, &dpgo_ros_agent_main);

  ros::sim::enter(monitor);
  bool timed_out = false;
  {
    Monitor mon(robots, rounds);
    if (disconnect_robot >= 0) mon.expectNothingFrom(disconnect_robot);
    while (ros::ok() && !mon.done()) {
      ros::spinOnce();
      if (disconnect_robot >= 0 && disconnect_at >= 0 && ros::sim::world().now >= disconnect_at) {
        ros::NodeHandle nh;
        for (int k = 0; k < robots; ++k) {   // what the network layer of every OTHER robot reports from now on
          if (k == disconnect_robot) continue;
          std_msgs::UInt16MultiArray peers;
          for (int j = 0; j < robots; ++j)
            if (j != disconnect_robot && j != k) peers.data.push_back((uint16_t)j);
          nh.advertise<std_msgs::UInt16MultiArray>("/kimera" + std::to_string(k) + "/connected_peer_ids", 5).publish(peers);
        }
        std_msgs::UInt16MultiArray alone;    // ... and what the lost robot's own reports
        nh.advertise<std_msgs::UInt16MultiArray>("/kimera" + std::to_string(disconnect_robot) + "/connected_peer_ids", 5).publish(alone);
        agents[disconnect_robot]->partitioned = true;
        disconnect_at = -1;
      }
      if (run_sim_seconds > 0 && ros::sim::world().now > run_sim_seconds) break;   // fixed-length run (asynchronous mode)
      if (ros::sim::world().now > max_sim_seconds) {
        timed_out = true;
        break;
      }
      ros::Duration(0.05).sleep();
    }
    const bool ok = mon.done() || run_sim_seconds > 0;
    mon.write(out, timed_out || !ok);
    timed_out = timed_out || !ok;
  }
  ros::shutdown();
  // keep handing the processor on until every node thread has left its main()
  for (;;) {
    bool others = false;
    for (auto &n : ros::sim::world().nodes)
      if (n.get() != monitor && n->alive) others = true;
    if (!others) break;
    ros::sim::sleep_for(0.1);
  }
  ros::sim::leave();
  for (auto &t : threads) t.join();
  if (node_failed) return 3;
  return timed_out ? 1 : 0;
}
