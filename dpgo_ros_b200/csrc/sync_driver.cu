// ROS-free replay of PGOAgentROS's synchronous per-iteration call sequence on N
// agents that live in this process, driven ONLY through the public C ABI with
// HOST buffers -- the way the ROS wrapper would use it:
//   UPDATE handler: non-selected robots iterate(false)          (src/PGOAgentROS.cpp:1185)
//   runOnce:        publishPublicPoses(false/true)              (:109-113, :662-690)
//   callbacks:      updateNeighborPoses / updateAuxNeighborPoses (:1255-1284)
//   selected robot: gate on neighbours' iteration (:136-149), iterate(true) (:160),
//                   publishStatus (:183), leader: shouldTerminate (:208)
// One OS thread per robot stands in for the one-process-per-robot deployment.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/dpgo_b200.h"

namespace {

struct SpinBarrier {
  explicit SpinBarrier(int n) : n_(n) {}
  void wait() {
    const unsigned g = gen_.load(std::memory_order_acquire);
    if (count_.fetch_add(1, std::memory_order_acq_rel) == n_ - 1) {
      count_.store(0, std::memory_order_relaxed);
      gen_.fetch_add(1, std::memory_order_release);
    } else {
      while (gen_.load(std::memory_order_acquire) == g) {
      }
    }
  }
  int n_;
  std::atomic<int> count_{0};
  std::atomic<unsigned> gen_{0};
};

struct Msg {
  int from = 0, to = 0, count = 0;
  std::vector<int> frames;
  std::vector<double> reg, aux;
};

}  // namespace

#pragma GCC visibility push(default)
extern "C" int dpgo_b200_sync_driver_run(dpgo_b200_agent_t *agents, int N, int steps, int accelerated,
                                         double *seconds, long long *payload_bytes, int *terminated_at) {
  if (!agents || N < 1 || steps < 0) return DPGO_B200_ERR_INVALID;
  std::vector<int> r(N);
  // outgoing messages per robot
  std::vector<std::vector<Msg>> out(N);
  std::vector<std::vector<Msg *>> in(N);
  for (int a = 0; a < N; ++a) {
    const int k = dpgo_b200_num_neighbors(agents[a]);
    std::vector<int> nb(k > 0 ? k : 1);
    if (k > 0 && dpgo_b200_get_neighbors(agents[a], nb.data(), k) != 0) return DPGO_B200_ERR_INVALID;
    out[a].resize(k);
    for (int i = 0; i < k; ++i) {
      Msg &m = out[a][i];
      m.from = a;
      m.to = nb[i];
      const int cnt = dpgo_b200_num_shared_poses(agents[a], nb[i]);
      m.frames.resize(cnt > 0 ? cnt : 1);
      m.reg.resize((size_t)(cnt > 0 ? cnt : 1) * 4 * 8);
      m.aux.resize((size_t)(cnt > 0 ? cnt : 1) * 4 * 8);
    }
  }
  for (int a = 0; a < N; ++a)
    for (auto &m : out[a])
      if (m.to >= 0 && m.to < N) in[m.to].push_back(&m);

  std::atomic<int> err{0};
  std::atomic<long long> bytes{0};
  std::atomic<int> term{-1};
  SpinBarrier bar(N);
  const int start_iter = dpgo_b200_iteration_number(agents[0]);
  dpgo_b200_status st0;
  dpgo_b200_get_status(agents[0], &st0);
  const int pose_doubles_hint = 0;
  (void)pose_doubles_hint;

  auto pack = [&](int a) {
    for (auto &m : out[a]) {
      int cnt = 0;
      int rc = dpgo_b200_get_shared_pose_dict(agents[a], m.to, 0, m.frames.data(), m.reg.data(),
                                              (int)m.frames.size(), &cnt);
      if (rc) err.store(rc);
      m.count = cnt;
      if (accelerated) {
        rc = dpgo_b200_get_shared_pose_dict(agents[a], m.to, 1, m.frames.data(), m.aux.data(), (int)m.frames.size(),
                                            &cnt);
        if (rc) err.store(rc);
      }
    }
  };
  auto deliver = [&](int b, int only_from, bool except) {
    for (Msg *m : in[b]) {
      if (except ? (m->from == only_from) : (m->from != only_from)) continue;
      int rc = dpgo_b200_update_neighbor_poses(agents[b], m->from, 0, m->frames.data(), m->reg.data(), m->count);
      if (rc) err.store(rc);
      if (accelerated) {
        rc = dpgo_b200_update_neighbor_poses(agents[b], m->from, 1, m->frames.data(), m->aux.data(), m->count);
        if (rc) err.store(rc);
      }
    }
  };

  const bool prof = getenv("DPGO_B200_DRIVER_PROFILE") != nullptr;
  double tp[6] = {0, 0, 0, 0, 0, 0};
  auto now = [] { return std::chrono::high_resolution_clock::now(); };
  auto since = [](std::chrono::high_resolution_clock::time_point t) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t).count();
  };
  auto worker = [&](int a) {
    for (int s = 0; s < steps; ++s) {
      const int sel = (start_iter + s) % N;
      auto t = now();
      if (accelerated) {
        if (a != sel) {
          if (dpgo_b200_iterate(agents[a], 0)) err.store(-1);
          if (prof && a == (sel + 1) % N) tp[0] += since(t);
          pack(a);
        }
        bar.wait();
        if (prof && a == 0) { tp[1] += since(t); t = now(); }
        deliver(a, sel, /*except=*/true);  // everything except the selected robot's (not sent yet)
        bar.wait();
        if (prof && a == 0) { tp[2] += since(t); t = now(); }
      } else if (a != sel) {
        if (dpgo_b200_iterate(agents[a], 0)) err.store(-1);
      }
      if (a == sel) {
        auto t2 = now();
        if (dpgo_b200_iterate(agents[a], 1)) err.store(-1);
        if (prof) tp[3] += since(t2);
        pack(a);
      }
      bar.wait();
      if (prof && a == 0) { tp[4] += since(t); t = now(); }
      deliver(a, sel, /*except=*/false);  // only the selected robot's poses
      if (a == 0) {
        // publishStatus: the leader hears everyone; shouldTerminate on the leader's own turn
        for (int b = 1; b < N; ++b) {
          dpgo_b200_status st;
          dpgo_b200_get_status(agents[b], &st);
          dpgo_b200_set_neighbor_status(agents[0], &st);
        }
        if (sel == 0 && term.load() < 0 && dpgo_b200_should_terminate(agents[0]) == 1) term.store(s + 1);
      }
      bar.wait();
      if (err.load()) return;
    }
  };

  // payload accounting: poses that crossed the host per step
  long long per_cycle = 0;
  for (int a = 0; a < N; ++a)
    for (auto &m : out[a]) per_cycle += (long long)m.frames.size();
  (void)per_cycle;

  const auto t0 = std::chrono::high_resolution_clock::now();
  std::vector<std::thread> th;
  for (int a = 1; a < N; ++a) th.emplace_back(worker, a);
  worker(0);
  for (auto &t : th) t.join();
  const double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  if (prof)
    fprintf(stderr, "[sync_driver] per step (us): iterate(false) %.1f | phase1 %.1f | deliver %.1f | iterate(true) %.1f | phase3 %.1f | total %.1f\n",
            tp[0] / steps * 1e6, tp[1] / steps * 1e6, tp[2] / steps * 1e6, tp[3] / steps * 1e6, tp[4] / steps * 1e6,
            dt / steps * 1e6);
  if (seconds) *seconds = dt;
  if (payload_bytes) *payload_bytes = bytes.load();
  if (terminated_at) *terminated_at = term.load();
  return err.load();
}
#pragma GCC visibility pop
