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

// ---------------------------------------------------------------------------------------------------
// The same replay with the robots spread over several PROCESSES of one node (one per GPU): the mailboxes, the
// barrier and the status board live in a POSIX shared-memory segment, standing in for the TCPROS transport between
// the per-robot processes of the reference deployment (launch/dpgo_demo.launch:21-123).  Every process calls
// dpgo_b200_sync_driver_run_shm with its own robots; the protocol (and the arithmetic) is the one above.
// ---------------------------------------------------------------------------------------------------
namespace {

constexpr int kShmMaxRobots = 64;
constexpr int kShmPoseDoubles = 4 * 8;  // r <= 8

struct ShmHeader {
  std::atomic<int> count;
  std::atomic<unsigned> gen;
  std::atomic<int> err, term;
  int cap;  // poses per mailbox
  dpgo_b200_status status[2][kShmMaxRobots];   // by step parity, like the mailboxes
};
struct ShmView {
  ShmHeader *h;
  unsigned char *boxes;
  size_t box_bytes;
  int N;
  // mailbox a -> b of step parity p: [int count | int frames[cap] | double reg[cap * 32] | double aux[cap * 32]]
  // (two sets: a robot may already post the next iteration's poses while a slower one still reads this iteration's)
  int p = 0;
  unsigned char *box(int a, int b) const { return boxes + (((size_t)p * N + a) * N + b) * box_bytes; }
  int *count(int a, int b) const { return reinterpret_cast<int *>(box(a, b)); }
  int *frames(int a, int b) const { return reinterpret_cast<int *>(box(a, b)) + 2; }
  double *reg(int a, int b) const { return reinterpret_cast<double *>(box(a, b) + 8 + ((size_t)h->cap * 4 + 7) / 8 * 8); }
  double *aux(int a, int b) const { return reg(a, b) + (size_t)h->cap * kShmPoseDoubles; }
  ShmView at(int parity) const {
    ShmView v = *this;
    v.p = parity;
    return v;
  }
};
size_t shm_box_bytes(int cap) { return 8 + ((size_t)cap * 4 + 7) / 8 * 8 + (size_t)2 * cap * kShmPoseDoubles * sizeof(double); }

void shm_barrier(ShmHeader *h, int n) {
  const unsigned g = h->gen.load(std::memory_order_acquire);
  if (h->count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
    h->count.store(0, std::memory_order_relaxed);
    h->gen.fetch_add(1, std::memory_order_release);
  } else {
    while (h->gen.load(std::memory_order_acquire) == g) {
    }
  }
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" size_t dpgo_b200_sync_driver_shm_bytes(int num_robots, int max_shared_poses) {
  return sizeof(ShmHeader) + (size_t)2 * num_robots * num_robots * shm_box_bytes(max_shared_poses);
}

extern "C" int dpgo_b200_sync_driver_run_shm(dpgo_b200_agent_t *agents, const int *robot_ids, int num_local,
                                             int num_robots, void *shm, int max_shared_poses, int steps,
                                             int accelerated, int start_iter, double *seconds, int *terminated_at) {
  if (!agents || !robot_ids || !shm || num_local < 1 || num_robots < num_local || num_robots > kShmMaxRobots || steps < 0)
    return DPGO_B200_ERR_INVALID;
  const int N = num_robots;
  ShmView V;
  V.h = static_cast<ShmHeader *>(shm);
  V.boxes = static_cast<unsigned char *>(shm) + sizeof(ShmHeader);
  V.box_bytes = shm_box_bytes(max_shared_poses);
  V.N = N;
  V.h->cap = max_shared_poses;  // (every rank writes the same value)
  std::vector<std::vector<int>> nbrs(num_local);
  for (int i = 0; i < num_local; ++i) {
    const int k = dpgo_b200_num_neighbors(agents[i]);
    nbrs[i].resize(k > 0 ? k : 0);
    if (k > 0 && dpgo_b200_get_neighbors(agents[i], nbrs[i].data(), k) != 0) return DPGO_B200_ERR_INVALID;
    for (int b : nbrs[i])
      if (dpgo_b200_num_shared_poses(agents[i], b) > max_shared_poses) return DPGO_B200_ERR_INVALID;
  }
  ShmHeader *H = V.h;
  auto fail = [&](int rc) { H->err.store(rc); };
  auto pack = [&](int i, const ShmView &V) {  // publishPublicPoses of local robot i
    const int a = robot_ids[i];
    for (int b : nbrs[i]) {
      int cnt = 0;
      int rc = dpgo_b200_get_shared_pose_dict(agents[i], b, 0, V.frames(a, b), V.reg(a, b), max_shared_poses, &cnt);
      if (rc) fail(rc);
      *V.count(a, b) = cnt;
      if (accelerated) {
        rc = dpgo_b200_get_shared_pose_dict(agents[i], b, 1, V.frames(a, b), V.aux(a, b), max_shared_poses, &cnt);
        if (rc) fail(rc);
      }
    }
  };
  auto deliver = [&](int i, int sel, bool except, const ShmView &V) {  // publicPosesCallback of local robot i
    const int b = robot_ids[i];
    for (int a : nbrs[i]) {
      if (except ? (a == sel) : (a != sel)) continue;
      const int cnt = *V.count(a, b);
      int rc = dpgo_b200_update_neighbor_poses(agents[i], a, 0, V.frames(a, b), V.reg(a, b), cnt);
      if (rc) fail(rc);
      if (accelerated) {
        rc = dpgo_b200_update_neighbor_poses(agents[i], a, 1, V.frames(a, b), V.aux(a, b), cnt);
        if (rc) fail(rc);
      }
    }
  };
  int my_term = -1;
  if (steps == 0) {
    // INITIALIZE (:1099-1101): everybody publishes once, everybody hears everybody
    std::vector<std::thread> th0;
    auto once = [&](int i) {
      pack(i, V);
      shm_barrier(H, N);
      deliver(i, -1, /*except=*/true, V);
      shm_barrier(H, N);
    };
    for (int i = 1; i < num_local; ++i) th0.emplace_back(once, i);
    once(0);
    for (auto &t : th0) t.join();
    if (seconds) *seconds = 0;
    if (terminated_at) *terminated_at = -1;
    return H->err.load();
  }
  auto worker = [&](int i) {
    const int a = robot_ids[i];
    for (int s = 0; s < steps; ++s) {
      const int sel = (start_iter + s) % N;
      const int p = s & 1;
      const ShmView Vp = V.at(p);
      if (accelerated) {
        if (a != sel) {
          if (dpgo_b200_iterate(agents[i], 0)) fail(-1);
          pack(i, Vp);
        }
        shm_barrier(H, N);
        deliver(i, sel, /*except=*/true, Vp);
        shm_barrier(H, N);
      } else if (a != sel) {
        if (dpgo_b200_iterate(agents[i], 0)) fail(-1);
      }
      if (a == sel) {
        if (dpgo_b200_iterate(agents[i], 1)) fail(-1);
        pack(i, Vp);
      }
      dpgo_b200_get_status(agents[i], &H->status[p][a]);  // publishStatus
      shm_barrier(H, N);
      deliver(i, sel, /*except=*/false, Vp);
      if (a == 0) {
        for (int b = 1; b < N; ++b) dpgo_b200_set_neighbor_status(agents[i], &H->status[p][b]);
        if (sel == 0 && H->term.load() < 0 && dpgo_b200_should_terminate(agents[i]) == 1) H->term.store(s + 1);
      }
      // no barrier at the end of a step (mailboxes and statuses are double-buffered by the step's parity; see the
      // in-process driver below).  An error does not end the loop early: without that barrier the threads would not
      // agree on the step at which to leave; the calls of a failed agent return at once, and the code is reported
    }
    shm_barrier(H, N);   // the call ends together: the next call restarts the parity at 0
  };
  const auto t0 = std::chrono::high_resolution_clock::now();
  std::vector<std::thread> th;
  for (int i = 1; i < num_local; ++i) th.emplace_back(worker, i);
  worker(0);
  for (auto &t : th) t.join();
  const double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  my_term = H->term.load();
  if (seconds) *seconds = dt;
  if (terminated_at) *terminated_at = my_term;
  return H->err.load();
}
#pragma GCC visibility pop

#pragma GCC visibility push(default)
extern "C" int dpgo_b200_sync_driver_run(dpgo_b200_agent_t *agents, int N, int steps, int accelerated,
                                         double *seconds, long long *payload_bytes, int *terminated_at) {
  if (!agents || N < 1 || steps < 0) return DPGO_B200_ERR_INVALID;
  // outgoing messages per robot, double-buffered by the parity of the step: a robot may already pack the next
  // iteration's poses while a slower one still reads this iteration's (messages in flight, as over TCPROS) -- which
  // is what lets the step run on three barriers instead of four
  std::vector<std::vector<Msg>> out[2];
  std::vector<std::vector<Msg *>> in[2];
  for (int p = 0; p < 2; ++p) {
    out[p].resize(N);
    in[p].resize(N);
    for (int a = 0; a < N; ++a) {
      const int k = dpgo_b200_num_neighbors(agents[a]);
      std::vector<int> nb(k > 0 ? k : 1);
      if (k > 0 && dpgo_b200_get_neighbors(agents[a], nb.data(), k) != 0) return DPGO_B200_ERR_INVALID;
      out[p][a].resize(k);
      for (int i = 0; i < k; ++i) {
        Msg &m = out[p][a][i];
        m.from = a;
        m.to = nb[i];
        const int cnt = dpgo_b200_num_shared_poses(agents[a], nb[i]);
        m.frames.resize(cnt > 0 ? cnt : 1);
        m.reg.resize((size_t)(cnt > 0 ? cnt : 1) * 4 * 8);
        m.aux.resize((size_t)(cnt > 0 ? cnt : 1) * 4 * 8);
      }
    }
    for (int a = 0; a < N; ++a)
      for (auto &m : out[p][a])
        if (m.to >= 0 && m.to < N) in[p][m.to].push_back(&m);
  }

  std::atomic<int> err{0};
  std::atomic<long long> bytes{0};
  std::atomic<int> term{-1};
  SpinBarrier bar(N);
  const int start_iter = dpgo_b200_iteration_number(agents[0]);
  // publishStatus: every robot posts its own, the leader reads them (by step parity, like the messages: without
  // acceleration a step has a single barrier, and a fast robot posts its next status while the leader still reads)
  std::vector<dpgo_b200_status> status_board[2] = {std::vector<dpgo_b200_status>(N), std::vector<dpgo_b200_status>(N)};

  auto pack = [&](int a, int p) {
    for (auto &m : out[p][a]) {
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
  auto deliver = [&](int b, int only_from, bool except, int p) {
    for (Msg *m : in[p][b]) {
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
      const int p = s & 1;
      auto t = now();
      if (accelerated) {
        if (a != sel) {
          if (dpgo_b200_iterate(agents[a], 0)) err.store(-1);
          if (prof && a == (sel + 1) % N) tp[0] += since(t);
          pack(a, p);
        }
        bar.wait();
        if (prof && a == 0) { tp[1] += since(t); t = now(); }
        deliver(a, sel, /*except=*/true, p);  // everything except the selected robot's (not sent yet)
        bar.wait();
        if (prof && a == 0) { tp[2] += since(t); t = now(); }
      } else if (a != sel) {
        if (dpgo_b200_iterate(agents[a], 0)) err.store(-1);
      }
      if (a == sel) {
        auto t2 = now();
        if (dpgo_b200_iterate(agents[a], 1)) err.store(-1);
        if (prof) tp[3] += since(t2);
        pack(a, p);
      }
      dpgo_b200_get_status(agents[a], &status_board[p][a]);   // publishStatus
      bar.wait();
      if (prof && a == 0) { tp[4] += since(t); t = now(); }
      deliver(a, sel, /*except=*/false, p);  // only the selected robot's poses
      if (a == 0) {
        // the leader hears everyone; shouldTerminate on the leader's own turn
        for (int b = 1; b < N; ++b) dpgo_b200_set_neighbor_status(agents[0], &status_board[p][b]);
        if (sel == 0 && term.load() < 0 && dpgo_b200_should_terminate(agents[0]) == 1) term.store(s + 1);
      }
      // no barrier here: the next step's first one orders this step's deliveries in front of everything that depends
      // on them (a robot iterates and delivers to itself on its own thread; messages are double-buffered; a status
      // is posted again only behind two barriers of the next step).  An error does not end the loop early -- the
      // threads would not agree on the step at which to leave -- it is reported when the call returns.
    }
  };

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
