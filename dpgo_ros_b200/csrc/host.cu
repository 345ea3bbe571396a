// Host side of the B200 RBCD path (see host.hpp).  No arithmetic of the hot
// path happens here: this file keeps the pose graph, lays the data out for the
// kernels, and sequences launches.  There is no CPU fallback.
#include "host.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace dpgo {

[[noreturn]] void fail(int code, const std::string &msg) { throw Error{code, msg}; }
void cuda_check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) fail(DPGO_B200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

static inline size_t roundup32(size_t x) { return (x + 31) / 32 * 32; }

// cudaSetDevice takes a driver lock even when nothing changes; cudaGetDevice is a thread-local read
static inline cudaError_t use_device(int device) {
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return cudaSuccess;
  return cudaSetDevice(device);
}

// agents alive per device: an armed launch (Agent::maybe_arm) keeps every SM of its device busy while it waits for
// the doorbell, which is only acceptable when nobody else can want that device
static std::atomic<int> g_agents_on_device[64];

// ============================================================================
// Agent
// ============================================================================
Agent::Agent(int id_, const dpgo_b200_params &p, int device_) : id(id_), P(p), device(device_), r(p.r) {
  if (p.d != 3) fail(DPGO_B200_ERR_INVALID, "only d = 3 is supported (src/PGOAgentROSNode.cpp:63)");
  if (p.r < 3 || p.r > 8) fail(DPGO_B200_ERR_INVALID, "relaxation rank r must satisfy 3 <= r <= 8");
  if (p.num_robots < 1 || p.num_robots > kMaxRobots) fail(DPGO_B200_ERR_INVALID, "num_robots out of range");
  if (id_ < 0 || id_ >= p.num_robots) fail(DPGO_B200_ERR_INVALID, "agent id out of range");
  if (p.cost_type != 0 && p.cost_type != 5)
    fail(DPGO_B200_ERR_INVALID, "only L2 and GNC_TLS robust costs are in scope (SURVEY §2 #4)");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= device_ || device_ < 0)
    fail(DPGO_B200_ERR_CUDA, "no usable CUDA device (the RBCD path has no CPU fallback)");
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  mu = p.gnc_init_mu;
  status.agent_id = id;
  own.reset(new Team(device_));
  own->add(this);
  ++g_agents_on_device[device_ & 63];
}

Agent::~Agent() {
  --g_agents_on_device[device & 63];
  if (armed) {
    try {
      disarm();
    } catch (...) {
    }
  }
  if (team && team != own.get()) team->remove(this);
  if (own) own->agents.clear();
  team = nullptr;
  if (pstream) {
    cudaStreamSynchronize(pstream);
    cudaStreamDestroy(pstream);
    cudaEventDestroy(pevent);
    cudaEventDestroy(pevent_in);
  }
  free_pinned();
}

void Agent::free_pinned() {
  if (h_inbox) cudaFreeHost(h_inbox);
  h_inbox = nullptr;
}

void Agent::add_measurement(const Meas &m) {
  if (m.r1 != id && m.r2 != id) return;  // irrelevant measurement, src/PGOAgentROS.cpp:273
  auto key = std::make_pair(std::make_pair(m.r1, m.p1), std::make_pair(m.r2, m.p2));
  if (have.count(key)) return;  // hasMeasurement, :276
  if (m.r1 < 0 || m.r2 < 0 || m.r1 >= P.num_robots || m.r2 >= P.num_robots || m.p1 < 0 || m.p2 < 0)
    fail(DPGO_B200_ERR_INVALID, "measurement index out of range");
  if (m.r1 == id && m.r2 == id) {
    if (m.p1 + 1 == m.p2) {
      have[key] = {0, (int)odom.size()};
      odom.push_back(m);
    } else {
      have[key] = {1, (int)plc.size()};
      plc.push_back(m);
    }
    n = std::max(n, std::max(m.p1, m.p2) + 1);
  } else {
    have[key] = {2, (int)slc.size()};
    slc.push_back(m);
    if (m.r1 == id) {
      n = std::max(n, m.p1 + 1);
      nbrs.insert(m.r2);
    } else {
      n = std::max(n, m.p2 + 1);
      nbrs.insert(m.r1);
    }
  }
  if (armed) disarm();
  if (la_used > 0 && !structure_dirty) materialize_lookahead();
  drop_lookahead();
  structure_dirty = values_dirty = precon_dirty = wiring_dirty = lc_dirty = weights_host_dirty = true;
  resid_valid = false;
  pub_frames_cache.clear();
  if (team) team->team_dirty = true;
}

Meas *Agent::find_measurement(int r1, int p1, int r2, int p2) {
  auto it = have.find({{r1, p1}, {r2, p2}});
  if (it == have.end()) return nullptr;
  return &(it->second.first == 0 ? odom : (it->second.first == 1 ? plc : slc))[it->second.second];
}
int Agent::measurement_index(int r1, int p1, int r2, int p2) const {
  auto it = have.find({{r1, p1}, {r2, p2}});
  if (it == have.end()) return -1;
  const int base = it->second.first == 0 ? 0 : (it->second.first == 1 ? (int)odom.size() : (int)(odom.size() + plc.size()));
  return base + it->second.second;
}

const std::vector<int> &Agent::my_public_frames(int nbr) const {
  auto it = pub_frames_cache.find(nbr);
  if (it != pub_frames_cache.end()) return it->second;
  std::set<int> s;
  for (const auto &m : slc) {
    if (m.r1 == id && m.r2 == nbr) s.insert(m.p1);
    if (m.r2 == id && m.r1 == nbr) s.insert(m.p2);
  }
  return pub_frames_cache.emplace(nbr, std::vector<int>(s.begin(), s.end())).first->second;
}

void Agent::set_lifting_matrix(const double *Y) {
  std::memcpy(ylift, Y, sizeof(double) * r * 3);
  have_lift = true;
}

void Agent::initialize(const double *T) {
  if (n == 0) fail(DPGO_B200_ERR_STATE, "initialize: empty pose graph");
  Tlocal.assign((size_t)12 * n, 0.0);
  if (T) {
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 4; ++c) Tlocal[(size_t)i * 12 + c * 3 + a] = T[(size_t)i * 12 + a * 4 + c];
  } else {
    // Odometry chain from identity (local_initialization_method Odometry, src/PGOAgentROSNode.cpp:106-108)
    std::vector<const Meas *> by_src(n, nullptr);
    for (const auto &m : odom) by_src[m.p1] = &m;
    for (int c = 0; c < 3; ++c) Tlocal[c * 3 + c] = 1.0;
    for (int i = 0; i + 1 < n; ++i) {
      const Meas *m = by_src[i];
      if (!m) fail(DPGO_B200_ERR_MISSING, "initialize: missing odometry edge");
      const double *Ti = &Tlocal[(size_t)i * 12];
      double *Tn = &Tlocal[(size_t)(i + 1) * 12];
      for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += Ti[k * 3 + a] * m->R[c * 3 + k];
          Tn[c * 3 + a] = s;
        }
      for (int a = 0; a < 3; ++a) {
        double s = Ti[9 + a];
        for (int k = 0; k < 3; ++k) s += Ti[k * 3 + a] * m->t[k];
        Tn[9 + a] = s;
      }
    }
  }
  state = 1;
}

// Chordal initialisation on the device (demo default, launch/dpgo_demo.launch:9): over the robot's
// own odometry + private loop closures with pose 0 fixed to the identity,
//   rotations   : rows x_i of R_i minimise sum k |x_j - x_i R_ij|^2      (linear), then the polar factor per pose,
//   translations: sum t |p_j - p_i - R_i t_ij|^2                         (linear).
// Both systems share one block matrix (blocks [-k R_ij 0; 0 -t], diagonal [sum k I3 0; 0 sum t]) whose dense inverse
// comes from the preconditioner machinery (dense_inverse.cu); the host only assembles the sparse blocks and the two
// right-hand sides.  Same arithmetic as oracle Agent::initializeChordal.
void Agent::initialize_chordal() {
  ProfSection prof_(".chordal");
  if (n == 0) fail(DPGO_B200_ERR_STATE, "initialize: empty pose graph");
  cuda_check(use_device(device), "cudaSetDevice");
  Tlocal.assign((size_t)12 * n, 0.0);
  for (int c = 0; c < 3; ++c) Tlocal[c * 3 + c] = 1.0;
  const int N = n - 1;
  if (N > 0) {
    std::vector<std::map<int, std::array<double, 16>>> cols(N);
    auto blk = [&](int i, int j) -> std::array<double, 16> & {
      auto it = cols[j].find(i);
      if (it == cols[j].end()) it = cols[j].emplace(i, std::array<double, 16>{}).first;
      return it->second;
    };
    std::vector<double> B1((size_t)3 * 4 * N, 0.0);  // 3 x 4N column-major
    std::vector<const Meas *> edges;
    for (const auto &m : odom) edges.push_back(&m);
    for (const auto &m : plc) edges.push_back(&m);
    for (const Meas *m : edges) {
      const double k = m->weight * m->kappa, t = m->weight * m->tau;
      const int i = m->p1 - 1, j = m->p2 - 1;
      if (i >= 0) {
        auto &D = blk(i, i);
        D[0] += k; D[5] += k; D[10] += k; D[15] += t;
      }
      if (j >= 0) {
        auto &D = blk(j, j);
        D[0] += k; D[5] += k; D[10] += k; D[15] += t;
      }
      if (i >= 0 && j >= 0) {
        auto &U = blk(i, j), &L = blk(j, i);
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < 3; ++a) {
            U[c * 4 + a] -= k * m->R[c * 3 + a];
            L[c * 4 + a] -= k * m->R[a * 3 + c];
          }
        U[15] -= t;
        L[15] -= t;
      } else if (i < 0 && j >= 0) {
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < 3; ++a) B1[((size_t)4 * j + c) * 3 + a] += k * m->R[c * 3 + a];
      } else if (j < 0 && i >= 0) {
        for (int c = 0; c < 3; ++c)
          for (int a = 0; a < 3; ++a) B1[((size_t)4 * i + c) * 3 + a] += k * m->R[a * 3 + c];
      }
    }
    std::vector<int> rowptr(N + 1, 0), colidx;
    std::vector<double> vals;
    for (int j = 0; j < N; ++j) {
      for (const auto &kv : cols[j]) {
        colidx.push_back(kv.first);
        vals.insert(vals.end(), kv.second.begin(), kv.second.end());
      }
      rowptr[j + 1] = (int)colidx.size();
    }
    const size_t npad = roundup32((size_t)4 * N);
    DevBuf<int> d_rp, d_ci, info;
    DevBuf<double> d_v, P, work, dB, dZ;
    d_rp.upload(rowptr);
    d_ci.upload(colidx);
    d_v.upload(vals);
    P.alloc(npad * npad);
    work.alloc(npad * npad, false);
    info.alloc(1);
    cuda_check(launch_scatter_blocks(P.p, npad, d_rp.p, d_ci.p, d_v.p, N, 0.0, (int)npad, 0), "scatter_blocks");
    cuda_check(spd_inverse(P.p, work.p, (int)npad, info.p, 0), "spd_inverse");
    int h_info = 0;
    cuda_check(cudaMemcpy(&h_info, info.p, sizeof(int), cudaMemcpyDeviceToHost), "D2H info");
    if (h_info != 0) fail(DPGO_B200_ERR_NUMERIC, "Chordal initialisation: the pose graph is not connected to pose 0");
    // stage 1: rotations
    dB.upload(B1);
    dZ.alloc((size_t)3 * 4 * N);
    cuda_check(launch_rows_times_sym(dB.p, P.p, npad, 3, 4 * N, dZ.p, 0), "rows_times_sym");
    cuda_check(launch_manifold_op(0, 3, N, dZ.p, nullptr, dB.p, std::max(1, std::min(148, (N + 31) / 32)), 0),
               "polar factor");
    std::vector<double> Rs((size_t)12 * N);
    cuda_check(cudaMemcpy(Rs.data(), dB.p, Rs.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H rotations");
    for (int i = 0; i < N; ++i)
      for (int c = 0; c < 3; ++c)
        for (int a = 0; a < 3; ++a) Tlocal[(size_t)(i + 1) * 12 + c * 3 + a] = Rs[((size_t)4 * i + c) * 3 + a];
    // stage 2: translations
    std::vector<double> B2((size_t)3 * 4 * N, 0.0);
    for (const Meas *m : edges) {
      const double t = m->weight * m->tau;
      const int i = m->p1 - 1, j = m->p2 - 1;
      const double *Ri = &Tlocal[(size_t)m->p1 * 12];
      for (int a = 0; a < 3; ++a) {
        double sum = 0;
        for (int kk = 0; kk < 3; ++kk) sum += Ri[kk * 3 + a] * m->t[kk];
        const double v = t * sum;
        if (j >= 0) B2[((size_t)4 * j + 3) * 3 + a] += v;
        if (i >= 0) B2[((size_t)4 * i + 3) * 3 + a] -= v;
      }
    }
    dB.upload(B2);
    cuda_check(launch_rows_times_sym(dB.p, P.p, npad, 3, 4 * N, dZ.p, 0), "rows_times_sym");
    std::vector<double> Ts((size_t)12 * N);
    cuda_check(cudaMemcpy(Ts.data(), dZ.p, Ts.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H translations");
    for (int i = 0; i < N; ++i)
      for (int a = 0; a < 3; ++a) Tlocal[(size_t)(i + 1) * 12 + 9 + a] = Ts[((size_t)4 * i + 3) * 3 + a];
  }
  drop_lookahead();
  state = 1;
}

void Agent::initialize_in_global_frame(const double *Tw_rm) {
  if (state == 0) fail(DPGO_B200_ERR_STATE, "initializeInGlobalFrame before initialize");
  if (!have_lift) fail(DPGO_B200_ERR_STATE, "initializeInGlobalFrame: lifting matrix not set");
  cuda_check(use_device(device), "cudaSetDevice");
  double Tw[12];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 4; ++c) Tw[c * 3 + a] = Tw_rm[a * 4 + c];
  std::vector<double> X((size_t)r * 4 * n);
  for (int i = 0; i < n; ++i) {
    double Tg[12];
    for (int c = 0; c < 4; ++c)
      for (int a = 0; a < 3; ++a) {
        double s = (c == 3) ? Tw[9 + a] : 0.0;
        for (int k = 0; k < 3; ++k) s += Tw[k * 3 + a] * Tlocal[(size_t)i * 12 + c * 3 + k];
        Tg[c * 3 + a] = s;
      }
    for (int c = 0; c < 4; ++c)
      for (int a = 0; a < r; ++a) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += ylift[k * r + a] * Tg[c * 3 + k];
        X[((size_t)i * 4 + c) * r + a] = s;
      }
  }
  if (armed) disarm();
  drop_lookahead();
  resid_valid = false;
  dX.upload(X);
  dXinit.upload(X);
  dY.upload(X);
  dV.upload(X);
  state = 2;
  status.state = 2;
  team->ctl.gamma = team->ctl.alpha = 0;
  team->gamma_state = 0;
  team->team_dirty = true;
  outbox_stale = true;
  warm();
}

// Allocate, assemble and wire everything but the dense inverse as soon as the robot is initialised.  Device and pinned
// allocations are implicit synchronisation points between streams: with every robot's allocations made here, in front
// of the round, the inverses that the robots of one GPU start at their first iterate() -- true or false -- run side by
// side instead of one after the other (config 3: four 5024 x 5024 inverses, most of a 9-iteration run).  Best effort:
// whatever is not ready yet (measurements still to come, ...) is left to the first iterate() as before.
void Agent::warm() {
  if (!team || team != own.get() || state != 2 || n == 0) return;
  try {
    team->precond_mode = 0;
    team->prepare(false, true);
  } catch (...) {
  }
  team->precond_mode = 2;
}

void Agent::reset() {
  if (armed) disarm();
  drop_lookahead();
  resid_valid = false;
  instance++;
  iter = 0;
  state = 0;
  status = dpgo_b200_status{};
  status.agent_id = id;
  status.instance_number = instance;
  team_status.clear();
  inactive_robots.clear();
  std::fill(inbox_valid_reg.begin(), inbox_valid_reg.end(), 0);
  std::fill(inbox_valid_aux.begin(), inbox_valid_aux.end(), 0);
  weight_update_count = 0;
  robust_inner_iter = 0;
  mu = P.gnc_init_mu;
  opt = dpgo_b200_opt_result{};
  if (team) {
    const unsigned ep = team->ctl.epoch;
    team->ctl = TeamCtl{};
    team->ctl.epoch = ep;
    team->gamma_state = 0;
    team->team_dirty = true;
  }
}

// ---- device structures --------------------------------------------------------
void Agent::build_structure() {
  ProfSection prof_(".build_structure");
  quiesce_preconditioner();
  // neighbour slots, ordered by (robot, frame)
  slot_of.clear();
  slot_key.clear();
  std::set<std::pair<int, int>> keys;
  for (const auto &m : slc) keys.insert(m.r1 == id ? std::make_pair(m.r2, m.p2) : std::make_pair(m.r1, m.p1));
  for (const auto &k : keys) {
    slot_of[k] = (int)slot_key.size();
    slot_key.push_back(k);
  }
  const int n_in = (int)slot_key.size();
  nbr_slots.clear();
  for (int sidx = 0; sidx < n_in; ++sidx) {
    NbrSlots &ns = nbr_slots[slot_key[sidx].first];
    if (ns.frames.empty()) ns.first = sidx;
    ns.frames.push_back(slot_key[sidx].second);
  }
  inbox_valid_reg.assign(n_in, 0);
  inbox_valid_aux.assign(n_in, 0);
  // Q structure by output pose
  std::vector<std::set<int>> rows(n);
  for (int j = 0; j < n; ++j) rows[j].insert(j);
  for (auto *vec : {&odom, &plc})
    for (const auto &m : *vec) {
      rows[m.p2].insert(m.p1);
      rows[m.p1].insert(m.p2);
    }
  h_q_rowptr.assign(n + 1, 0);
  h_q_col.clear();
  for (int j = 0; j < n; ++j) {
    for (int i : rows[j]) h_q_col.push_back(i);
    h_q_rowptr[j + 1] = (int)h_q_col.size();
  }
  // shared-edge terms by my pose, in slc order
  std::vector<std::vector<int>> sl(n);
  for (size_t e = 0; e < slc.size(); ++e) {
    const auto &m = slc[e];
    sl[m.r1 == id ? m.p1 : m.p2].push_back((int)e);
  }
  h_s_rowptr.assign(n + 1, 0);
  h_s_slot.clear();
  std::vector<int> s_edge;
  for (int j = 0; j < n; ++j) {
    for (int e : sl[j]) {
      const auto &m = slc[e];
      h_s_slot.push_back(slot_of[m.r1 == id ? std::make_pair(m.r2, m.p2) : std::make_pair(m.r1, m.p1)]);
      s_edge.push_back(e);
    }
    h_s_rowptr[j + 1] = (int)h_s_slot.size();
  }
  // publication entries by my pose: (neighbour, frame)
  std::vector<std::set<int>> pubs(n);
  for (const auto &m : slc) {
    if (m.r1 == id)
      pubs[m.p1].insert(m.r2);
    else
      pubs[m.p2].insert(m.r1);
  }
  h_pub_rowptr.assign(n + 1, 0);
  h_pub_entries.clear();
  for (int j = 0; j < n; ++j) {
    for (int b : pubs[j]) h_pub_entries.push_back({b, j});
    h_pub_rowptr[j + 1] = (int)h_pub_entries.size();
  }
  outbox_range.clear();
  outbox_total = 0;
  for (int b : nbrs) {
    const int cnt = (int)my_public_frames(b).size();
    outbox_range[b] = {outbox_total, cnt};
    outbox_total += cnt;
  }
  // ---- the measurements on the device: [odom | plc | slc]
  {
    const int M = num_meas();
    std::vector<double> R((size_t)M * 9), t((size_t)M * 3), ka(M), ta(M);
    std::vector<int> src(M), dst(M);
    std::vector<unsigned char> fl(M);
    for (int e = 0; e < M; ++e) {
      const Meas &m = meas_at(e);
      const bool sr = m.r1 != id, dr = m.r2 != id;
      fl[e] = (unsigned char)((sr ? 1 : 0) | (dr ? 2 : 0));
      src[e] = sr ? slot_of[{m.r1, m.p1}] : m.p1;
      dst[e] = dr ? slot_of[{m.r2, m.p2}] : m.p2;
      std::memcpy(&R[(size_t)e * 9], m.R, 9 * sizeof(double));
      std::memcpy(&t[(size_t)e * 3], m.t, 3 * sizeof(double));
      ka[e] = m.kappa;
      ta[e] = m.tau;
    }
    d_m_R.upload(R);
    d_m_t.upload(t);
    d_m_kappa.upload(ka);
    d_m_tau.upload(ta);
    d_m_src.upload(src);
    d_m_dst.upload(dst);
    d_m_flags.upload(fl);
    d_m_w.alloc(M);
    d_m_skip.alloc(M);
    d_m_resid.alloc(M);
    weights_host_dirty = true;
    resid_valid = false;
  }
  // ---- which (measurement, role) pairs land in which Q slot, in measurement order (k_assemble_values)
  {
    const size_t nq = h_q_col.size();
    std::vector<std::vector<int>> items(nq);
    auto slot = [&](int col, int row) {  // block multiplying X_col in output pose `row`
      auto b = h_q_col.begin() + h_q_rowptr[row], e = h_q_col.begin() + h_q_rowptr[row + 1];
      return (size_t)(std::lower_bound(b, e, col) - h_q_col.begin());
    };
    const int no = (int)odom.size(), np = (int)plc.size();
    for (int e = 0; e < no + np; ++e) {
      const Meas &m = meas_at(e);
      items[slot(m.p1, m.p1)].push_back(e * 4 + 0);
      items[slot(m.p2, m.p2)].push_back(e * 4 + 1);
      items[slot(m.p1, m.p2)].push_back(e * 4 + 2);
      items[slot(m.p2, m.p1)].push_back(e * 4 + 3);
    }
    for (int k = 0; k < (int)slc.size(); ++k) {
      const Meas &m = slc[k];
      const int e = no + np + k;
      if (m.r1 == id)
        items[slot(m.p1, m.p1)].push_back(e * 4 + 0);
      else
        items[slot(m.p2, m.p2)].push_back(e * 4 + 1);
    }
    std::vector<int> ptr(nq + 1, 0), flat;
    for (size_t q = 0; q < nq; ++q) {
      flat.insert(flat.end(), items[q].begin(), items[q].end());
      ptr[q + 1] = (int)flat.size();
    }
    if (flat.empty()) flat.push_back(0);
    d_qc_ptr.upload(ptr);
    d_qc_item.upload(flat);
    // ELL(8) layout of Q + overflow: columns here, values by the kernel
    constexpr int W = 8;
    std::vector<int> ec((size_t)n * W, -1), orp(n + 1, 0), oc, qdst(std::max<size_t>(nq, 1), 0);
    for (int j = 0; j < n; ++j) {
      int k = 0;
      for (int e = h_q_rowptr[j]; e < h_q_rowptr[j + 1]; ++e, ++k) {
        if (k < W) {
          ec[(size_t)j * W + k] = h_q_col[e];
          qdst[e] = j * W + k;
        } else {
          qdst[e] = -(1 + (int)oc.size());
          oc.push_back(h_q_col[e]);
        }
      }
      orp[j + 1] = (int)oc.size();
    }
    d_q_dst.upload(qdst);
    d_qe_col.upload(ec);
    d_qe_val.alloc((size_t)n * W * 16);
    d_qo_rowptr.upload(orp);
    d_qo_val.alloc(std::max<size_t>(oc.size(), 1) * 16);
    if (oc.empty()) oc.push_back(0);
    d_qo_col.upload(oc);
    d_q_val.alloc(std::max<size_t>(nq, 1) * 16);
  }
  // ---- linear-term blocks: one per shared edge, in the order of h_s_slot (per pose, slc order)
  {
    const int base = (int)(odom.size() + plc.size());
    const size_t ns = s_edge.size();
    std::vector<int> item(std::max<size_t>(ns, 1), 0), sdst(std::max<size_t>(ns, 1), 0);
    for (size_t k = 0; k < ns; ++k) item[k] = (base + s_edge[k]) * 2 + (slc[s_edge[k]].r1 == id ? 0 : 1);
    constexpr int W = 4;
    std::vector<int> ec((size_t)n * W, -1), orp(n + 1, 0), oc;
    for (int j = 0; j < n; ++j) {
      int k = 0;
      for (int e = h_s_rowptr[j]; e < h_s_rowptr[j + 1]; ++e, ++k) {
        if (k < W) {
          ec[(size_t)j * W + k] = h_s_slot[e];
          sdst[e] = j * W + k;
        } else {
          sdst[e] = -(1 + (int)oc.size());
          oc.push_back(h_s_slot[e]);
        }
      }
      orp[j + 1] = (int)oc.size();
    }
    d_s_item.upload(item);
    d_s_dst.upload(sdst);
    d_se_slot.upload(ec);
    d_se_val.alloc((size_t)n * W * 16);
    d_so_rowptr.upload(orp);
    d_so_val.alloc(std::max<size_t>(oc.size(), 1) * 16);
    if (oc.empty()) oc.push_back(0);
    d_so_slot.upload(oc);
    d_s_val.alloc(std::max<size_t>(ns, 1) * 16);
  }
  // ---- LARGE agents: per-pose incidence lists of the edge-record gradient (edge_grad.cu)
  if (n >= kEdgeGradMinPoses) {
    const int M = num_meas();
    std::vector<int> ptr(n + 1, 0);
    auto mine = [&](const Meas &m, bool dst_end) { return dst_end ? (m.r2 == id) : (m.r1 == id); };
    for (int e = 0; e < M; ++e) {
      const Meas &m = meas_at(e);
      if (mine(m, false)) ptr[m.p1 + 1]++;
      if (mine(m, true)) ptr[m.p2 + 1]++;
    }
    for (int j = 0; j < n; ++j) ptr[j + 1] += ptr[j];
    std::vector<int2> items(std::max(1, ptr[n]));
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int e = 0; e < M; ++e) {
      const Meas &m = meas_at(e);
      const int src = m.r1 == id ? m.p1 : -(slot_of[{m.r1, m.p1}] + 1);
      const int dst = m.r2 == id ? m.p2 : -(slot_of[{m.r2, m.p2}] + 1);
      if (m.r1 == id) items[fill[m.p1]++] = make_int2(e * 2 + 0, dst);   // this pose is the source: other end = dst
      if (m.r2 == id) items[fill[m.p2]++] = make_int2(e * 2 + 1, src);
    }
    d_inc_ptr.upload(ptr);
    d_inc_item.upload(items);
    d_er_rec.alloc((size_t)std::max(1, M) * 16, false);
    d_eg_partials.alloc((size_t)edge_grad_grid(n) * 2);
    d_eg_marks.alloc(2);
  } else {
    d_inc_ptr.release();
    d_inc_item.release();
    d_er_rec.release();
  }
  lc_dirty = true;

  d_q_rowptr.upload(h_q_rowptr);
  d_q_col.upload(h_q_col);
  d_s_rowptr.upload(h_s_rowptr);
  d_s_slot.upload(h_s_slot);
  d_pub_rowptr.upload(h_pub_rowptr);
  const size_t vec = (size_t)r * 4 * n;
  if (inbox_ext) fail(DPGO_B200_ERR_STATE, "the pose graph changed after the team's fabric window was laid out");
  d_inbox.alloc((size_t)2 * std::max(1, n_in) * 4 * r);
  free_pinned();
  cuda_check(cudaMallocHost((void **)&h_inbox, d_inbox.n * sizeof(double)), "cudaMallocHost");
  std::memset(h_inbox, 0, d_inbox.n * sizeof(double));
  inbox_dirty = false;
  outbox_mirror_valid = false;
  for (auto *b : {&dG, &dRg, &dRgT, &dZ, &dEta, &dDlt0, &dDlt1, &dHd, &dHdT, &dRv, &dRvT, &dRw, &dRwT, &dX2, &dX3, &dRg2, &dRg2T,
                  &dZeta})
    b->alloc(vec);
  dS.alloc((size_t)6 * n);
  dS2.alloc((size_t)6 * n);
  if (P.acceleration) dLX.alloc((size_t)kLaMax * vec);
  if (dX.n != vec) {  // not initialised yet: allocate so the views are valid
    dX.alloc(vec);
    dY.alloc(vec);
    dV.alloc(vec);
    dXinit.alloc(vec);
  }
  structure_dirty = false;
  values_dirty = true;
  precon_dirty = true;
  wiring_dirty = true;
}

// loop closures subject to re-weighting: every non-fixed private / shared loop closure; the lower-ID robot owns a
// shared edge's weight (src/PGOAgentROS.cpp:732, 1340)
void Agent::build_lc_list() {
  ProfSection prof_(".build_lc_list");
  lc_meas.clear();
  lc_mask.clear();
  const int no = (int)odom.size(), np = (int)plc.size();
  for (int k = 0; k < np; ++k)
    if (!plc[k].fixed) {
      lc_meas.push_back(no + k);
      lc_mask.push_back(1);
    }
  for (int k = 0; k < (int)slc.size(); ++k)
    if (!slc[k].fixed) {
      const int other = slc[k].r1 == id ? slc[k].r2 : slc[k].r1;
      lc_meas.push_back(no + np + k);
      lc_mask.push_back(other >= id ? 1 : 0);
    }
  std::vector<int> lm = lc_meas;
  std::vector<unsigned char> mk = lc_mask;
  if (lm.empty()) {
    lm.push_back(0);
    mk.push_back(0);
  }
  d_lc_meas.upload(lm);
  d_lc_mask.upload(mk);
  d_lc_residual.alloc(lm.size());
  lc_dirty = false;
}

void Agent::upload_weights() {
  const int M = num_meas();
  std::vector<double> w(M);
  for (int e = 0; e < M; ++e) w[e] = meas_at(e).weight;
  d_m_w.upload(w);
  weights_host_dirty = false;
}

MeasDev Agent::meas_view() const {
  return MeasDev{num_meas(), d_m_R.p, d_m_t.p, d_m_kappa.p, d_m_tau.p, d_m_w.p, d_m_skip.p, d_m_src.p, d_m_dst.p,
                 d_m_flags.p};
}
AssembleDev Agent::assemble_view() const {
  AssembleDev A{};
  A.nq = (int)h_q_col.size();
  A.ns = (int)h_s_slot.size();
  A.qc_ptr = d_qc_ptr.p; A.qc_item = d_qc_item.p; A.q_dst = d_q_dst.p;
  A.s_item = d_s_item.p; A.s_dst = d_s_dst.p;
  A.q_val = d_q_val.p; A.qe_val = d_qe_val.p; A.qo_val = d_qo_val.p;
  A.s_val = d_s_val.p; A.se_val = d_se_val.p; A.so_val = d_so_val.p;
  return A;
}

// Q and the linear-term blocks from the measurements, on the device (assemble.cu).  What crosses from the host is at
// most the weight vector (when the host changed a weight: setMeasurementWeight, a neighbour's weights) and the
// per-measurement "neighbour deactivated" flags -- never a matrix.
void Agent::build_values() {
  ProfSection prof_(".build_values");
  quiesce_preconditioner();
  if (lc_dirty) build_lc_list();
  if (weights_host_dirty) upload_weights();
  {
    // shared loop closures with a deactivated neighbour leave the problem (upstream's default; the alternative,
    // useInactiveNeighbors(true), is commented out in the wrapper: src/PGOAgentROS.cpp:151-156)
    const int M = num_meas(), base = (int)(odom.size() + plc.size());
    std::vector<unsigned char> skip(M, 0);
    if (!inactive_robots.empty())
      for (int k = 0; k < (int)slc.size(); ++k)
        if (inactive_robots.count(slc[k].r1 == id ? slc[k].r2 : slc[k].r1)) skip[base + k] = 1;
    d_m_skip.upload(skip);
  }
  cuda_check(launch_assemble_values(meas_view(), assemble_view(), 0), "k_assemble_values");
  if (has_edge_arrays()) cuda_check(launch_pack_edge_records(meas_view(), d_er_rec.p, 0), "k_pack_edge_records");
  values_dirty = false;
  precon_dirty = true;
  resid_valid = false;
}

void Agent::build_preconditioner() {
  start_preconditioner();
  finish_preconditioner();
}

// (what the kernels are wired to is dPinv: the workspace never appears in a device view)
bool Agent::precond_reserved() const {
  const size_t npad = roundup32((size_t)4 * n);
  return !need_preconditioner() || dPinv.n == npad * npad;
}

void Agent::reserve_preconditioner(bool with_workspace) {
  if (!need_preconditioner()) return;
  const size_t npad = roundup32((size_t)4 * n);
  dPinv.alloc(npad * npad, false);
  // the factorisation workspace stays allocated between rebuilds (a GNC weight update rebuilds the inverse of every
  // robot: cudaMalloc / cudaFree of a few MB each time cost more than the kernels on the tunnels robots, and
  // cudaFree synchronises the device) -- except for agents whose workspace is measured in GB: those allocate it when
  // the build starts and give it back when it ends
  if (with_workspace || npad * npad * sizeof(double) <= ((size_t)1 << 30)) dPwork.alloc(npad * npad, false);
  dPinfo.alloc(1);
  if (!pstream) {
    cuda_check(cudaStreamCreateWithFlags(&pstream, cudaStreamNonBlocking), "streamCreate (preconditioner)");
    cuda_check(cudaEventCreateWithFlags(&pevent, cudaEventDisableTiming), "eventCreate");
    cuda_check(cudaEventCreateWithFlags(&pevent_in, cudaEventDisableTiming), "eventCreate");
  }
}

void Agent::start_preconditioner() {
  if (precon_inflight) return;
  if (!need_preconditioner()) {
    precon_dirty = false;
    return;
  }
  ProfSection prof_(".start_preconditioner");
  const size_t npad = roundup32((size_t)4 * n);
  reserve_preconditioner(true);
  // the block values were assembled on the legacy stream, which a non-blocking stream does not wait for by itself
  cuda_check(cudaEventRecord(pevent_in, 0), "eventRecord");
  cuda_check(cudaStreamWaitEvent(pstream, pevent_in, 0), "streamWaitEvent");
  cuda_check(launch_scatter_blocks(dPinv.p, npad, d_q_rowptr.p, d_q_col.p, d_q_val.p, n, P.precond_lambda,
                                   (int)npad, pstream),
             "scatter_blocks");
  cuda_check(spd_inverse(dPinv.p, dPwork.p, (int)npad, dPinfo.p, pstream), "spd_inverse");
  cuda_check(cudaEventRecord(pevent, pstream), "eventRecord");
  precon_inflight = true;
}

void Agent::finish_preconditioner() {
  if (!precon_inflight) return;
  ProfSection prof_(".build_preconditioner");
  cuda_check(cudaEventSynchronize(pevent), "dense inverse");
  precon_inflight = false;
  int h_info = 0;
  cuda_check(cudaMemcpyAsync(&h_info, dPinfo.p, sizeof(int), cudaMemcpyDeviceToHost, pstream), "D2H info");
  cuda_check(cudaStreamSynchronize(pstream), "D2H info");
  const size_t npad = roundup32((size_t)4 * n);
  if (npad * npad * sizeof(double) > ((size_t)1 << 30)) dPwork.release();
  if (h_info != 0) fail(DPGO_B200_ERR_NUMERIC, "preconditioner: Q + lambda I is not positive definite");
  precon_dirty = false;
}

void Agent::quiesce_preconditioner() {
  if (!precon_inflight) return;
  cudaEventSynchronize(pevent);
  precon_inflight = false;  // precon_dirty stays set: whoever changes the values rebuilds
}

void Agent::ensure_device(int precond_mode) {
  cuda_check(use_device(device), "cudaSetDevice");
  if (structure_dirty) build_structure();
  if (values_dirty) build_values();
  if (lc_dirty) build_lc_list();
  if (precon_dirty) {
    if (precond_mode == 0) {
      reserve_preconditioner(false);
    } else {
      start_preconditioner();
      if (precond_mode == 2) finish_preconditioner();
    }
  }
  // (allocated here, not at the first launch: no cudaMalloc once kernels of several ranks wait for each other)
  if (has_edge_arrays() && need_preconditioner() && !P.acceleration && P.method == 1 && getenv("DPGO_B200_SYM_PRECOND"))
    ensure_sym_buffers();
}

AgentDev Agent::dev_view() const {
  AgentDev A{};
  A.id = id; A.n = n; A.r = r; A.n_in = (int)slot_key.size();
  A.conv_ok = weights_converged() ? 1 : 0;
  A.X = dX.p; A.Y = dY.p; A.V = dV.p; A.Xinit = dXinit.p;
  A.q_rowptr = d_q_rowptr.p; A.q_col = d_q_col.p; A.q_val = d_q_val.p;
  A.s_rowptr = d_s_rowptr.p; A.s_slot = d_s_slot.p; A.s_val = d_s_val.p;
  A.qe_col = d_qe_col.p; A.qe_val = d_qe_val.p; A.qo_rowptr = d_qo_rowptr.p; A.qo_col = d_qo_col.p; A.qo_val = d_qo_val.p;
  A.se_slot = d_se_slot.p; A.se_val = d_se_val.p; A.so_rowptr = d_so_rowptr.p; A.so_slot = d_so_slot.p; A.so_val = d_so_val.p;
  A.inbox_reg = d_inbox_reg(); A.inbox_aux = d_inbox_aux();
  A.LX = dLX.p;
  A.la_out = d_la_out;
  A.outbox_base = d_outbox;
  A.la_stride = la_stride;
  A.inbox_src = h_inbox;  // pinned + UVA: the same pointer is valid on the device
  A.inbox_doubles = (int)d_inbox.n;
  A.pub_rowptr = d_pub_rowptr.p; A.pub_dst_reg = d_pub_dst_reg.p; A.pub_dst_aux = d_pub_dst_aux.p;
  A.Pinv = dPinv.p;
  A.G = dG.p; A.Rg = dRg.p; A.RgT = dRgT.p; A.Z = dZ.p; A.eta = dEta.p; A.dlt0 = dDlt0.p; A.dlt1 = dDlt1.p;
  A.Hd = dHd.p; A.HdT = dHdT.p; A.rv = dRv.p; A.rvT = dRvT.p; A.rw = dRw.p; A.rwT = dRwT.p; A.X2 = dX2.p; A.X3 = dX3.p; A.Rg2 = dRg2.p; A.Rg2T = dRg2T.p;
  A.zeta = dZeta.p; A.S = dS.p; A.S2 = dS2.p; A.stat = d_stat;
  return A;
}

// (accepted + rejected) / total loop closures >= robustOptMinConvergenceRatio; counted like PoseGraph::statistics()
// (src/PGOAgentROS.cpp:1058-1067): weight exactly 1 = accepted, exactly 0 = rejected
bool Agent::weights_converged() const {
  if (!(P.robust_opt_min_convergence_ratio > 0.0)) return true;
  size_t total = 0, settled = 0;
  for (auto *vec : {&plc, &slc})
    for (const auto &m : *vec) {
      ++total;
      if (m.weight == 1.0 || m.weight == 0.0) ++settled;
    }
  return total == 0 || (double)settled >= P.robust_opt_min_convergence_ratio * (double)total;
}

bool Agent::all_inbox_valid(bool aux) const {
  // Slots of a DEACTIVATED neighbour never fill (it sends nothing, src/PGOAgentROS.cpp:377-400) and are not read:
  // its shared loop closures have zero blocks in Q and G (build_values), as in the oracle (Agent::iterate skips
  // inactive neighbours before the pose look-up).
  const auto &v = aux ? inbox_valid_aux : inbox_valid_reg;
  for (size_t s = 0; s < v.size(); ++s)
    if (!v[s] && !inactive_robots.count(slot_key[s].first)) return false;
  return true;
}

// ---- iterate (standalone path: a team of one, neighbours fed through the inbox)
bool Agent::iterate(bool do_opt) {
  cuda_check(use_device(device), "cudaSetDevice");
  resid_valid = false;
  Team *tm = team;
  if (state != 2) {
    iter++;
    tm->ctl.iter = iter;
    if (P.cost_type != 0) tm->ctl.robust_inner_iter = ++robust_inner_iter;
    return false;
  }
  {
    ProfSection prof_(".prepare");
    // iterate(false) never applies the preconditioner and does not build it.  (Starting the build there, so that the
    // inverses of the robots that share a GPU overlap, was measured: config 3 through the wrapper 0.042 -> 0.121 s --
    // the selected robot's cooperative solve kernel then queues behind the other robots' factorisation kernels.)
    tm->precond_mode = do_opt ? 2 : 0;
    tm->prepare(false, true);
    tm->precond_mode = 2;
  }
  const bool accel = P.acceleration != 0;
  if (!do_opt && la_used < la_valid && lookahead_usable()) {
    // this iterate(false) was computed ahead of time by the last launch (phase_lookahead): no GPU round trip
    if (la_used == 0) {
      volatile TeamCtl *c = reinterpret_cast<volatile TeamCtl *>(tm->h_result);
      while (c->la_seq != la_launch) {
      }
      std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (la_restart[la_used]) la_vsrc = la_used;
    tm->gamma_state = la_gamma[la_used];
    ++la_used;
    ++iter;
    tm->ctl.iter = iter;
    tm->ctl.gamma = tm->gamma_state;
    if (P.cost_type != 0) tm->ctl.robust_inner_iter = ++robust_inner_iter;
    status.iteration_number = iter;
    team_status[id] = get_status();
    publish_requested = true;
    if (la_used == la_valid) maybe_arm();   // the next call is iterate(true): have its kernel wait on the GPU
    return false;
  }
  const bool restart = accel && ((iter + 2) % P.restart_interval == 0);
  bool can_opt = do_opt;
  if (do_opt && !all_inbox_valid(accel && !restart)) can_opt = false;  // data matrices cannot be built
  if (armed) {
    if (can_opt) {
      // the solve kernel of this very call has been on the GPU since the last neighbour poses arrived (Agent::arm):
      // ring the doorbell -- the kernel pulls the staged inbox itself -- and wait for its result
      std::atomic_thread_fence(std::memory_order_release);
      *tm->doorbell() = tm->pending.args.seq * 2ull + 1ull;
      const unsigned long long seq = tm->pending.args.seq;
      volatile TeamCtl *c = reinterpret_cast<volatile TeamCtl *>(tm->h_result);
      bool expired = false;
      for (unsigned spins = 0; c->seq != seq; ++spins) {
        if (*tm->arm_state() == seq * 2ull) {
          expired = true;
          break;
        }
        if ((spins & 0xffff) == 0xffff) {
          const cudaError_t q = cudaStreamQuery(tm->stream);
          if (q != cudaSuccess && q != cudaErrorNotReady) cuda_check(q, "armed k_team_run");
        }
      }
      armed = false;
      if (!expired) {
        inbox_dirty = false;
        float ms = 0;
        tm->launch_finish(tm->pending, &ms);
        stats_pending = true;
        opt.f_opt = opt.gradnorm_opt = std::nan("");
        publish_requested = true;
        return true;
      }
      // the kernel gave up before it saw the doorbell: it has committed the consumed lookahead and left
      std::atomic_thread_fence(std::memory_order_acquire);
      drop_lookahead();
      outbox_stale = true;
      arm_backoff = 64;
    } else {
      disarm();
    }
  }
  if (!accel && !can_opt) {
    iter++;
    tm->ctl.iter = iter;
    if (P.cost_type != 0) tm->ctl.robust_inner_iter = ++robust_inner_iter;
    return false;
  }
  if (do_opt && arm_backoff > 0) --arm_backoff;
  {
    ProfSection prof_(can_opt ? ".solve" : ".step_without_solve");
    tm->run_forced(can_opt ? local_index : -1);
  }
  publish_requested = accel || can_opt;  // mPublishPublicPosesRequested, src/PGOAgentROS.cpp:109
  return can_opt;
}

int Agent::get_shared_pose_dict(int nbr, bool aux, int *frames, double *poses, int cap) {
  if (state != 2) fail(DPGO_B200_ERR_STATE, "getSharedPoseDictWithNeighbor: agent not initialized");
  if (aux && !P.acceleration) fail(DPGO_B200_ERR_STATE, "auxiliary poses need acceleration");
  cuda_check(use_device(device), "cudaSetDevice");
  team->prepare(false, true);
  const std::vector<int> &fr = my_public_frames(nbr);
  const int cnt = (int)fr.size();
  if (cnt > cap) fail(DPGO_B200_ERR_INVALID, "getSharedPoseDictWithNeighbor: buffer too small");
  if (cnt == 0) return 0;
  for (int k = 0; k < cnt; ++k) frames[k] = fr[k];
  bool colocated = false;
  for (Agent *o : team->agents)
    if (o->id == nbr) colocated = true;
  const size_t pb = (size_t)4 * r * sizeof(double);
  if (!colocated && la_used > 0) {
    // poses of a speculated iterate(false): X = Y there, one copy serves the regular and the auxiliary dictionary
    const auto rg = outbox_range.at(nbr);
    std::memcpy(poses, h_la_out + (size_t)(la_used - 1) * la_stride + (size_t)rg.first * 4 * r, pb * cnt);
  } else if (!colocated) {
    if (outbox_stale || !outbox_mirror_valid) team->exchange_all();
    const auto rg = outbox_range.at(nbr);
    const size_t o = (aux ? (size_t)std::max(1, outbox_total) * 4 * r : 0) + (size_t)rg.first * 4 * r;
    if (h_outbox)
      std::memcpy(poses, h_outbox + o, pb * cnt);
    else
      cuda_check(cudaMemcpy(poses, d_outbox + o, pb * cnt, cudaMemcpyDeviceToHost), "D2H outbox");
  } else {
    const double *base = aux ? dY.p : dX.p;
    for (int k = 0; k < cnt; ++k)
      cuda_check(cudaMemcpy(poses + (size_t)k * 4 * r, base + (size_t)fr[k] * 4 * r, pb, cudaMemcpyDeviceToHost),
                 "D2H pose");
  }
  return cnt;
}

void Agent::update_neighbor_poses(int nbr, bool aux, const int *frames, const double *poses, int count) {
  cuda_check(use_device(device), "cudaSetDevice");
  if (structure_dirty) build_structure();
  resid_valid = false;
  // stage in pinned host memory; the next launch uploads the inbox with one async copy
  double *inbox = h_inbox + (aux ? (size_t)slot_key.size() * 4 * r : 0);
  auto &valid = aux ? inbox_valid_aux : inbox_valid_reg;
  const size_t pb = (size_t)4 * r * sizeof(double);
  // the usual message holds exactly the poses this agent needs from `nbr`, in frame order: one copy
  auto ns = nbr_slots.find(nbr);
  if (ns != nbr_slots.end() && (int)ns->second.frames.size() == count && count > 0 &&
      std::memcmp(ns->second.frames.data(), frames, sizeof(int) * count) == 0) {
    std::memcpy(inbox + (size_t)ns->second.first * 4 * r, poses, pb * count);
    std::memset(valid.data() + ns->second.first, 1, count);
    inbox_dirty = true;
    maybe_arm(nbr, aux);
    return;
  }
  for (int k = 0; k < count; ++k) {
    auto it = slot_of.find({nbr, frames[k]});
    if (it == slot_of.end()) continue;  // a pose this agent does not need
    std::memcpy(inbox + (size_t)it->second * 4 * r, poses + (size_t)k * 4 * r, pb);
    valid[it->second] = 1;
    inbox_dirty = true;
  }
}

// Arm the next iterate(true).  Called when neighbour poses arrive (updateNeighborPoses): if this accelerated
// stand-alone RGD agent has answered all the iterate(false) calls its last launch speculated (N - 1 of them in the
// RoundRobin schedule), the next call the wrapper makes is iterate(true) -- so its solve kernel is launched NOW, with
// everything it can do without the neighbours' latest poses in front (lookahead commit, Nesterov phase, TMA prefetch
// of the preconditioner slab), and waits for the doorbell.  Nothing else may want this GPU in between: any other
// device-touching call disarms first, and a kernel nobody rings within the time-out leaves by itself.
void Agent::maybe_arm(int from_nbr, bool from_aux) {
  const bool disabled = getenv("DPGO_B200_NO_ARM") != nullptr;
  if (disabled || armed || !lookahead_usable() || P.method != 1 || P.cost_type != 0) return;
  // the waiting kernel occupies the whole device: only when this robot has it to itself (one robot per GPU, the
  // deployment of BASELINE config 2 at 8 GPUs).  With several robots on one device the kernel of whoever holds the
  // UPDATE token would queue behind it (measured: the 300 us time-out every iteration).
  static const bool shared_ok = getenv("DPGO_B200_ARM_SHARED") != nullptr;   // tests: exercise go / expiry / abort on one GPU
  if (!shared_ok && g_agents_on_device[device & 63].load() != 1) {
    // Several robots share this GPU.  The one moment at which nobody else wants it before this robot's own
    // iterate(true): the robot that holds the UPDATE token of the current iteration -- RoundRobin: my predecessor --
    // has just delivered the poses its solve produced (its auxiliary poses come last, src/PGOAgentROS.cpp:662-690), and
    // every other robot answers its iterate(false) of the next iteration from its lookahead without a launch.  Arming
    // on any earlier delivery put this kernel in front of the predecessor's solve (the 300 us time-out every iteration).
    static const bool multi_ok = getenv("DPGO_B200_NO_ARM_MULTI") == nullptr;
    const int N = P.num_robots;
    if (!multi_ok || N < 2 || !from_aux || from_nbr != (id + N - 1) % N || iter % N != id) return;
  }
  if (la_valid <= 0 || la_used != la_valid) return;
  if (structure_dirty || values_dirty || precon_dirty || wiring_dirty || lc_dirty || team->team_dirty) return;
  if (arm_backoff > 0) return;
  Team *tm = team;
  cuda_check(use_device(device), "cudaSetDevice");
  RunArgs args{};
  args.max_iters = 1;
  args.force_selected = local_index;
  args.pull_mask = 1u << local_index;
  args.skip_stats = 1;
  args.armed = 1;
  static const double timeout_us = getenv("DPGO_B200_ARM_TIMEOUT_US") ? atof(getenv("DPGO_B200_ARM_TIMEOUT_US")) : 300.0;
  args.arm_timeout_ns = (unsigned long long)(timeout_us * 1e3);
  args.arm_decision = reinterpret_cast<int *>(tm->dBar.p + 3);
  *tm->doorbell() = 0;
  *tm->arm_state() = 0;
  std::atomic_thread_fence(std::memory_order_release);
  tm->launch_begin(args, tm->grid, false, tm->pending);
  armed = true;
}

void Agent::disarm() {
  if (!armed) return;
  Team *tm = team;
  const unsigned long long seq = tm->pending.args.seq;
  *tm->doorbell() = seq * 2ull;
  for (unsigned spins = 0; *tm->arm_state() != seq * 2ull; ++spins)
    if ((spins & 0xffff) == 0xffff) {
      const cudaError_t q = cudaStreamQuery(tm->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) cuda_check(q, "armed k_team_run (abort)");
      if (q == cudaSuccess && *tm->arm_state() != seq * 2ull) break;   // (the kernel left some other way)
    }
  std::atomic_thread_fence(std::memory_order_acquire);
  armed = false;
  drop_lookahead();      // the kernel committed the consumed steps before it waited
  outbox_stale = true;
  arm_backoff = 8;
}

bool Agent::lookahead_usable() const {
  return P.acceleration && state == 2 && team == own.get() && h_la_out != nullptr && !team->window;
}

void Agent::materialize_lookahead() {
  if (armed) disarm();
  if (la_used == 0) return;
  cuda_check(use_device(device), "cudaSetDevice");
  RunArgs args{};
  args.max_iters = 1;
  args.force_selected = -1;
  args.commit_only = 1;
  team->launch_and_read(args, team->small_grid, false, nullptr);  // commits la_used steps (fills la_commit itself)
  drop_lookahead();
  outbox_stale = true;
}

// mLocalOptResult.fOpt / gradNormOpt (src/PGOAgentROS.cpp:169-172) when the launch skipped them: one
// gradient pass at X+ against the G of the solve (still cached: G is rebuilt by the next solve only)
void Agent::finish_opt_stats() {
  if (!stats_pending) return;
  if (armed) disarm();
  cuda_check(use_device(device), "cudaSetDevice");
  team->prepare(false, true);  // needs X+ (X2) and G only
  const int grid = team->grid;
  if (d_stat_partials.n != (size_t)grid * 2) d_stat_partials.alloc((size_t)grid * 2);
  cuda_check(launch_post_stats(team->T.ag[local_index], dX2.p, d_stat_partials.p, grid, team->stream), "k_post_stats");
  std::vector<double> part((size_t)grid * 2);
  cuda_check(cudaMemcpyAsync(part.data(), d_stat_partials.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost,
                             team->stream),
             "D2H partials");
  cuda_check(cudaStreamSynchronize(team->stream), "k_post_stats");
  double f = 0, g2 = 0;
  for (int b = 0; b < grid; ++b) {
    f += part[(size_t)b * 2];
    g2 += part[(size_t)b * 2 + 1];
  }
  opt.f_opt = f;
  opt.gradnorm_opt = std::sqrt(g2);
  stats_pending = false;
}

dpgo_b200_status Agent::get_status() const {
  dpgo_b200_status s = status;
  s.agent_id = id;
  s.state = state;
  s.instance_number = instance;
  s.iteration_number = iter;
  return s;
}

bool Agent::should_terminate() const {
  if (iter > P.max_num_iters) return true;
  if (P.cost_type != 0 && weight_update_count < P.robust_opt_num_weight_updates) return false;
  for (int rid = 0; rid < P.num_robots; ++rid) {
    if (inactive_robots.count(rid)) continue;   // a robot the leader has deactivated no longer has a say (src/PGOAgentROS.cpp:195)
    auto it = team_status.find(rid);
    if (it == team_status.end()) return false;
    if (it->second.state != 2 || !it->second.ready_to_terminate) return false;
  }
  return true;
}

bool Agent::should_update_weights() const {
  if (P.cost_type == 0) return false;
  if (weight_update_count >= P.robust_opt_num_weight_updates) return false;
  if (robust_inner_iter >= P.robust_opt_inner_iters) return true;
  for (int rid = 0; rid < P.num_robots; ++rid) {
    if (inactive_robots.count(rid)) continue;   // a robot the leader has deactivated no longer has a say (src/PGOAgentROS.cpp:195)
    auto it = team_status.find(rid);
    if (it == team_status.end()) return false;
    if (it->second.state != 2 || !it->second.ready_to_terminate) return false;
  }
  return true;
}

double Agent::robust_weight(double res) const {
  if (P.cost_type == 0) return 1.0;
  const double bsq = P.gnc_barc * P.gnc_barc;
  const double upper = (mu + 1.0) / mu * bsq, lower = mu / (mu + 1.0) * bsq, rsq = res * res;
  if (rsq >= upper) return 0.0;
  if (rsq <= lower) return 1.0;
  return std::sqrt(bsq * mu * (mu + 1.0) / rsq) - mu;
}

void Agent::update_measurement_weights() {
  if (state != 2) return;
  team->gnc_update_all();
}

// computeMeasurementResidual (src/PGOAgentROS.cpp:1049).  The TERMINATE handler asks for every active loop closure in
// turn (:1044-1057; ~1000 per robot on the tunnels dataset): the first call runs ONE launch over all measurements and
// the rest are served from the cached array until X, the inbox or the graph change.
void Agent::refresh_residuals() {
  cuda_check(use_device(device), "cudaSetDevice");
  team->prepare();  // materialises a consumed lookahead, flushes the staged inbox
  const int M = num_meas();
  h_resid.assign(M, 0.0);
  if (M > 0) {
    ResidualJob J{M, nullptr, nullptr, d_m_resid.p, 0.0, 0.0, 0};
    cuda_check(launch_measurement_residuals(meas_view(), J, r, dX.p, d_inbox_reg(), team->stream), "k_measurement_residuals");
    cuda_check(cudaMemcpyAsync(h_resid.data(), d_m_resid.p, sizeof(double) * M, cudaMemcpyDeviceToHost, team->stream),
               "D2H residuals");
    cuda_check(cudaStreamSynchronize(team->stream), "k_measurement_residuals");
  }
  resid_valid = true;
}

bool Agent::compute_residual(int r1, int p1, int r2, int p2, double *res) {
  if (state != 2) return false;
  const int idx = measurement_index(r1, p1, r2, p2);
  if (idx < 0) fail(DPGO_B200_ERR_MISSING, "computeMeasurementResidual: no such measurement");
  if (structure_dirty) resid_valid = false;
  if (!resid_valid) refresh_residuals();
  const Meas &m = meas_at(idx);
  if (m.r1 != id) {
    auto it = slot_of.find({m.r1, m.p1});
    if (it == slot_of.end() || !inbox_valid_reg[it->second]) return false;
  }
  if (m.r2 != id) {
    auto it = slot_of.find({m.r2, m.p2});
    if (it == slot_of.end() || !inbox_valid_reg[it->second]) return false;
  }
  *res = h_resid[idx];
  return true;
}

// ============================================================================
// Team
// ============================================================================
Team::Team(int device_) : device(device_) {
  cuda_check(use_device(device), "cudaSetDevice");
  cuda_check(cudaEventCreate(&ev0), "eventCreate");
  cuda_check(cudaEventCreate(&ev1), "eventCreate");
  cuda_check(cudaStreamCreate(&stream), "streamCreate");  // blocking stream: ordered against legacy-stream setup work
}

Team::~Team() {
  for (Agent *a : agents) {
    if (a->team != this) continue;
    if (a->own.get() == this || !a->own) {
      a->team = nullptr;
      continue;
    }
    a->team = a->own.get();
    a->own->agents.assign(1, a);
    a->local_index = 0;
    a->wiring_dirty = true;
    a->own->team_dirty = true;
  }
  fabric_close();
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (stream) cudaStreamDestroy(stream);
  if (h_result) cudaFreeHost(h_result);
}

void Team::add(Agent *a) {
  if ((int)agents.size() >= kMaxLocal) fail(DPGO_B200_ERR_INVALID, "too many agents on one device (max 8)");
  if (a->device != device) fail(DPGO_B200_ERR_INVALID, "agent lives on another device");
  if (!agents.empty() && (agents[0]->r != a->r || agents[0]->P.num_robots != a->P.num_robots))
    fail(DPGO_B200_ERR_INVALID, "agents of one team must share r and num_robots");
  if (a->la_used > 0) a->materialize_lookahead();
  a->drop_lookahead();
  if (a->team && a->team != this && a->team != a->own.get()) a->team->remove(a);
  if (agents.empty() && a->own && a->own.get() != this) {
    const unsigned ep = ctl.epoch;
    ctl = a->own->ctl;
    ctl.epoch = ep;
    gamma_state = a->own->gamma_state;
  }
  if (a->own && a->own.get() != this) a->own->agents.clear();
  a->team = this;
  a->local_index = (int)agents.size();
  agents.push_back(a);
  for (Agent *o : agents) o->wiring_dirty = true;
  team_dirty = true;
}

void Team::remove(Agent *a) {
  auto it = std::find(agents.begin(), agents.end(), a);
  if (it == agents.end()) return;
  agents.erase(it);
  for (size_t i = 0; i < agents.size(); ++i) {
    agents[i]->local_index = (int)i;
    agents[i]->wiring_dirty = true;
  }
  team_dirty = true;
  if (a->own && a->own.get() != this) {
    a->team = a->own.get();
    a->own->agents.assign(1, a);
    a->local_index = 0;
    a->wiring_dirty = true;
    a->own->team_dirty = true;
  } else {
    a->team = nullptr;
  }
}

void Team::prepare(bool need_inbox, bool keep_lookahead) {
  cuda_check(use_device(device), "cudaSetDevice");
  if (agents.empty()) fail(DPGO_B200_ERR_STATE, "team has no agents");
  if (!keep_lookahead)
    for (Agent *a : agents) {
      if (a->armed) a->disarm();
      if (a->la_used > 0) a->materialize_lookahead();  // whoever comes next reads X / Y / V directly
    }
  bool rewire = team_dirty;
  for (Agent *a : agents) {
    // (a stale preconditioner alone changes no pointer once its buffers exist: no re-wiring for it)
    if (a->structure_dirty || a->values_dirty || (a->precon_dirty && !a->precond_reserved())) rewire = true;
    a->ensure_device(std::min(precond_mode, 1));   // every agent's dense inverse is in flight before the first is waited for
    if (a->wiring_dirty) rewire = true;
  }
  if (precond_mode == 2)
    for (Agent *a : agents) a->ensure_device(2);
  if (need_inbox) flush_inboxes();
  if (!rewire) return;
  layout_result();
  const int r = agents[0]->r;
  for (Agent *a : agents) {
    std::vector<double *> reg(a->h_pub_entries.size()), aux(a->h_pub_entries.size());
    std::map<int, std::vector<int>> frames_of;
    for (size_t e = 0; e < a->h_pub_entries.size(); ++e) {
      const int b = a->h_pub_entries[e].first, f = a->h_pub_entries[e].second;
      Agent *peer = nullptr;
      for (Agent *o : agents)
        if (o->id == b) peer = o;
      if (peer) {
        auto it = peer->slot_of.find({a->id, f});
        if (it == peer->slot_of.end()) {
          // the peer does not hold this shared edge (measurements not synchronised): park in the outbox
          peer = nullptr;
        } else {
          reg[e] = peer->d_inbox_reg() + (size_t)it->second * 4 * r;
          aux[e] = peer->d_inbox_aux() + (size_t)it->second * 4 * r;
        }
      }
      if (!peer) {
        auto &fr = frames_of[b];
        if (fr.empty()) fr = a->my_public_frames(b);
        const int idx = (int)(std::lower_bound(fr.begin(), fr.end(), f) - fr.begin());
        auto rt = routes.find({a->id, b});
        if (rt != routes.end() && peer_base[rt->second.peer]) {
          // remote neighbour reachable over the fabric: store straight into its inbox on the other GPU
          unsigned char *pb = peer_base[rt->second.peer];
          reg[e] = reinterpret_cast<double *>(pb + rt->second.off_reg) + (size_t)idx * 4 * r;
          aux[e] = reinterpret_cast<double *>(pb + rt->second.off_aux) + (size_t)idx * 4 * r;
        } else {
          const size_t o = (size_t)(a->outbox_range.at(b).first + idx) * 4 * r;
          reg[e] = a->d_outbox_reg() + o;
          aux[e] = a->d_outbox_aux() + o;
        }
      }
    }
    if (reg.empty()) {
      reg.push_back(nullptr);
      aux.push_back(nullptr);
    }
    a->d_pub_dst_reg.upload(reg);
    a->d_pub_dst_aux.upload(aux);
    a->wiring_dirty = false;
  }
  // TeamDev
  std::memset(&T, 0, sizeof(T));
  T.num_local = (int)agents.size();
  T.num_robots = agents[0]->P.num_robots;
  for (int i = 0; i < kMaxRobots; ++i) T.local_of_robot[i] = -1;
  T.pose_prefix[0] = 0;
  for (size_t i = 0; i < agents.size(); ++i) {
    T.local_of_robot[agents[i]->id] = (int)i;
    T.ag[i] = agents[i]->dev_view();
    T.pose_prefix[i + 1] = T.pose_prefix[i] + agents[i]->n;
  }
  const dpgo_b200_params &P = agents[0]->P;
  T.p.method = P.method;
  T.p.rgd_stepsize = P.rgd_stepsize;
  T.p.rgd_use_precond = P.rgd_use_preconditioner;
  T.p.rtr_iterations = P.rtr_iterations;
  T.p.rtr_tcg_iterations = P.rtr_tcg_iterations;
  T.p.rtr_initial_radius = P.rtr_initial_radius;
  T.p.gradnorm_tol = P.gradnorm_tol;
  T.p.acceleration = P.acceleration;
  T.p.restart_interval = P.restart_interval;
  T.p.robust = P.cost_type != 0;
  T.p.robust_num_weight_updates = P.robust_opt_num_weight_updates;
  T.p.robust_inner_iters = P.robust_opt_inner_iters;
  T.p.max_num_iters = P.max_num_iters;
  T.p.rel_change_tol = P.rel_change_tol;
  if (grid <= 0) {
    grid = max_coop_grid(device);
    if (const char *g = getenv("DPGO_B200_GRID")) {   // diagnostics: CTAs of the persistent kernel
      const int v = atoi(g);
      if (v > 0 && v < grid) grid = v;
    }
  }
  dBar.alloc(4);  // [0]: full-grid launches, [1]: small-grid launches, [2]: last-block-done counter, [3]: armed-launch decision
  {
    int total = 0;
    for (Agent *a : agents) total += a->n;
    small_grid = std::max(1, std::min(grid, (total + kGroupsPerCta - 1) / kGroupsPerCta));
  }
  dSlots.alloc((size_t)2 * grid * kRed);
  dDefer.alloc((size_t)agents.size() * grid * (kThreads / 32) * 8);
  T.gs.counter = dBar.p;
  T.gs.slots = dSlots.p;
  T.defer = dDefer.p;
  T.ctl = reinterpret_cast<TeamCtl *>(d_result);
  T.prof = prof_iters > 0 ? dProf.p : nullptr;
  T.prof_iters = prof_iters;
  T.prof_cta = prof_cta;
  T.done_counter = dBar.p + 2;
  if (fab_world > 1 && window) {
    Fabric &Fb = T.fab;
    Fb.world = fab_world;
    Fb.rank = fab_rank;
    Fb.flags = reinterpret_cast<unsigned long long *>(window);
    Fb.payload = Fb.flags + kMaxRanks;
    Fb.prog = Fb.flags + 3 * kMaxRanks;
    for (int s = 0; s < fab_world; ++s) {
      unsigned long long *pf = reinterpret_cast<unsigned long long *>(peer_base[s]);
      Fb.peer_flags[s] = pf;
      Fb.peer_payload[s] = pf ? pf + kMaxRanks : nullptr;
      Fb.peer_prog[s] = pf ? pf + 3 * kMaxRanks : nullptr;
    }
    // ranks that host a neighbour of a local robot: the only ones the point-to-point protocol talks to
    Fb.nbr_ranks = 0;
    for (size_t i = 0; i < agents.size(); ++i) {
      unsigned m = 0;
      for (int b : agents[i]->nbrs) {
        auto rt = routes.find({agents[i]->id, b});
        if (rt != routes.end()) m |= 1u << rt->second.peer;
      }
      Fb.agent_nbr_ranks[i] = m;
      Fb.nbr_ranks |= m;
    }
    Fb.local_mask = 0;
    for (Agent *a : agents) Fb.local_mask |= 1ull << a->id;
    Fb.timeout_ns = (unsigned long long)(fab_timeout_s * 1e9);
  }
  team_dirty = false;
}

// ---- multi-GPU fabric ---------------------------------------------------------
void Team::fabric_init(int world, int rank) {
  if (world < 2 || world > kMaxRanks || rank < 0 || rank >= world)
    fail(DPGO_B200_ERR_INVALID, "fabric_init: need 2 <= world <= 8 and 0 <= rank < world");
  if (window) fail(DPGO_B200_ERR_STATE, "fabric_init: already initialised");
  prepare();
  auto up = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
  size_t off = kFabricHeaderBytes;
  std::vector<size_t> offs;
  for (Agent *a : agents) {
    offs.push_back(off);
    off += up(a->d_inbox.n * sizeof(double), 256);
  }
  window_bytes = off;
  cuda_check(cudaMalloc((void **)&window, window_bytes), "cudaMalloc fabric window");
  cuda_check(cudaMemset(window, 0, window_bytes), "cudaMemset fabric window");
  cuda_check(cudaStreamSynchronize(stream), "sync before moving the inboxes");
  for (size_t i = 0; i < agents.size(); ++i) {
    Agent *a = agents[i];
    double *dst = reinterpret_cast<double *>(window + offs[i]);
    cuda_check(cudaMemcpy(dst, a->d_inbox.p, a->d_inbox.n * sizeof(double), cudaMemcpyDeviceToDevice), "move inbox");
    a->inbox_ext = dst;
    a->wiring_dirty = true;
  }
  // no allocation (cudaMalloc / cudaFree synchronise the device) may happen once the ranks' kernels wait for each other
  dGammaTab.alloc((size_t)1 << 16, false);
  if (dProf.n < 4096 + 256) dProf.alloc((size_t)4096 + 256);
  fab_world = world;
  fab_rank = rank;
  fab_seq = 0;
  fab_step = 0;
  peer_base[rank] = window;
  team_dirty = true;
}

void Team::fabric_import(int peer, const cudaIpcMemHandle_t *handle, void *same_process_base) {
  if (!window) fail(DPGO_B200_ERR_STATE, "fabric_import before fabric_init");
  if (peer < 0 || peer >= fab_world || peer == fab_rank) fail(DPGO_B200_ERR_INVALID, "fabric_import: bad peer rank");
  cuda_check(use_device(device), "cudaSetDevice");
  if (same_process_base) {
    // a team of this process on another device: plain peer access
    cudaPointerAttributes at{};
    cuda_check(cudaPointerGetAttributes(&at, same_process_base), "fabric_import: pointer attributes");
    if (at.device != device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled)
        cudaGetLastError();
      else
        cuda_check(e, "cudaDeviceEnablePeerAccess");
    }
    peer_base[peer] = static_cast<unsigned char *>(same_process_base);
    peer_is_ipc[peer] = false;
  } else {
    if (!handle) fail(DPGO_B200_ERR_INVALID, "fabric_import: null handle");
    void *p = nullptr;
    cuda_check(cudaIpcOpenMemHandle(&p, *handle, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    peer_base[peer] = static_cast<unsigned char *>(p);
    peer_is_ipc[peer] = true;
  }
  for (Agent *a : agents) a->wiring_dirty = true;
  team_dirty = true;
}

void Team::fabric_route(int robot, int nbr, int peer, size_t off_reg, size_t off_aux) {
  if (!window) fail(DPGO_B200_ERR_STATE, "fabric_route before fabric_init");
  if (peer < 0 || peer >= fab_world || peer == fab_rank) fail(DPGO_B200_ERR_INVALID, "fabric_route: bad peer rank");
  Agent *me = nullptr;
  for (Agent *a : agents)
    if (a->id == robot) me = a;
  if (!me) fail(DPGO_B200_ERR_INVALID, "fabric_route: robot is not in this team");
  if (!me->nbrs.count(nbr)) fail(DPGO_B200_ERR_MISSING, "fabric_route: not a neighbour");
  routes[{robot, nbr}] = Route{peer, off_reg, off_aux};
  me->wiring_dirty = true;
  team_dirty = true;
}

void Team::fabric_close() {
  if (!window) return;
  cudaSetDevice(device);
  cudaStreamSynchronize(stream);
  for (int s = 0; s < kMaxRanks; ++s) {
    if (peer_base[s] && peer_is_ipc[s]) cudaIpcCloseMemHandle(peer_base[s]);
    peer_base[s] = nullptr;
    peer_is_ipc[s] = false;
  }
  for (Agent *a : agents) {
    // give the agents their private inboxes back (contents preserved)
    if (a->inbox_ext && a->d_inbox.p)
      cudaMemcpy(a->d_inbox.p, a->inbox_ext, a->d_inbox.n * sizeof(double), cudaMemcpyDeviceToDevice);
    a->inbox_ext = nullptr;
    a->wiring_dirty = true;
  }
  cudaFree(window);
  window = nullptr;
  window_bytes = 0;
  fab_world = 0;
  routes.clear();
  std::memset(&T.fab, 0, sizeof(T.fab));
  team_dirty = true;
}

// Every rank calls this with the same arguments at the same point of the schedule; the kernels of
// the ranks meet at the fabric barriers.  Returns when the launch ended on THIS rank -- all ranks
// leave at the same global iteration (the control decisions are taken from identical data).
dpgo_b200_run_result Team::fabric_run(int max_iters, bool stop_on_terminate) {
  dpgo_b200_run_result res{};
  if (fab_world < 2 || !window) fail(DPGO_B200_ERR_STATE, "fabric_run: fabric_init first");
  prepare();
  for (int s = 0; s < fab_world; ++s)
    if (!peer_base[s]) fail(DPGO_B200_ERR_STATE, "fabric_run: a peer window has not been imported");
  for (Agent *a : agents) {
    if (a->state != 2) fail(DPGO_B200_ERR_STATE, "fabric_run: every agent must be initialized in the global frame");
    for (int b : a->nbrs) {
      bool local = false;
      for (Agent *o : agents) local |= (o->id == b);
      if (!local && !routes.count({a->id, b})) fail(DPGO_B200_ERR_STATE, "fabric_run: a remote neighbour has no route");
    }
    if (!a->all_inbox_valid(false) || (a->P.acceleration && !a->all_inbox_valid(true)))
      fail(DPGO_B200_ERR_MISSING, "fabric_run: neighbour poses missing (exchange_all on every rank, then mark the inboxes)");
  }
  RunArgs args{};
  args.max_iters = std::min(max_iters, 1 << 16);
  args.force_selected = -2;
  args.stop_on_terminate = stop_on_terminate ? 1 : 0;
  args.fabric = 1;
  args.parallel = parallel_schedule_checked();
  if (const char *fv = getenv("DPGO_B200_FAB_VARIANT")) args.fab_variant = atoi(fv);
  if (edge_grad_loop(grid)) {
    // LARGE agents: one tick per launch (the gradient runs in k_edge_grad in front of it); every rank takes the same
    // decision, so the ranks still meet launch by launch
    dpgo_b200_run_result acc{};
    for (int it = 0; it < max_iters; ++it) {
      RunArgs a1 = args;
      a1.max_iters = 1;
      a1.skip_stats = it + 1 < max_iters ? 1 : 0;
      cuda_check(launch_fabric_rendezvous(T.fab, fab_seq + 1, stream), "k_fabric_rendezvous");
      fab_seq += 1;
      T.fab.seq0 = fab_seq;
      T.fab.step0 = fab_step;
      float ms1 = 0;
      const int l0 = launches;
      launch_and_read(a1, grid, true, &ms1);
      fab_seq = ctl.fab_seq;
      fab_step = ctl.fab_step;
      acc.device_ms += ms1;
      acc.kernel_launches += launches - l0;
      acc.iterations += ctl.iters_done;
      acc.stop_reason = ctl.stop_reason;
      if (ctl.stop_reason == -1) {
        cuda_check(cudaMemset(dBar.p, 0, dBar.n * sizeof(unsigned long long)), "reset barrier");
        fail(DPGO_B200_ERR_CUDA, "fabric_run: timed out waiting for a peer GPU");
      }
      if (ctl.stop_reason == 1) acc.terminated = 1;
      if ((ctl.stop_reason == 1 && stop_on_terminate) || ctl.stop_reason == 2 || ctl.iters_done == 0) break;
    }
    return acc;
  }
  // start together: a one-warp kernel in front of the timed launch meets the other ranks (all-rank barrier in the
  // windows), so the CUDA events around the persistent kernel do not count the skew between the processes' launch
  // calls -- with K = 20 steps of 20 us that skew used to be most of the measurement
  const int use_grid = grid;
  cuda_check(launch_fabric_rendezvous(T.fab, fab_seq + 1, stream), "k_fabric_rendezvous");
  fab_seq += 1;
  T.fab.seq0 = fab_seq;
  T.fab.step0 = fab_step;
  // diagnostics: DPGO_B200_FAB_PROF=<cta> records clock64() marks of that CTA for the first 64 steps of every launch
  static const char *fab_prof = getenv("DPGO_B200_FAB_PROF");
  if (fab_prof) {
    cuda_check(cudaMemsetAsync(dProf.p, 0, dProf.n * sizeof(long long), stream), "clear prof");
    T.prof = dProf.p;
    T.prof_iters = std::min(64, args.max_iters);
    T.prof_cta = atoi(fab_prof);
  }
  float ms = 0;
  launch_and_read(args, use_grid, true, &ms);
  fab_seq = ctl.fab_seq;
  fab_step = ctl.fab_step;
  if (fab_prof && ctl.iters_done >= 64) {
    std::vector<long long> h(64 * 16);
    cuda_check(cudaMemcpy(h.data(), dProf.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost), "D2H prof");
    // marks: 0 step start, 7 after the write-after-read wait, 1 Nesterov done, 2 barrier + post, 8 after the gate,
    // 3 gradient done, 4 barrier + post, 5 step done, 6 end
    const int order[] = {0, 7, 1, 2, 8, 3, 4, 5, 6};
    const char *nm[] = {"war_wait", "nesterov", "barrier_post", "gate", "gradient", "barrier_post2", "step", "tail"};
    double acc_sel[9] = {0}, acc_non[9] = {0};
    int nsel = 0, nnon = 0;
    for (int it = 8; it < 63; ++it) {
      const long long *m = &h[(size_t)it * 16];
      const bool sel = m[3] != 0;
      long long prev = m[0];
      double seg[9];
      for (int k = 1; k < 9; ++k) {
        long long cur = m[order[k]];
        if (cur == 0 || cur < prev) cur = prev;   // mark not hit in this step
        seg[k - 1] = (double)(cur - prev);
        prev = cur;
      }
      seg[8] = (double)(h[(size_t)(it + 1) * 16] - m[0]);   // whole step
      for (int k = 0; k < 9; ++k) (sel ? acc_sel : acc_non)[k] += seg[k];
      (sel ? nsel : nnon)++;
    }
    char path[256];
    snprintf(path, sizeof path, "gpurun_out/fabprof_rank%d.txt", fab_rank);
    if (FILE *fp = fopen(path, "a")) {
      fprintf(fp, "launch of %d steps, cta %d, cycles per segment\n  with the selected robot here (%d steps):", ctl.iters_done, T.prof_cta, nsel);
      for (int k = 0; k < 8; ++k) fprintf(fp, " %s=%.0f", nm[k], nsel ? acc_sel[k] / nsel : 0.0);
      fprintf(fp, " step_total=%.0f\n  without (%d steps):", nsel ? acc_sel[8] / nsel : 0.0, nnon);
      for (int k = 0; k < 8; ++k) fprintf(fp, " %s=%.0f", nm[k], nnon ? acc_non[k] / nnon : 0.0);
      fprintf(fp, " step_total=%.0f\n", nnon ? acc_non[8] / nnon : 0.0);
      fclose(fp);
    }
  }
  res.device_ms = ms;
  res.kernel_launches = 1;
  res.iterations = ctl.iters_done;
  res.stop_reason = ctl.stop_reason;
  if (ctl.stop_reason == -1) {
    // a peer never showed up: the barrier counter is poisoned, start clean next time
    cuda_check(cudaMemset(dBar.p, 0, dBar.n * sizeof(unsigned long long)), "reset barrier");
    fail(DPGO_B200_ERR_CUDA, "fabric_run: timed out waiting for a peer GPU");
  }
  if (ctl.stop_reason == 1) res.terminated = 1;
  return res;
}

void Team::layout_result() {
  auto up = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
  size_t off = up(kCtlBytes + kStatBytes * agents.size(), 256);
  std::vector<size_t> offs;
  size_t dev_doubles = 0;
  for (Agent *a : agents) {
    if (device_outbox) {
      offs.push_back(dev_doubles);
      dev_doubles += up(a->outbox_doubles(), 32);
    } else {
      offs.push_back(off);
      off += up(a->outbox_doubles() * sizeof(double), 256);
    }
  }
  if (device_outbox) dOutboxAll.alloc(std::max<size_t>(dev_doubles, 32));
  // a stand-alone accelerated agent also gets kLaMax lookahead outboxes (regular half only: X = Y there)
  size_t la_off = 0;
  const bool la_team = agents.size() == 1 && agents[0]->own.get() == this && agents[0]->P.acceleration && !device_outbox;
  if (la_team) {
    la_off = off;
    off += up((size_t)kLaMax * (agents[0]->outbox_doubles() / 2) * sizeof(double), 256);
  }
  if (off != result_bytes) {
    cuda_check(cudaDeviceSynchronize(), "sync before result realloc");
    if (h_result) cudaFreeHost(h_result);
    d_result = h_result = nullptr;
    // zero-copy: the kernels write control state, statistics and outboxes straight into mapped
    // pinned host memory, so the per-robot API needs no D2H copy call after a launch
    cuda_check(cudaHostAlloc((void **)&h_result, off, cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc");
    cuda_check(cudaHostGetDevicePointer((void **)&d_result, h_result, 0), "cudaHostGetDevicePointer");
    result_bytes = off;
  }
  cuda_check(cudaDeviceSynchronize(), "sync before result reset");
  std::memset(h_result, 0, result_bytes);
  for (size_t i = 0; i < agents.size(); ++i) {
    Agent *a = agents[i];
    a->d_stat = reinterpret_cast<AgentStat *>(d_result + kCtlBytes + kStatBytes * i);
    a->h_stat = reinterpret_cast<AgentStat *>(h_result + kCtlBytes + kStatBytes * i);
    if (device_outbox) {
      a->d_outbox = dOutboxAll.p + offs[i];
      a->h_outbox = nullptr;
    } else {
      a->d_outbox = reinterpret_cast<double *>(d_result + offs[i]);
      a->h_outbox = reinterpret_cast<double *>(h_result + offs[i]);
    }
    a->outbox_stale = true;
    a->outbox_mirror_valid = false;
    a->drop_lookahead();
    a->d_la_out = a->h_la_out = nullptr;
    a->la_stride = 0;
    if (la_team) {
      a->d_la_out = reinterpret_cast<double *>(d_result + la_off);
      a->h_la_out = reinterpret_cast<double *>(h_result + la_off);
      a->la_stride = (int)(a->outbox_doubles() / 2);
    }
  }
}

void Team::flush_inboxes() {
  for (Agent *a : agents)
    if (a->inbox_dirty && a->d_inbox.n) {
      cuda_check(cudaMemcpyAsync(a->inbox_base(), a->h_inbox, a->d_inbox.n * sizeof(double), cudaMemcpyHostToDevice,
                                 stream),
                 "H2D inbox");
      a->inbox_dirty = false;
    }
}

void Team::read_back() {
  ctl = *h_ctl();
  for (Agent *a : agents) {
    const AgentStat st = *a->h_stat;
    a->iter = ctl.iter;
    a->robust_inner_iter = ctl.robust_inner_iter;
    if (st.optimized) {
      a->opt.success = 1;
      a->opt.f_init = st.f_init;
      a->opt.f_opt = st.f_opt;
      a->opt.gradnorm_init = st.gn_init;
      a->opt.gradnorm_opt = st.gn_opt;
      a->opt.relative_change = st.relchange;
      a->opt.tcg_iters = st.tcg_iters;
      a->opt.rtr_outer_iters = st.rtr_outer;
      a->opt.rtr_rejections = st.rtr_rej;
      a->status.relative_change = st.relchange;
      a->status.ready_to_terminate = st.ready;
    }
    a->status.iteration_number = a->iter;
    a->status.state = a->state;
    a->team_status[a->id] = a->get_status();
    a->outbox_mirror_valid = true;
    a->outbox_stale = false;
  }
}

void Agent::ensure_sym_buffers() {
  if (sym_ntiles > 0 && dZt.n == (size_t)r * 4 * n) return;
  std::vector<int> first;
  sym_ntiles = sym_precond_tiles(4 * n, r, first);
  sym_npanels = (int)first.size() - 1;
  d_sym_first_tile.upload(first);
  d_sym_partials.alloc(sym_precond_partial_doubles(sym_ntiles, r), false);
  d_sym_counter.alloc(1);
  dZt.alloc((size_t)r * 4 * n);
}

bool Team::edge_grad_loop(int use_grid) const {
  if (agents.empty()) return false;
  const dpgo_b200_params &P = agents[0]->P;
  if (P.method != 1 || !P.rgd_use_preconditioner || P.acceleration) return false;
  const bool disabled = getenv("DPGO_B200_NO_EDGE_GRAD") != nullptr;   // diagnostics: keep the in-kernel gradient
  if (disabled) return false;
  if (!team_needs_streaming(T, use_grid)) return false;
  for (const Agent *a : agents)
    if (!a->has_edge_arrays()) return false;
  return true;
}

// local agents (bit mask) whose gradient of this launch comes from k_edge_grad
unsigned Team::ext_grad_mask(const RunArgs &args, int use_grid) const {
  if (args.max_iters != 1 || args.mode != 0 || args.commit_only || !edge_grad_loop(use_grid)) return 0;
  if (args.parallel) return (agents.size() >= 32) ? ~0u : ((1u << agents.size()) - 1u);
  int sel = args.force_selected;
  if (sel == -2) sel = T.local_of_robot[ctl.selected];
  return sel >= 0 ? (1u << sel) : 0u;
}

// launch the persistent kernel (control state travels as a kernel argument), queue ONE read-back
// copy of the result block [ctl | stats | outboxes], synchronise once
void Team::launch_and_read(const RunArgs &args_in, int use_grid, bool timed, float *ms) {
  PendingLaunch pl;
  launch_begin(args_in, use_grid, timed, pl);
  launch_finish(pl, ms);
}

void Team::launch_begin(const RunArgs &args_in, int use_grid, bool timed, PendingLaunch &pl) {
  RunArgs args = args_in;
  args.seq = ++seq;
  for (Agent *a : agents) a->resid_valid = false;
  // Nesterov sequences (a7): gamma_k = (1 + sqrt(1 + 4 N^2 gamma_{k-1}^2)) / 2N, alpha_k = 1 / (gamma_k N),
  // reset on restart iterations ((iter + 1) % restartInterval == 0)
  const dpgo_b200_params &P = agents[0]->P;
  const int N = P.num_robots;
  args.gamma_tab = nullptr;
  // stand-alone lookahead: commit what the host consumed, speculate the next N-1 iterate(false) steps
  Agent *la = (agents.size() == 1 && agents[0]->lookahead_usable() && args.max_iters == 1 && args.mode == 0 &&
               args.force_selected >= -1)
                  ? agents[0]
                  : nullptr;
  int la_depth = 0;
  if (la) {
    args.la_commit = la->la_used;
    args.la_vsrc = la->la_vsrc;
    if (!args.commit_only && !P.cost_type) la_depth = std::min(kLaMax, std::max(0, N - 1));
    args.la_depth = la_depth;
  }
  if (P.acceleration && args.commit_only) {
    // nothing advances
  } else if (P.acceleration) {
    const int K = args.max_iters + la_depth;
    h_gamma_tab.resize(K);
    h_gamma_state.resize(K);
    double g = gamma_state;
    for (int i = 0; i < K; ++i) {
      g = next_gamma(g, N);
      h_gamma_tab[i] = make_double2(g, 1.0 / (g * N));
      if ((ctl.iter + i + 2) % P.restart_interval == 0) g = 0;  // restart after this iteration
      h_gamma_state[i] = g;
    }
    if (args.mode == 2) {  // second half of a split iteration: the sequences were advanced by the first half
      args.gamma0 = last_gamma_use;
      args.alpha0 = last_alpha_use;
    } else if (args.max_iters == 1) {
      args.gamma0 = last_gamma_use = h_gamma_tab[0].x;
      args.alpha0 = last_alpha_use = h_gamma_tab[0].y;
      for (int j = 0; j < la_depth; ++j) {
        const bool rst = (ctl.iter + (j + 1) + 2) % P.restart_interval == 0;
        args.la_tab[j] = make_double2(h_gamma_tab[1 + j].y, rst ? 1.0 : 0.0);
      }
    } else {
      dGammaTab.alloc(std::max<size_t>(dGammaTab.n, (size_t)K), false);
      cuda_check(cudaMemcpyAsync(dGammaTab.p, h_gamma_tab.data(), sizeof(double2) * K, cudaMemcpyHostToDevice, stream),
                 "H2D gamma table");
      args.gamma_tab = dGammaTab.p;
    }
  }
  ctl.gamma = gamma_state;
  args.ctl_in = ctl;
  const TeamDev &Tl = T;
  const auto hp0 = std::chrono::steady_clock::now();
  if (timed) cuda_check(cudaEventRecord(ev0, stream), "eventRecord");
  args.ext_grad_mask = ext_grad_mask(args, use_grid);
  for (size_t i = 0; i < agents.size(); ++i)
    if ((args.ext_grad_mask >> i) & 1u) {
      Agent *a = agents[i];
      EdgeGradArgs eg{};
      eg.n = a->n;
      eg.build_g = 1;
      eg.rec = a->d_er_rec.p;
      eg.inc_ptr = a->d_inc_ptr.p;
      eg.inc_item = a->d_inc_item.p;
      eg.Xin = a->dX.p;
      eg.inbox = a->d_inbox_reg();
      eg.G = a->dG.p;
      eg.Rg = a->dRg.p;
      eg.RgT = a->dRgT.p;
      eg.partials = a->d_eg_partials.p;
      eg.tmarks = nullptr;
      if (a->eg_profile) {
        cuda_check(cudaMemsetAsync(a->d_eg_marks.p, 0xff, sizeof(unsigned long long), stream), "marks");
        cuda_check(cudaMemsetAsync(a->d_eg_marks.p + 1, 0, sizeof(unsigned long long), stream), "marks");
        eg.tmarks = a->d_eg_marks.p;
      }
      cuda_check(launch_edge_grad(eg, a->r, stream), "launch k_edge_grad");
      args.ext_partials[i] = a->d_eg_partials.p;
      args.ext_grid[i] = edge_grad_grid(a->n);
      ++launches;
      // Z^T = (Rg Pinv)^T with one triangle of the dense inverse streamed (sym_precond.cu).  OPT-IN
      // (DPGO_B200_SYM_PRECOND=1): correct to 1e-12 and half the HBM bytes, but on a B200 it runs at 1.8 TB/s against
      // 3.8 TB/s for the full pass inside the persistent kernel (5.57 vs 5.27 ms per step on config 5)
      const bool use_sym = getenv("DPGO_B200_SYM_PRECOND") != nullptr;
      args.ext_zt[i] = nullptr;
      if (use_sym) {
        a->ensure_sym_buffers();
        SymPrecondArgs sp{};
        sp.P = a->dPinv.p;
        sp.ld = roundup32((size_t)4 * a->n);
        sp.n4 = 4 * a->n;
        sp.VT = a->dRgT.p;
        sp.Zt = a->dZt.p;
        sp.partials = a->d_sym_partials.p;
        sp.first_tile = a->d_sym_first_tile.p;
        sp.npanels = a->sym_npanels;
        sp.ntiles = a->sym_ntiles;
        sp.tile_counter = a->d_sym_counter.p;
        cuda_check(launch_sym_precond(sp, a->r, max_coop_grid(device), stream), "launch k_sym_precond");
        args.ext_zt[i] = a->dZt.p;
        launches += 2;
      }
    }
  if ((args.force_selected == -1 || args.mode == 1) && args.max_iters == 1 && args.mode != 2)
    cuda_check(launch_nesterov_only(Tl, args, use_grid, stream), "launch k_nesterov_only");
  else
    cuda_check(launch_team_run(Tl, args, use_grid, stream), "launch k_team_run");
  if (timed) cuda_check(cudaEventRecord(ev1, stream), "eventRecord");
  ++launches;
  pl.args = args;
  pl.la = la;
  pl.la_depth = la_depth;
  pl.timed = timed;
  pl.hp0 = hp0;
}

void Team::launch_finish(PendingLaunch &pl, float *ms) {
  const RunArgs &args = pl.args;
  Agent *la = pl.la;
  const int la_depth = pl.la_depth;
  const bool timed = pl.timed;
  const auto hp0 = pl.hp0;
  const dpgo_b200_params &P = agents[0]->P;
  if (timed) {
    cuda_check(cudaEventSynchronize(ev1), "k_team_run");
    if (ms) cudaEventElapsedTime(ms, ev0, ev1);
  }
  const auto hp1 = std::chrono::steady_clock::now();
  wait_result(args.seq);
  const auto hp2 = std::chrono::steady_clock::now();
  read_back();
  host_prof[0] += std::chrono::duration<double>(hp1 - hp0).count();  // launch call
  host_prof[1] += std::chrono::duration<double>(hp2 - hp1).count();  // launch returned -> result visible
  host_prof[2] += 1.0;
  if (P.acceleration && ctl.iters_done > 0 && args.mode != 2) gamma_state = h_gamma_state[ctl.iters_done - 1];
  ctl.gamma = gamma_state;
  if (la) {
    la->drop_lookahead();
    if (la_depth > 0 && ctl.iters_done == 1) {
      la->la_valid = la_depth;
      la->la_launch = args.seq;
      for (int j = 0; j < la_depth; ++j) {
        la->la_gamma[j] = h_gamma_state[1 + j];
        la->la_restart[j] = args.la_tab[j].y != 0.0;
      }
    }
  }
}

// completion: poll the sequence number the kernel publishes (after a system-scope fence) in the
// mapped result block -- cheaper than a stream synchronisation for 10-us kernels
void Team::wait_result(unsigned long long expect) {
  volatile TeamCtl *c = reinterpret_cast<volatile TeamCtl *>(h_result);
  unsigned spins = 0;
  while (c->seq != expect) {
    if ((++spins & 0x3fff) == 0) {
      const cudaError_t q = cudaStreamQuery(stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) cuda_check(q, "k_team_run");
      if (q == cudaSuccess && c->seq != expect) fail(DPGO_B200_ERR_CUDA, "kernel finished without publishing its result");
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
}

void Team::run_forced(int sel_local) {
  prepare(false, true);  // consumed lookahead steps are committed by this very launch
  RunArgs args{};
  args.max_iters = 1;
  args.force_selected = sel_local;
  // poses staged by updateNeighborPoses since the last solve: the solve kernel pulls them from the pinned host
  // block itself (one coalesced PCIe read hidden behind its first phase) instead of a cudaMemcpyAsync per call
  if (sel_local >= 0)
    for (size_t i = 0; i < agents.size(); ++i)
      if (agents[i]->inbox_dirty && agents[i]->d_inbox.n && agents[i]->h_inbox) {
        args.pull_mask |= 1u << i;
        agents[i]->inbox_dirty = false;
      }
  // iterate(true) of a stand-alone agent returns as soon as X+ is published; fOpt / gradNormOpt of
  // mLocalOptResult cost a second gradient pass and are evaluated when somebody reads them
  args.skip_stats = 1;
  static const bool time_it = getenv("DPGO_B200_TIME_LAUNCHES") != nullptr;  // diagnostics
  float ms = 0;
  launch_and_read(args, sel_local < 0 ? small_grid : grid, time_it, &ms);
  host_prof[3] += ms * 1e-3;
  if (sel_local >= 0 && agents[sel_local]->P.method == 1) {
    Agent *a = agents[sel_local];
    a->stats_pending = true;
    a->opt.f_opt = a->opt.gradnorm_opt = std::nan("");
  }
}

// One global iteration for the LOCAL agents of a team that holds only part of the robots
// (multi-GPU): mode 0 = whole iterate; mode 1 = Nesterov half (everyone's Y, non-selected X = Y,
// publication) ; mode 2 = the selected robot's local solve, after the neighbours' poses of this
// iteration arrived (the gate of src/PGOAgentROS.cpp:136-149).
void Team::step(int selected_robot, int mode) {
  prepare();
  const dpgo_b200_params &P = agents[0]->P;
  int sel_local = -1;
  for (size_t i = 0; i < agents.size(); ++i)
    if (agents[i]->id == selected_robot) sel_local = (int)i;
  if (mode == 2 && sel_local < 0) return;
  if (!P.acceleration && sel_local < 0) {  // plain RBCD: iterate(false) only counts
    if (mode != 2) {
      ctl.iter++;
      if (P.cost_type != 0) ctl.robust_inner_iter++;
      for (Agent *a : agents) {
        a->iter = ctl.iter;
        a->robust_inner_iter = ctl.robust_inner_iter;
      }
    }
    return;
  }
  if (!P.acceleration && mode == 1) {  // nothing to do before the exchange; the solve half counts the iteration
    return;
  }
  RunArgs args{};
  args.max_iters = 1;
  args.force_selected = sel_local;
  args.mode = P.acceleration ? mode : 0;
  const bool small = (args.mode == 1) || sel_local < 0;
  launch_and_read(args, small ? small_grid : grid, false, nullptr);
}

int Team::parallel_schedule_checked() const {
  if (schedule == 0) return 0;
  const dpgo_b200_params &P = agents[0]->P;
  if (P.method != 1 || P.acceleration || P.cost_type != 0)
    fail(DPGO_B200_ERR_INVALID,
         "the parallel (asynchronous-mode) schedule runs RGD without acceleration on the L2 cost "
         "(src/PGOAgentROSNode.cpp:80-93)");
  return 1;
}

dpgo_b200_run_result Team::run(int max_iters, bool stop_on_terminate) {
  dpgo_b200_run_result res{};
  prepare();
  for (Agent *a : agents) {
    if (a->state != 2) fail(DPGO_B200_ERR_STATE, "team_run: every agent must be initialized in the global frame");
    if (T.local_of_robot[a->id] < 0) fail(DPGO_B200_ERR_STATE, "team_run: inconsistent team");
  }
  if ((int)agents.size() != T.num_robots)
    fail(DPGO_B200_ERR_STATE, "team_run needs every robot of the problem in the team (use iterate + exchange otherwise)");
  for (Agent *a : agents)
    if (!a->all_inbox_valid(false) || (a->P.acceleration && !a->all_inbox_valid(true)))
      fail(DPGO_B200_ERR_MISSING, "team_run: neighbour poses missing; call team_exchange_all first");
  int remaining = max_iters;
  const bool eg_loop = edge_grad_loop(grid);
  while (remaining > 0) {
    RunArgs args{};
    // LARGE agents: one iteration per launch, the gradient in k_edge_grad in front of it; the statistics of
    // mLocalOptResult (a second gradient pass) only with the last iteration of this call
    args.max_iters = eg_loop ? 1 : std::min(remaining, 1 << 16);
    args.skip_stats = (eg_loop && remaining > 1) ? 1 : 0;
    args.force_selected = -2;
    args.stop_on_terminate = stop_on_terminate ? 1 : 0;
    args.parallel = parallel_schedule_checked();
    float ms = 0;
    const int l0 = launches;
    launch_and_read(args, grid, true, &ms);
    res.device_ms += ms;
    res.kernel_launches += launches - l0;
    res.iterations += ctl.iters_done;
    remaining -= ctl.iters_done;
    if (ctl.stop_reason == 2) {
      gnc_update_all();
      res.weight_updates++;
      continue;
    }
    if (ctl.stop_reason == 1) {
      res.terminated = 1;
      if (stop_on_terminate) break;
    }
    if (ctl.iters_done == 0) break;   // nothing ran (cannot happen with max_iters >= 1; guards the loop)
  }
  res.stop_reason = ctl.stop_reason;
  return res;
}

void Team::exchange_all() {
  prepare();
  // publish X (and Y) of every agent through the same publication lists the
  // persistent kernel uses: a forced iteration count of zero does nothing, so
  // use the dedicated kernel
  cuda_check(launch_publish_all(T, grid, stream), "publish_all");
  cuda_check(cudaStreamSynchronize(stream), "publish_all");
  for (Agent *a : agents) {
    if (a->state != 2) continue;
    a->outbox_stale = false;
    a->outbox_mirror_valid = true;
    for (Agent *o : agents) {
      if (o == a) continue;
      for (size_t s = 0; s < o->slot_key.size(); ++s)
        if (o->slot_key[s].first == a->id) {
          o->inbox_valid_reg[s] = 1;
          if (a->P.acceleration) o->inbox_valid_aux[s] = 1;
        }
    }
  }
}

// UPDATE_WEIGHT (src/PGOAgentROS.cpp:1211-1233): residual + GNC-TLS weight on the
// device, ownership rule and Q / G / preconditioner rebuild on the host side.
void Team::gnc_update_all() {
  gnc_compute_weights();
  gnc_finish_update();
}

void Team::gnc_compute_weights() {
  prepare();
  for (Agent *a : agents) {
    if (a->state != 2) continue;
    const size_t L = a->lc_meas.size();
    if (L) {
      // residual + GNC-TLS weight on the device; the weights this agent owns are rewritten in MeasDev::w, which is
      // what k_assemble_values reads when gnc_finish_update marks the data matrices stale
      ResidualJob J{(int)L, a->d_lc_meas.p, a->d_lc_mask.p, a->d_lc_residual.p, a->P.gnc_barc * a->P.gnc_barc, a->mu,
                    a->P.cost_type};
      cuda_check(launch_measurement_residuals(a->meas_view(), J, a->r, a->dX.p, a->d_inbox_reg(), stream), "gnc_weights");
      // host mirror (get_lc_weights, the owner -> neighbour weight messages): owned entries only
      std::vector<double> w(a->num_meas());
      cuda_check(cudaMemcpyAsync(w.data(), a->d_m_w.p, w.size() * sizeof(double), cudaMemcpyDeviceToHost, stream),
                 "D2H weights");
      cuda_check(cudaStreamSynchronize(stream), "gnc_weights");
      for (size_t e = 0; e < L; ++e)
        if (a->lc_mask[e]) a->meas_at(a->lc_meas[e]).weight = w[a->lc_meas[e]];
    }
  }
  // publishMeasurementWeights (:721-754) -> measurementWeightsCallback (:1315-1353)
  for (Agent *a : agents)
    for (const auto &m : a->slc) {
      const int other = (m.r1 == a->id) ? m.r2 : m.r1;
      if (other > a->id)
        for (Agent *o : agents)
          if (o->id == other) {
            Meas *mm = o->find_measurement(m.r1, m.p1, m.r2, m.p2);
            if (mm) {
              if (mm->fixed != m.fixed) o->lc_dirty = true;
              mm->weight = m.weight;
              mm->fixed = m.fixed;
              o->weights_host_dirty = true;
            }
          }
    }
}

void Team::gnc_finish_update() {
  for (Agent *a : agents) {
    if (a->state != 2) continue;
    if (a->P.cost_type == 5) a->mu *= a->P.gnc_mu_step;
    a->weight_update_count++;
    a->robust_inner_iter = 0;
    a->values_dirty = a->precon_dirty = true;
    a->resid_valid = false;
    const size_t bytes = (size_t)a->r * 4 * a->n * sizeof(double);
    if (a->weight_update_count <= a->P.robust_opt_num_resets)
      cuda_check(cudaMemcpy(a->dX.p, a->dXinit.p, bytes, cudaMemcpyDeviceToDevice), "reset X");
    if (a->P.acceleration) {
      cuda_check(cudaMemcpy(a->dV.p, a->dX.p, bytes, cudaMemcpyDeviceToDevice), "V = X");
      cuda_check(cudaMemcpy(a->dY.p, a->dX.p, bytes, cudaMemcpyDeviceToDevice), "Y = X");
    }
  }
  ctl.weight_update_count++;
  ctl.robust_inner_iter = 0;
  if (agents[0]->P.acceleration) {
    ctl.gamma = ctl.alpha = 0;
    gamma_state = 0;
  }
  team_dirty = true;
  exchange_all();
}

double Team::global_cost() {
  prepare();
  std::map<int, std::vector<double>> X;
  for (Agent *a : agents) {
    std::vector<double> h((size_t)a->r * 4 * a->n);
    cuda_check(cudaMemcpy(h.data(), a->dX.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost), "D2H X");
    X[a->id] = std::move(h);
  }
  const int r = agents[0]->r;
  auto edge_cost = [&](const Meas &m, const double *Xi, const double *Xj) {
    double rot = 0, tr = 0;
    for (int c = 0; c < 3; ++c)
      for (int a = 0; a < r; ++a) {
        double s = -Xj[(size_t)c * r + a];
        for (int k = 0; k < 3; ++k) s += Xi[(size_t)k * r + a] * m.R[c * 3 + k];
        rot += s * s;
      }
    for (int a = 0; a < r; ++a) {
      double s = Xj[(size_t)3 * r + a] - Xi[(size_t)3 * r + a];
      for (int k = 0; k < 3; ++k) s -= Xi[(size_t)k * r + a] * m.t[k];
      tr += s * s;
    }
    return m.weight * (m.kappa * rot + m.tau * tr);
  };
  double cost = 0;
  for (Agent *a : agents) {
    const double *Xa = X[a->id].data();
    for (auto *vec : {&a->odom, &a->plc})
      for (const auto &m : *vec) cost += edge_cost(m, Xa + (size_t)m.p1 * 4 * r, Xa + (size_t)m.p2 * 4 * r);
    for (const auto &m : a->slc)
      if (m.r1 == a->id && X.count(m.r2))
        cost += edge_cost(m, Xa + (size_t)m.p1 * 4 * r, X[m.r2].data() + (size_t)m.p2 * 4 * r);
  }
  return cost;
}

}  // namespace dpgo
