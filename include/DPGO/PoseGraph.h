/*
 * DPGO/PoseGraph.h -- host-side mirror of the agent's pose graph: the wrapper reads counts and
 * measurement ranges from it and REPLACES it (`mPoseGraph = std::make_shared<PoseGraph>(mID, r, d)`,
 * src/PGOAgentROS.cpp:237), so it must stay a plain shared_ptr-managed object with an (id, r, d)
 * constructor.  The data matrices themselves (Q, G, the preconditioner) live on the GPU;
 * clearDataMatrices() (:1351) forwards through the owning agent.
 */
#ifndef DPGO_SHIM_POSEGRAPH_H
#define DPGO_SHIM_POSEGRAPH_H
#include <algorithm>
#include <functional>
#include <set>
#include <vector>

#include "DPGO/DPGO_types.h"
#include "DPGO/RelativeSEMeasurement.h"

namespace DPGO {
class PoseGraph {
 public:
  struct Statistics {
    double total_loop_closures = 0, accept_loop_closures = 0, reject_loop_closures = 0, undecided_loop_closures = 0;
  };
  PoseGraph(unsigned int id, unsigned int r, unsigned int d) : id_(id), r_(r), d_(d) {}
  unsigned int r() const { return r_; }
  unsigned int d() const { return d_; }
  unsigned int n() const { return n_; }
  unsigned int numOdometry() const { return (unsigned)odometry_.size(); }                     // :343, :1292
  unsigned int numPrivateLoopClosures() const { return (unsigned)private_lcs_.size(); }       // :344
  unsigned int numSharedLoopClosures() const { return (unsigned)shared_lcs_.size(); }         // :345
  unsigned int numMeasurements() const { return numOdometry() + numPrivateLoopClosures() + numSharedLoopClosures(); }
  bool hasMeasurement(const PoseID &src, const PoseID &dst) const {                           // :276
    return edges_.count({{src.robot_id, src.frame_id}, {dst.robot_id, dst.frame_id}}) != 0;
  }
  // returns false when the measurement does not involve this robot or is a duplicate
  bool addMeasurement(const RelativeSEMeasurement &m) {
    if (m.r1 != id_ && m.r2 != id_) return false;
    const auto key = std::make_pair(std::make_pair((unsigned)m.r1, (unsigned)m.p1), std::make_pair((unsigned)m.r2, (unsigned)m.p2));
    if (!edges_.insert(key).second) return false;
    if (m.r1 == id_ && m.r2 == id_) {
      (m.p1 + 1 == m.p2 ? odometry_ : private_lcs_).push_back(m);
      n_ = std::max<unsigned>(n_, (unsigned)std::max(m.p1, m.p2) + 1);
    } else {
      shared_lcs_.push_back(m);
      if (m.r1 == id_) {
        n_ = std::max<unsigned>(n_, (unsigned)m.p1 + 1);
        nbr_ids_.insert((unsigned)m.r2);
        nbr_pose_ids_.insert(PoseID((unsigned)m.r2, (unsigned)m.p2));
      } else {
        n_ = std::max<unsigned>(n_, (unsigned)m.p2 + 1);
        nbr_ids_.insert((unsigned)m.r1);
        nbr_pose_ids_.insert(PoseID((unsigned)m.r1, (unsigned)m.p1));
      }
    }
    return true;
  }
  std::vector<RelativeSEMeasurement> &odometry() { return odometry_; }
  std::vector<RelativeSEMeasurement> &privateLoopClosures() { return private_lcs_; }           // :770
  std::vector<RelativeSEMeasurement> &sharedLoopClosures() { return shared_lcs_; }            // :706, :725, :800
  const std::vector<RelativeSEMeasurement> &sharedLoopClosures() const { return shared_lcs_; }
  // loop closures with / without an active neighbour (all robots are active on the GPU fabric)
  std::vector<RelativeSEMeasurement *> activeLoopClosures() {                                 // :1048, :1431
    std::vector<RelativeSEMeasurement *> v;
    for (auto &m : private_lcs_) v.push_back(&m);
    for (auto &m : shared_lcs_)
      if (active_(m.r1 == id_ ? (unsigned)m.r2 : (unsigned)m.r1)) v.push_back(&m);
    return v;
  }
  std::vector<RelativeSEMeasurement *> inactiveLoopClosures() {                               // :1445
    std::vector<RelativeSEMeasurement *> v;
    for (auto &m : shared_lcs_)
      if (!active_(m.r1 == id_ ? (unsigned)m.r2 : (unsigned)m.r1)) v.push_back(&m);
    return v;
  }
  std::set<unsigned> activeNeighborIDs() const {                                              // :137
    std::set<unsigned> s;
    for (unsigned b : nbr_ids_)
      if (active_(b)) s.insert(b);
    return s;
  }
  const std::set<unsigned> &neighborIDs() const { return nbr_ids_; }
  std::set<PoseID, ComparePoseID> activeNeighborPublicPoseIDs() const {                       // :1394
    std::set<PoseID, ComparePoseID> s;
    for (const auto &p : nbr_pose_ids_)
      if (active_(p.robot_id)) s.insert(p);
    return s;
  }
  RelativeSEMeasurement *findMeasurement(const PoseID &src, const PoseID &dst) {
    for (auto *vec : {&odometry_, &private_lcs_, &shared_lcs_})
      for (auto &m : *vec)
        if (m.r1 == src.robot_id && m.p1 == src.frame_id && m.r2 == dst.robot_id && m.p2 == dst.frame_id) return &m;
    return nullptr;
  }
  Statistics statistics() const {                                                             // :1058-1067
    Statistics st;
    for (auto *vec : {&private_lcs_, &shared_lcs_})
      for (const auto &m : *vec) {
        // fixed weights count too: the wrapper prints these numbers right after it has turned its rejects into
        // (weight 0, fixedWeight) (src/PGOAgentROS.cpp:1051-1067)
        st.total_loop_closures += 1;
        if (m.weight == 1.0)
          st.accept_loop_closures += 1;
        else if (m.weight == 0.0)
          st.reject_loop_closures += 1;
        else
          st.undecided_loop_closures += 1;
      }
    return st;
  }
  void clearDataMatrices() {                                                                  // :1351
    if (on_clear_) on_clear_();
  }
  void useInactiveNeighbors(bool) {}                                                          // (:156, commented out upstream)
  void setNeighborActive(unsigned id, bool active) { (active ? inactive_.erase(id) : (inactive_.insert(id), size_t(0))); }
  void bindClear(std::function<void()> fn) { on_clear_ = std::move(fn); }

 private:
  bool active_(unsigned id) const { return inactive_.count(id) == 0; }
  unsigned int id_, r_, d_, n_ = 0;
  std::vector<RelativeSEMeasurement> odometry_, private_lcs_, shared_lcs_;
  std::set<std::pair<std::pair<unsigned, unsigned>, std::pair<unsigned, unsigned>>> edges_;
  std::set<unsigned> nbr_ids_, inactive_;
  std::set<PoseID, ComparePoseID> nbr_pose_ids_;
  std::function<void()> on_clear_;
};
}  // namespace DPGO
#endif
