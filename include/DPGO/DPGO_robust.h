/* DPGO/DPGO_robust.h -- RobustCostParameters / RobustCost as dpgo_ros uses them
 * (src/PGOAgentROSNode.cpp:174-221: cost type, GNC parameters, computeErrorThresholdAtQuantile;
 * src/PGOAgentROS.cpp:1050: mRobustCost.weight(residual) in the TERMINATE handler).
 * Only L2 and GNC_TLS are computed by the B200 path; the weight itself comes
 * from the library, whose mu schedule lives with the agent. */
#ifndef DPGO_SHIM_ROBUST_H
#define DPGO_SHIM_ROBUST_H
#include <functional>

#include "DPGO/DPGO_utils.h"

namespace DPGO {
struct RobustCostParameters {
  enum class Type { L2, L1, Huber, TLS, GM, GNC_TLS };
  Type costType = Type::L2;
  unsigned GNCMaxNumIters = 10000;
  double GNCBarc = 5.0, GNCMuStep = 2.0, GNCInitMu = 1e-5;
};
class RobustCost {
 public:
  RobustCost() = default;
  explicit RobustCost(const RobustCostParameters &p) : mParams(p) {}
  // weight of a measurement with residual r under the CURRENT mu (forwarded to dpgo_b200_robust_weight)
  double weight(double r) const { return mWeightFn ? mWeightFn(r) : 1.0; }
  // sqrt of the chi-square quantile: residual threshold below which a measurement is an inlier with probability q
  static double computeErrorThresholdAtQuantile(double quantile, size_t dimension) {
    return std::sqrt(chi2inv(quantile, dimension));
  }
  void bind(std::function<double(double)> fn) { mWeightFn = std::move(fn); }

 private:
  RobustCostParameters mParams;
  std::function<double(double)> mWeightFn;
};
}  // namespace DPGO
#endif
