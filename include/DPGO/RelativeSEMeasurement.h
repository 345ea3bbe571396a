/* DPGO/RelativeSEMeasurement.h -- the measurement type crossing the boundary
 * (8-argument ctor and field names pinned by tests/testUtils.cpp:43-52; weight / fixedWeight
 * by src/utils.cpp:144-149). */
#ifndef DPGO_SHIM_RELATIVESEMEASUREMENT_H
#define DPGO_SHIM_RELATIVESEMEASUREMENT_H
#include "DPGO/DPGO_types.h"

namespace DPGO {
struct RelativeSEMeasurement {
  size_t r1 = 0, r2 = 0, p1 = 0, p2 = 0;  // source / destination robot and pose index
  Matrix R;                               // d x d
  Matrix t;                               // d x 1
  double kappa = 0, tau = 0;              // rotation / translation precision
  double weight = 1.0;                    // GNC weight
  bool fixedWeight = false;               // excluded from reweighting (odometry, src/utils.cpp:147-149)
  RelativeSEMeasurement() = default;
  RelativeSEMeasurement(size_t first_robot, size_t second_robot, size_t first_pose, size_t second_pose,
                        const Matrix &relative_rotation, const Matrix &relative_translation,
                        double rotational_precision, double translational_precision)
      : r1(first_robot), r2(second_robot), p1(first_pose), p2(second_pose), R(relative_rotation),
        t(relative_translation), kappa(rotational_precision), tau(translational_precision) {}
};
}  // namespace DPGO
#endif
