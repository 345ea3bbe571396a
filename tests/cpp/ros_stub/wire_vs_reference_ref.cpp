// Half of oracle/_ref/dpgo_ros_wire_vs_reference: the REFERENCE's codecs (src/utils.cpp, compiled unmodified by
// oracle/Makefile.ref) behind plain C functions, so that the other half can hold include/dpgo_ros_wire/wire.h -- which
// mirrors the same names in the same namespace -- without the two meeting in one translation unit.  TEST INFRASTRUCTURE.
#include <dpgo_ros/utils.h>

extern "C" {
// MatrixToMsg (src/utils.cpp:51-57): column-major in, the message's rows / cols / row-major values out
void ref_matrix_to_msg(const double *colmajor, int rows, int cols, unsigned *orows, unsigned *ocols, double *values) {
  DPGO::Matrix M(rows, cols);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) M(i, j) = colmajor[(size_t)j * rows + i];
  const dpgo_ros::MatrixMsg msg = dpgo_ros::MatrixToMsg(M);
  *orows = msg.rows;
  *ocols = msg.cols;
  for (size_t k = 0; k < msg.values.size(); ++k) values[k] = msg.values[k];
}
// MatrixFromMsg (src/utils.cpp:59-61)
void ref_matrix_from_msg(int rows, int cols, const double *values, double *colmajor) {
  dpgo_ros::MatrixMsg msg;
  msg.rows = rows;
  msg.cols = cols;
  msg.values.assign(values, values + (size_t)rows * cols);
  const DPGO::Matrix M = dpgo_ros::MatrixFromMsg(msg);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) colmajor[(size_t)j * rows + i] = M(i, j);
}
// statusToMsg / statusFromMsg (src/utils.cpp:262-281): fields of the message, then of the status read back from it
void ref_status_roundtrip(unsigned id, int state, unsigned instance, unsigned iteration, int ready, double rel_change,
                          double *msg_fields /* 6 */, double *back_fields /* 6 */) {
  const DPGO::PGOAgentStatus st(id, static_cast<DPGO::PGOAgentState>(state), instance, iteration, ready != 0, rel_change);
  const dpgo_ros::Status msg = dpgo_ros::statusToMsg(st);
  const double m[6] = {(double)msg.robot_id, (double)msg.state, (double)msg.instance_number, (double)msg.iteration_number,
                       (double)msg.ready_to_terminate, (double)msg.relative_change};
  const DPGO::PGOAgentStatus b = dpgo_ros::statusFromMsg(msg);
  const double q[6] = {(double)b.agentID, (double)b.state, (double)b.instanceNumber, (double)b.iterationNumber,
                       (double)b.readyToTerminate, b.relativeChange};
  for (int k = 0; k < 6; ++k) {
    msg_fields[k] = m[k];
    back_fields[k] = q[k];
  }
}
// RelativeMeasurementToMsg -> RelativeMeasurementFromMsg (src/utils.cpp:108-152): what survives the PoseGraphEdge message
void ref_measurement_roundtrip(const double *R_rowmajor, const double *t, double *R_out_rowmajor, double *t_out,
                               double *kappa_tau_fixed /* 3 */, int odometry) {
  DPGO::Matrix R(3, 3), tv(3, 1);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R(i, j) = R_rowmajor[i * 3 + j];
    tv(i, 0) = t[i];
  }
  const DPGO::RelativeSEMeasurement m(0, odometry ? 0 : 1, 4, 5, R, tv, 3.0, 7.0);
  const DPGO::RelativeSEMeasurement b = dpgo_ros::RelativeMeasurementFromMsg(dpgo_ros::RelativeMeasurementToMsg(m));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R_out_rowmajor[i * 3 + j] = b.R(i, j);
    t_out[i] = b.t(i, 0);
  }
  kappa_tau_fixed[0] = b.kappa;
  kappa_tau_fixed[1] = b.tau;
  kappa_tau_fixed[2] = b.fixedWeight ? 1 : 0;
}
}
