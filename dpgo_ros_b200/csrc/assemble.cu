// Data-matrix assembly on the device and the per-measurement residual / GNC-TLS weight kernel (a8).
//
// The pose graph lives on the device as ONE structure-of-arrays over the agent's measurements, in the order
// [odometry | private loop closures | shared loop closures] (MeasDev).  Everything that depends on the
// measurement weights is produced from it by k_assemble_values:
//   * the 4x4 blocks of Q  (PoseGraph::quadraticMatrix; reference call sites: addMeasurement
//     src/PGOAgentROS.cpp:277,1307, clearDataMatrices :1351 after setMeasurementWeight :1341):
//       private edge (i -> j):  Q_ii += T Om T^T,  Q_jj += Om,  Q_ij = -T Om,  Q_ji = -(T Om)^T
//       shared edge:            only this agent's diagonal block
//     with T = [R t; 0 1], Om = w diag(kappa, kappa, kappa, tau);
//   * the 4x4 blocks of the linear term (PoseGraph::linearMatrix), one per shared edge:
//       outgoing  G_i -= X_j^nbr (Om T^T),   incoming  G_j -= X_i^nbr (T Om).
// Accumulation is segmented BY DESTINATION SLOT (block-CSR by output pose): the host lists, once per graph
// change, which (measurement, role) pairs land in which slot; sixteen threads own the sixteen entries of a slot
// and add the contributions in list order -- no atomics, bitwise reproducible, and the same kernel writes the
// CSR copy (read by the dense-inverse scatter) and the ELL / overflow copy (read by the hot phases).  A GNC
// weight update therefore costs one small kernel + the dense inverse, with no host assembly and no upload of Q.
#include "kernels.h"

namespace dpgo {

void count_launch();

// entry (i, j) of T = [R t; 0 1]; R column-major 3x3
__device__ __forceinline__ double t_entry(const double *R, const double *t, int i, int j) {
  if (i < 3) return j < 3 ? R[j * 3 + i] : t[i];
  return j < 3 ? 0.0 : 1.0;
}

// entry (i, j) of the contribution of measurement m in `role` to a Q block
//   0: T Om T^T   1: Om   2: -T Om   3: -(T Om)^T
__device__ __forceinline__ double q_contrib(const MeasDev &M, int m, int role, int i, int j) {
  const double w = M.skip[m] ? 0.0 : M.w[m];
  const double k = w * M.kappa[m], tau = w * M.tau[m];
  const double *R = M.R + (size_t)m * 9, *t = M.t + (size_t)m * 3;
  if (role == 1) return i == j ? (i < 3 ? k : tau) : 0.0;
  if (role == 2) return -(t_entry(R, t, i, j) * (j < 3 ? k : tau));
  if (role == 3) return -(t_entry(R, t, j, i) * (i < 3 ? k : tau));
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) s += (t_entry(R, t, i, q) * (q < 3 ? k : tau)) * t_entry(R, t, j, q);
  return s;
}

__device__ __forceinline__ void store_block_entry(double v, int slot, int tid16, int dst, double *csr, double *ell,
                                                  double *ovf) {
  if (csr) csr[(size_t)slot * 16 + tid16] = v;
  if (dst >= 0)
    ell[(size_t)dst * 16 + tid16] = v;
  else
    ovf[(size_t)(-dst - 1) * 16 + tid16] = v;
}

// one 16-thread group per destination slot: first the nq slots of Q, then the ns blocks of the linear term
__global__ void __launch_bounds__(256) k_assemble_values(const __grid_constant__ MeasDev M,
                                                         const __grid_constant__ AssembleDev A) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, tid16 = threadIdx.x & 15;
  const int i = tid16 & 3, j = tid16 >> 2;  // blocks are column-major
  if (g < A.nq) {
    double s = 0.0;
    for (int c = A.qc_ptr[g]; c < A.qc_ptr[g + 1]; ++c) {
      const int item = A.qc_item[c];
      s += q_contrib(M, item >> 2, item & 3, i, j);
    }
    store_block_entry(s, g, tid16, A.q_dst[g], A.q_val, A.qe_val, A.qo_val);
  } else if (g < A.nq + A.ns) {
    const int e = g - A.nq;
    const int item = A.s_item[e], m = item >> 1;
    // outgoing (this agent is the source): -(Om T^T)(i, j) = -Om_i T(j, i); incoming: -(T Om)(i, j)
    const double v = q_contrib(M, m, (item & 1) ? 2 : 3, i, j);
    store_block_entry(v, e, tid16, A.s_dst[e], A.s_val, A.se_val, A.so_val);
  }
}

// computeMeasurementResidual (src/PGOAgentROS.cpp:1049) for a list of measurements, optionally followed by the
// GNC-TLS weight (RobustCost::weight, :1050) written back into the measurement's weight -- the weight update
// stays on the device (k_assemble_values reads it next).
//   residual = sqrt(kappa |Y_j - Y_i R|^2 + tau |p_j - p_i - Y_i t|^2)
__global__ void k_measurement_residuals(MeasDev M, ResidualJob J, int r, const double *X, const double *inbox) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= J.count) return;
  const int m = J.meas ? J.meas[e] : e;
  const unsigned char fl = M.flags[m];
  const double *Xi = ((fl & 1) ? inbox : X) + (size_t)M.src[m] * 4 * r;
  const double *Xj = ((fl & 2) ? inbox : X) + (size_t)M.dst[m] * 4 * r;
  const double *Rm = M.R + (size_t)m * 9;  // column-major
  const double *tm = M.t + (size_t)m * 3;
  double rot = 0, tr = 0;
  for (int a = 0; a < r; ++a) {
    const double y0 = Xi[a], y1 = Xi[r + a], y2 = Xi[2 * r + a];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double s = y0 * Rm[c * 3] + y1 * Rm[c * 3 + 1] + y2 * Rm[c * 3 + 2] - Xj[c * r + a];
      rot += s * s;
    }
    const double s = Xj[3 * r + a] - Xi[3 * r + a] - (y0 * tm[0] + y1 * tm[1] + y2 * tm[2]);
    tr += s * s;
  }
  const double rsq = M.kappa[m] * rot + M.tau[m] * tr;
  J.residual[e] = sqrt(rsq);
  if (J.update_mask && J.update_mask[e]) {
    double w = 1.0;
    if (J.cost_type == 5) {
      const double upper = (J.mu + 1.0) / J.mu * J.barc_sq;
      const double lower = J.mu / (J.mu + 1.0) * J.barc_sq;
      if (rsq >= upper)
        w = 0.0;
      else if (rsq <= lower)
        w = 1.0;
      else
        w = sqrt(J.barc_sq * J.mu * (J.mu + 1.0) / rsq) - J.mu;
    }
    M.w[m] = w;
  }
}

cudaError_t launch_assemble_values(const MeasDev &M, const AssembleDev &A, cudaStream_t s) {
  const long long groups = (long long)A.nq + A.ns;
  if (groups == 0) return cudaSuccess;
  count_launch();
  k_assemble_values<<<(unsigned)((groups * 16 + 255) / 256), 256, 0, s>>>(M, A);
  return cudaGetLastError();
}

cudaError_t launch_measurement_residuals(const MeasDev &M, const ResidualJob &J, int r, const double *X,
                                         const double *inbox, cudaStream_t s) {
  if (J.count == 0) return cudaSuccess;
  count_launch();
  k_measurement_residuals<<<(J.count + 127) / 128, 128, 0, s>>>(M, J, r, X, inbox);
  return cudaGetLastError();
}

}  // namespace dpgo
