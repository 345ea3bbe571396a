/*
 * DPGO/DPGO_utils.h -- free functions of the DPGO:: surface used by dpgo_ros:
 * read_g2o_file (src/PGODatasetPublisherNode.cpp:80), plus the small dense helpers the
 * shim itself needs for the global-frame read-out (projection to SO(3), chi-square
 * quantile for RobustCost::computeErrorThresholdAtQuantile, src/PGOAgentROSNode.cpp:201).
 * Host-side, not on the hot path.
 */
#ifndef DPGO_SHIM_UTILS_H
#define DPGO_SHIM_UTILS_H

#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "DPGO/DPGO_types.h"
#include "DPGO/RelativeSEMeasurement.h"

namespace DPGO {

namespace detail {
// inverse of a symmetric positive definite 3x3 block
inline bool inv3(const double (&A)[3][3], double (&B)[3][3]) {
  const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                     A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
  if (det == 0.0) return false;
  const double s = 1.0 / det;
  B[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) * s;
  B[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * s;
  B[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * s;
  B[1][0] = (A[1][2] * A[2][0] - A[1][0] * A[2][2]) * s;
  B[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * s;
  B[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * s;
  B[2][0] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) * s;
  B[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * s;
  B[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * s;
  return true;
}
}  // namespace detail

// 3-D g2o (EDGE_SE3:QUAT) -> measurements with robot id 0 and global pose indices.  Precisions follow the SE-Sync
// rule: tau = 3 / tr(I_t^-1), kappa = 3 / (2 tr(I_R^-1)).
inline std::vector<RelativeSEMeasurement> read_g2o_file(const std::string &filename, size_t &num_poses) {
  std::vector<RelativeSEMeasurement> out;
  std::ifstream in(filename);
  if (!in) throw std::runtime_error("read_g2o_file: cannot open " + filename);
  std::string line, tag;
  num_poses = 0;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    ss >> tag;
    if (tag != "EDGE_SE3:QUAT") continue;
    size_t i, j;
    double t[3], q[4], info[21];
    ss >> i >> j >> t[0] >> t[1] >> t[2] >> q[0] >> q[1] >> q[2] >> q[3];
    for (double &v : info) ss >> v;
    double I[6][6];
    int k = 0;
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) I[a][b] = I[b][a] = info[k++];
    double It[3][3], Ir[3][3], Ct[3][3], Cr[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        It[a][b] = I[a][b];
        Ir[a][b] = I[3 + a][3 + b];
      }
    if (!detail::inv3(It, Ct) || !detail::inv3(Ir, Cr)) throw std::runtime_error("read_g2o_file: singular information");
    const double tau = 3.0 / (Ct[0][0] + Ct[1][1] + Ct[2][2]);
    const double kappa = 3.0 / (2.0 * (Cr[0][0] + Cr[1][1] + Cr[2][2]));
    const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double x = q[0] / nq, y = q[1] / nq, z = q[2] / nq, w = q[3] / nq;
    Matrix R(3, 3), tv(3, 1);
    R << 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y);
    tv << t[0], t[1], t[2];
    out.emplace_back(0, 0, i, j, R, tv, kappa, tau);
    num_poses = std::max(num_poses, std::max(i, j) + 1);
  }
  return out;
}

// nearest rotation to a 3x3 matrix (polar factor with det +1) by Newton iteration on the polar
// decomposition; inputs here are Y_a^T Y_i of nearly orthonormal blocks, so it converges in a few steps
inline Matrix projectToRotationGroup(const Matrix &M) {
  assert(M.rows() == 3 && M.cols() == 3);
  Matrix X = M;
  for (int it = 0; it < 100; ++it) {
    // X <- (X + X^-T) / 2
    double A[3][3], B[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i][j] = X(i, j);
    if (!detail::inv3(A, B)) break;
    double diff = 0;
    Matrix N(3, 3);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        N(i, j) = 0.5 * (X(i, j) + B[j][i]);
        diff += (N(i, j) - X(i, j)) * (N(i, j) - X(i, j));
      }
    X = N;
    if (diff < 1e-30) break;
  }
  const double det = X(0, 0) * (X(1, 1) * X(2, 2) - X(1, 2) * X(2, 1)) - X(0, 1) * (X(1, 0) * X(2, 2) - X(1, 2) * X(2, 0)) +
                     X(0, 2) * (X(1, 0) * X(2, 1) - X(1, 1) * X(2, 0));
  if (det < 0)
    for (int i = 0; i < 3; ++i) X(i, 2) = -X(i, 2);  // (only reached for reflected input; flips the weakest axis choice)
  return X;
}

// Robust single-transform averaging used by the cross-robot initialisation (multirobot_initialization,
// src/PGOAgentROSNode.cpp:120; robustInitMinInliers :220): every shared loop closure with an initialised neighbour
// votes for this robot's world-frame transform; the candidate with the largest consensus set (rotation within rotTol
// in Frobenius norm, translation within tranTol) wins and its consensus set is chordal-averaged.  Returns false when
// the best consensus set has fewer than minInliers members.  Candidates are 3 x 4 [R | t].
inline bool robustTransformAverage(const std::vector<Matrix> &candidates, double rotTol, double tranTol,
                                   unsigned minInliers, Matrix &out, unsigned *numInliers = nullptr) {
  size_t best = 0, bestCount = 0;
  std::vector<size_t> bestSet;
  for (size_t c = 0; c < candidates.size(); ++c) {
    std::vector<size_t> set;
    for (size_t k = 0; k < candidates.size(); ++k) {
      const Matrix D = candidates[k] - candidates[c];
      if (D.block(0, 0, 3, 3).norm() < rotTol && D.block(0, 3, 3, 1).norm() < tranTol) set.push_back(k);
    }
    if (set.size() > bestCount) {
      best = c;
      bestCount = set.size();
      bestSet = set;
    }
  }
  (void)best;
  if (numInliers) *numInliers = (unsigned)bestCount;
  if (bestCount == 0 || bestCount < minInliers) return false;
  Matrix mean(3, 4);
  for (size_t k : bestSet) mean = mean + candidates[k];
  mean = (1.0 / (double)bestCount) * mean;
  out = Matrix(3, 4);
  out.block(0, 0, 3, 3) = projectToRotationGroup(mean.block(0, 0, 3, 3));
  out.block(0, 3, 3, 1) = mean.block(0, 3, 3, 1);
  return true;
}
// rigid transforms as 3 x 4 [R | t]
inline Matrix se3Compose(const Matrix &A, const Matrix &B) {
  Matrix C(3, 4);
  C.block(0, 0, 3, 3) = A.block(0, 0, 3, 3) * B.block(0, 0, 3, 3);
  C.block(0, 3, 3, 1) = A.block(0, 0, 3, 3) * B.block(0, 3, 3, 1) + A.block(0, 3, 3, 1);
  return C;
}
inline Matrix se3Inverse(const Matrix &A) {
  Matrix C(3, 4);
  const Matrix Rt = A.block(0, 0, 3, 3).transpose();
  C.block(0, 0, 3, 3) = Rt;
  C.block(0, 3, 3, 1) = -1.0 * (Rt * A.block(0, 3, 3, 1));
  return C;
}

// regularised lower incomplete gamma P(a, x) and the chi-square quantile built on it
inline double gammaP(double a, double x) {
  if (x <= 0) return 0.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {  // series
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 1000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-16) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  double b = x + 1.0 - a, c = 1e300, d = 1.0 / b, h = d;  // continued fraction for Q
  for (int i = 1; i < 1000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < 1e-300) d = 1e-300;
    c = b + an / c;
    if (std::fabs(c) < 1e-300) c = 1e-300;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}
inline double chi2inv(double quantile, size_t dof) {
  double lo = 0.0, hi = 1.0;
  while (gammaP(0.5 * dof, 0.5 * hi) < quantile) hi *= 2.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    (gammaP(0.5 * dof, 0.5 * mid) < quantile ? lo : hi) = mid;
  }
  return 0.5 * (lo + hi);
}

// a fixed YLift in St(d, r) shared by every robot (the wrapper broadcasts the leader's, src/PGOAgentROS.cpp:402-410;
// this is the deterministic matrix the repo's tests use: Gram-Schmidt of a closed-form full-rank r x d matrix)
inline Matrix fixedStiefelVariable(unsigned d, unsigned r) {
  Matrix A(r, d);
  for (unsigned i = 0; i < r; ++i)
    for (unsigned j = 0; j < d; ++j)
      A(i, j) = std::cos(0.7 * (i + 1.0) * (j + 1.0)) + 0.3 * std::sin(1.3 * i - 0.4 * j) + (i == j ? 1.0 : 0.0);
  for (unsigned j = 0; j < d; ++j) {  // modified Gram-Schmidt == QR with diag(R) > 0
    for (unsigned k = 0; k < j; ++k) {
      double dot = 0;
      for (unsigned i = 0; i < r; ++i) dot += A(i, k) * A(i, j);
      for (unsigned i = 0; i < r; ++i) A(i, j) -= dot * A(i, k);
    }
    double nrm = 0;
    for (unsigned i = 0; i < r; ++i) nrm += A(i, j) * A(i, j);
    nrm = std::sqrt(nrm);
    for (unsigned i = 0; i < r; ++i) A(i, j) /= nrm;
  }
  return A;
}

}  // namespace DPGO
#endif
