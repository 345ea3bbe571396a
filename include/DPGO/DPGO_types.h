/*
 * DPGO/DPGO_types.h -- value types of the DPGO:: surface that dpgo_ros consumes
 * (include sites: include/dpgo_ros/utils.h:10-12, src/utils.cpp:8-9).
 *
 * Part of the source-compatible shim that lets `class PGOAgentROS : public PGOAgent`
 * (include/dpgo_ros/PGOAgentROS.h:121) sit on libdpgo_b200.so.  Upstream these are
 * Eigen types; Eigen is not a dependency here, so DPGO::Matrix is a minimal dense
 * column-major class with exactly the operations the wrapper uses (SURVEY 8b,
 * "Matrix type"): (rows, cols) ctor, comma initialisation, operator()(i, j),
 * rows()/cols(), Zero/Identity, block(i, j, p, q), product, sum/difference,
 * transpose, norm.
 */
#ifndef DPGO_SHIM_TYPES_H
#define DPGO_SHIM_TYPES_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <functional>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <utility>
#include <vector>

// upstream's headers pull glog in: src/utils.cpp:284 uses CHECK with no include of its own
#if defined(__has_include)
#if __has_include(<glog/logging.h>)
#include <glog/logging.h>
#endif
#endif

namespace DPGO {

// the wrapper says `using namespace DPGO;` and then writes `vector<...>` unqualified (src/PGOAgentROS.cpp:288)
using std::vector;

// Storage of a Matrix: up to 32 coefficients (a lifted pose is r x (d+1) <= 8 x 4) live inside the object, anything
// larger on the heap.  The wrapper builds and copies one Matrix per public pose per message (src/utils.cpp:100-110,
// src/PGOAgentROS.cpp:1255-1284: millions per run); with heap storage those allocations were most of its host time.
class MatrixStorage {
 public:
  static constexpr size_t kInline = 32;
  MatrixStorage() = default;
  MatrixStorage(size_t n, double v) : n_(n) {
    if (n_ > kInline) heap_ = new double[n_];
    std::fill(data(), data() + n_, v);
  }
  MatrixStorage(const MatrixStorage &o) : n_(o.n_) {
    if (n_ > kInline) heap_ = new double[n_];
    std::copy(o.data(), o.data() + n_, data());
  }
  MatrixStorage(MatrixStorage &&o) noexcept : n_(o.n_), heap_(o.heap_) {
    if (!heap_) std::copy(o.buf_, o.buf_ + n_, buf_);
    o.heap_ = nullptr;
    o.n_ = 0;
  }
  MatrixStorage &operator=(const MatrixStorage &o) {
    if (this != &o) {
      MatrixStorage t(o);
      swap(t);
    }
    return *this;
  }
  MatrixStorage &operator=(MatrixStorage &&o) noexcept {
    if (this != &o) swap(o);
    return *this;
  }
  ~MatrixStorage() { delete[] heap_; }
  void swap(MatrixStorage &o) noexcept {
    std::swap_ranges(buf_, buf_ + kInline, o.buf_);
    std::swap(n_, o.n_);
    std::swap(heap_, o.heap_);
  }
  size_t size() const { return n_; }
  double *data() { return heap_ ? heap_ : buf_; }
  const double *data() const { return heap_ ? heap_ : buf_; }
  double &operator[](size_t i) { return data()[i]; }
  double operator[](size_t i) const { return data()[i]; }
  double *begin() { return data(); }
  double *end() { return data() + n_; }
  const double *begin() const { return data(); }
  const double *end() const { return data() + n_; }

 private:
  size_t n_ = 0;
  double *heap_ = nullptr;
  double buf_[kInline];
};

class Matrix {
 public:
  Matrix() = default;
  Matrix(size_t rows, size_t cols) : r_(rows), c_(cols), a_(rows * cols, 0.0) {}
  static Matrix Zero(size_t rows, size_t cols) { return Matrix(rows, cols); }
  static Matrix Identity(size_t rows, size_t cols) {
    Matrix M(rows, cols);
    for (size_t i = 0; i < rows && i < cols; ++i) M(i, i) = 1.0;
    return M;
  }
  size_t rows() const { return r_; }
  size_t cols() const { return c_; }
  size_t size() const { return a_.size(); }
  double *data() { return a_.data(); }              // column-major, like Eigen::MatrixXd
  const double *data() const { return a_.data(); }
  double &operator()(size_t i, size_t j) { return a_[j * r_ + i]; }
  double operator()(size_t i, size_t j) const { return a_[j * r_ + i]; }
  double &operator()(size_t i) { return a_[i]; }    // vectors
  double operator()(size_t i) const { return a_[i]; }

  // M << a, b, c, ...;   row-major fill order, as Eigen's comma initialiser (src/utils.cpp:68-71)
  class CommaInit {
   public:
    CommaInit(Matrix &M, double first) : M_(M), k_(0) { put(first); }
    CommaInit &operator,(double v) {
      put(v);
      return *this;
    }

   private:
    void put(double v) {
      if (k_ >= M_.size()) throw std::out_of_range("DPGO::Matrix: too many coefficients");
      M_(k_ / M_.cols(), k_ % M_.cols()) = v;
      ++k_;
    }
    Matrix &M_;
    size_t k_;
  };
  CommaInit operator<<(double first) { return CommaInit(*this, first); }

  // block view: readable, assignable from a Matrix (src/utils.cpp:163-164, src/PGOAgentROS.cpp:777)
  class Block {
   public:
    Block(Matrix &M, size_t i, size_t j, size_t p, size_t q) : M_(M), i_(i), j_(j), p_(p), q_(q) {}
    Block &operator=(const Matrix &B) {
      assert(B.rows() == p_ && B.cols() == q_);
      for (size_t c = 0; c < q_; ++c)
        for (size_t r = 0; r < p_; ++r) M_(i_ + r, j_ + c) = B(r, c);
      return *this;
    }
    Block &operator=(const Block &B) { return *this = static_cast<Matrix>(B); }   // element copy, like Eigen
    Block(const Block &) = default;
    operator Matrix() const {
      Matrix B(p_, q_);
      for (size_t c = 0; c < q_; ++c)
        for (size_t r = 0; r < p_; ++r) B(r, c) = M_(i_ + r, j_ + c);
      return B;
    }
    double operator()(size_t r, size_t c) const { return M_(i_ + r, j_ + c); }
    double operator()(size_t k) const { return q_ == 1 ? M_(i_ + k, j_) : M_(i_, j_ + k); }  // vector blocks
    Matrix transpose() const { return static_cast<Matrix>(*this).transpose(); }
    double norm() const { return static_cast<Matrix>(*this).norm(); }

   private:
    Matrix &M_;
    size_t i_, j_, p_, q_;
  };
  Block block(size_t i, size_t j, size_t p, size_t q) { return Block(*this, i, j, p, q); }
  Matrix block(size_t i, size_t j, size_t p, size_t q) const {
    Matrix B(p, q);
    for (size_t c = 0; c < q; ++c)
      for (size_t r = 0; r < p; ++r) B(r, c) = (*this)(i + r, j + c);
    return B;
  }
  Matrix col(size_t j) const { return block(0, j, r_, 1); }

  Matrix transpose() const {
    Matrix T(c_, r_);
    for (size_t j = 0; j < c_; ++j)
      for (size_t i = 0; i < r_; ++i) T(j, i) = (*this)(i, j);
    return T;
  }
  double norm() const {  // Frobenius (tests/testUtils.cpp:25)
    double s = 0;
    for (double v : a_) s += v * v;
    return std::sqrt(s);
  }
  double squaredNorm() const {
    double s = 0;
    for (double v : a_) s += v * v;
    return s;
  }
  void setZero() { std::fill(a_.begin(), a_.end(), 0.0); }

 private:
  size_t r_ = 0, c_ = 0;
  MatrixStorage a_;
};

inline Matrix operator*(const Matrix &A, const Matrix &B) {  // src/PGOAgentROS.cpp:1419
  assert(A.cols() == B.rows());
  Matrix C(A.rows(), B.cols());
  for (size_t j = 0; j < B.cols(); ++j)
    for (size_t k = 0; k < A.cols(); ++k) {
      const double b = B(k, j);
      for (size_t i = 0; i < A.rows(); ++i) C(i, j) += A(i, k) * b;
    }
  return C;
}
inline Matrix operator*(double s, const Matrix &A) {
  Matrix C = A;
  for (size_t k = 0; k < C.size(); ++k) C.data()[k] *= s;
  return C;
}
inline Matrix operator+(const Matrix &A, const Matrix &B) {
  assert(A.rows() == B.rows() && A.cols() == B.cols());
  Matrix C = A;
  for (size_t k = 0; k < C.size(); ++k) C.data()[k] += B.data()[k];
  return C;
}
inline Matrix operator-(const Matrix &A, const Matrix &B) {
  assert(A.rows() == B.rows() && A.cols() == B.cols());
  Matrix C = A;
  for (size_t k = 0; k < C.size(); ++k) C.data()[k] -= B.data()[k];
  return C;
}
inline std::ostream &operator<<(std::ostream &os, const Matrix &M) {
  for (size_t i = 0; i < M.rows(); ++i) {
    for (size_t j = 0; j < M.cols(); ++j) os << (j ? " " : "") << M(i, j);
    os << "\n";
  }
  return os;
}

// column vector (src/PGOAgentROS.cpp:1465: Vector::Zero(r))
class Vector : public Matrix {
 public:
  Vector() = default;
  explicit Vector(size_t n) : Matrix(n, 1) {}
  Vector(const Matrix &M) : Matrix(M) { assert(M.cols() == 1 || M.size() == 0); }
  static Vector Zero(size_t n) { return Vector(n); }
};

// ---- identifiers (src/PGOAgentROS.cpp:271, 684-685, 1434-1435; PGOAgentROS.h:189,192) ------------------
struct PoseID {
  unsigned int robot_id = 0, frame_id = 0;
  PoseID() = default;
  PoseID(unsigned int robot, unsigned int frame) : robot_id(robot), frame_id(frame) {}
  bool operator==(const PoseID &o) const { return robot_id == o.robot_id && frame_id == o.frame_id; }
};
struct ComparePoseID {
  bool operator()(const PoseID &a, const PoseID &b) const {
    return std::make_pair(a.robot_id, a.frame_id) < std::make_pair(b.robot_id, b.frame_id);
  }
};
struct EdgeID {
  PoseID src_pose_id, dst_pose_id;
  EdgeID() = default;
  EdgeID(const PoseID &src, const PoseID &dst) : src_pose_id(src), dst_pose_id(dst) {}
  bool operator==(const EdgeID &o) const { return src_pose_id == o.src_pose_id && dst_pose_id == o.dst_pose_id; }
  bool isOdometry() const {
    return src_pose_id.robot_id == dst_pose_id.robot_id && src_pose_id.frame_id + 1 == dst_pose_id.frame_id;
  }
  bool isPrivateLoopClosure() const {
    return src_pose_id.robot_id == dst_pose_id.robot_id && src_pose_id.frame_id + 1 != dst_pose_id.frame_id;
  }
  bool isSharedLoopClosure() const { return src_pose_id.robot_id != dst_pose_id.robot_id; }
};
struct HashEdgeID {
  size_t operator()(const EdgeID &e) const {
    size_t h = std::hash<unsigned>()(e.src_pose_id.robot_id);
    for (unsigned v : {e.src_pose_id.frame_id, e.dst_pose_id.robot_id, e.dst_pose_id.frame_id})
      h ^= std::hash<unsigned>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};

// ---- poses (src/PGOAgentROS.cpp:353-357, 1420-1421, 1463-1466; src/utils.cpp:154-164) --------------------
// r x (d+1) block [Y | p]; r == d for an unlifted pose
class LiftedPose {
 public:
  LiftedPose() : LiftedPose(3, 3) {}
  LiftedPose(unsigned int r, unsigned int d) : r_(r), d_(d), X_(Matrix::Identity(r, d + 1)) {}
  LiftedPose(const Matrix &X) : r_((unsigned)X.rows()), d_((unsigned)X.cols() - 1), X_(X) {}  // PoseDict::emplace(id, Matrix)
  unsigned int r() const { return r_; }
  unsigned int d() const { return d_; }
  const Matrix &getData() const { return X_; }
  void setData(const Matrix &X) {
    assert(X.rows() == r_ && X.cols() == d_ + 1);
    X_ = X;
  }
  Matrix pose() const { return X_; }
  Matrix::Block rotation() { return X_.block(0, 0, r_, d_); }
  Matrix rotation() const { return static_cast<const Matrix &>(X_).block(0, 0, r_, d_); }
  Matrix::Block translation() { return X_.block(0, d_, r_, 1); }
  Matrix translation() const { return static_cast<const Matrix &>(X_).block(0, d_, r_, 1); }

 protected:
  unsigned int r_, d_;
  Matrix X_;
};
class Pose : public LiftedPose {
 public:
  explicit Pose(unsigned int d = 3) : LiftedPose(d, d) {}
  Pose(const Matrix &T) : LiftedPose(T) {}
  Pose inverse() const {
    const Matrix R = rotation(), t = translation();
    Matrix T(d_, d_ + 1);
    T.block(0, 0, d_, d_) = R.transpose();
    T.block(0, d_, d_, 1) = -1.0 * (R.transpose() * t);
    return Pose(T);
  }
  Pose operator*(const Pose &o) const {
    Matrix T(d_, d_ + 1);
    T.block(0, 0, d_, d_) = rotation() * o.rotation();
    T.block(0, d_, d_, 1) = rotation() * o.translation() + translation();
    return Pose(T);
  }
};

// n poses side by side: r x (d+1) n  (src/utils.cpp:156-157: d x (d+1) n for a trajectory)
class LiftedPoseArray {
 public:
  LiftedPoseArray(unsigned int r, unsigned int d, unsigned int n) : r_(r), d_(d), n_(n), X_(r, (size_t)(d + 1) * n) {
    for (unsigned i = 0; i < n; ++i)
      for (unsigned k = 0; k < d && k < r; ++k) X_(k, (size_t)i * (d + 1) + k) = 1.0;
  }
  unsigned int r() const { return r_; }
  unsigned int d() const { return d_; }
  unsigned int n() const { return n_; }
  const Matrix &getData() const { return X_; }
  void setData(const Matrix &X) {
    assert(X.rows() == r_ && X.cols() == (size_t)(d_ + 1) * n_);
    X_ = X;
  }
  Matrix::Block pose(unsigned int i) { return X_.block(0, (size_t)i * (d_ + 1), r_, d_ + 1); }
  Matrix pose(unsigned int i) const { return static_cast<const Matrix &>(X_).block(0, (size_t)i * (d_ + 1), r_, d_ + 1); }
  Matrix::Block rotation(unsigned int i) { return X_.block(0, (size_t)i * (d_ + 1), r_, d_); }
  Matrix rotation(unsigned int i) const { return static_cast<const Matrix &>(X_).block(0, (size_t)i * (d_ + 1), r_, d_); }
  Matrix::Block translation(unsigned int i) { return X_.block(0, (size_t)i * (d_ + 1) + d_, r_, 1); }
  Matrix translation(unsigned int i) const {
    return static_cast<const Matrix &>(X_).block(0, (size_t)i * (d_ + 1) + d_, r_, 1);
  }

 protected:
  unsigned int r_, d_, n_;
  Matrix X_;
};
class PoseArray : public LiftedPoseArray {
 public:
  PoseArray(unsigned int d, unsigned int n) : LiftedPoseArray(d, d, n) {}
};

typedef std::map<PoseID, LiftedPose, ComparePoseID> PoseDict;

// ---- agent state / status (tests/testUtils.cpp:56-69, msg/Status.msg:1-11, src/utils.cpp:262-281) ----------
enum PGOAgentState { WAIT_FOR_DATA = 0, WAIT_FOR_INITIALIZATION = 1, INITIALIZED = 2 };

struct PGOAgentStatus {
  unsigned agentID = 0;
  PGOAgentState state = WAIT_FOR_DATA;
  unsigned instanceNumber = 0, iterationNumber = 0;
  bool readyToTerminate = false;
  double relativeChange = 0;
  PGOAgentStatus() = default;
  PGOAgentStatus(unsigned id, PGOAgentState s, unsigned instance, unsigned iteration, bool ready, double relChange)
      : agentID(id), state(s), instanceNumber(instance), iterationNumber(iteration), readyToTerminate(ready),
        relativeChange(relChange) {}
};

enum class InitializationMethod { Odometry, Chordal, GNC_TLS };

// local solver parameters (src/PGOAgentROSNode.cpp:82-100)
struct ROptParameters {
  enum class ROptMethod { RTR, RGD };
  ROptMethod method = ROptMethod::RTR;
  bool verbose = false;
  double gradnorm_tol = 1e-2;
  double RGD_stepsize = 1e-3;
  bool RGD_use_preconditioner = true;
  double RTR_initial_radius = 100;
  int RTR_iterations = 3, RTR_tCG_iterations = 50;   // int: read with ros::param::get (src/PGOAgentROSNode.cpp:98-99)
};

// mLocalOptResult (src/PGOAgentROS.cpp:169-172)
struct ROPTResult {
  bool success = false;
  double fInit = 0, gradNormInit = 0, fOpt = 0, gradNormOpt = 0, relativeChange = 0, elapsedMs = 0;
};

}  // namespace DPGO
#endif
