/* tf/tf.h of the ROS stand-in: the quaternion <-> rotation-matrix conversions src/utils.cpp:60-106 uses.
 * Standard formulas (unit quaternion -> matrix; matrix -> quaternion by the largest-pivot rule).  TEST INFRASTRUCTURE. */
#ifndef ROS_STUB_TF_H
#define ROS_STUB_TF_H
#include <cmath>

#include "geometry_msgs/Point.h"
namespace tf {
class Vector3 {
 public:
  Vector3() = default;
  Vector3(double x, double y, double z) : v_{x, y, z} {}
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  double operator[](int i) const { return v_[i]; }
  double &operator[](int i) { return v_[i]; }

 private:
  double v_[3] = {0, 0, 0};
};
class Quaternion {
 public:
  Quaternion() = default;
  Quaternion(double x, double y, double z, double w) : q_{x, y, z, w} {}
  double x() const { return q_[0]; }
  double y() const { return q_[1]; }
  double z() const { return q_[2]; }
  double w() const { return q_[3]; }

 private:
  double q_[4] = {0, 0, 0, 1};
};
class Matrix3x3 {
 public:
  Matrix3x3() = default;
  explicit Matrix3x3(const Quaternion &q) {
    const double d = q.x() * q.x() + q.y() * q.y() + q.z() * q.z() + q.w() * q.w();
    const double s = 2.0 / d;
    const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    m_[0] = Vector3(1.0 - (yy + zz), xy - wz, xz + wy);
    m_[1] = Vector3(xy + wz, 1.0 - (xx + zz), yz - wx);
    m_[2] = Vector3(xz - wy, yz + wx, 1.0 - (xx + yy));
  }
  Matrix3x3(double xx, double xy, double xz, double yx, double yy, double yz, double zx, double zy, double zz) {
    m_[0] = Vector3(xx, xy, xz);
    m_[1] = Vector3(yx, yy, yz);
    m_[2] = Vector3(zx, zy, zz);
  }
  const Vector3 &operator[](int i) const { return m_[i]; }
  Vector3 &operator[](int i) { return m_[i]; }
  void getRotation(Quaternion &q) const {
    const double trace = m_[0][0] + m_[1][1] + m_[2][2];
    double t[4];
    if (trace > 0.0) {
      double s = std::sqrt(trace + 1.0);
      t[3] = s * 0.5;
      s = 0.5 / s;
      t[0] = (m_[2][1] - m_[1][2]) * s;
      t[1] = (m_[0][2] - m_[2][0]) * s;
      t[2] = (m_[1][0] - m_[0][1]) * s;
    } else {
      const int i = m_[0][0] < m_[1][1] ? (m_[1][1] < m_[2][2] ? 2 : 1) : (m_[0][0] < m_[2][2] ? 2 : 0);
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      double s = std::sqrt(m_[i][i] - m_[j][j] - m_[k][k] + 1.0);
      t[i] = s * 0.5;
      s = 0.5 / s;
      t[3] = (m_[k][j] - m_[j][k]) * s;
      t[j] = (m_[j][i] + m_[i][j]) * s;
      t[k] = (m_[k][i] + m_[i][k]) * s;
    }
    q = Quaternion(t[0], t[1], t[2], t[3]);
  }

 private:
  Vector3 m_[3];
};
inline void quaternionMsgToTF(const geometry_msgs::Quaternion &m, Quaternion &q) { q = Quaternion(m.x, m.y, m.z, m.w); }
inline void quaternionTFToMsg(const Quaternion &q, geometry_msgs::Quaternion &m) {
  m.x = q.x();
  m.y = q.y();
  m.z = q.z();
  m.w = q.w();
}
inline void pointMsgToTF(const geometry_msgs::Point &m, Vector3 &v) { v = Vector3(m.x, m.y, m.z); }
inline void pointTFToMsg(const Vector3 &v, geometry_msgs::Point &m) {
  m.x = v.x();
  m.y = v.y();
  m.z = v.z();
}
}  // namespace tf
#endif
