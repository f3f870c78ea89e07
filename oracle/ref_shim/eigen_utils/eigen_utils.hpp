// eigen_utils.hpp -- stand-in for MIT's eigen_utils pod (absent, no version pinned: SURVEY.md 8c), restricted to
// what the reference's state-estimator/src/mav_state_est/rbis.{hpp,cpp} uses.  TEST INFRASTRUCTURE ONLY.
// Every semantic here is the [RECALLED] restatement tabulated in SURVEY.md 8c -- the same assumptions the
// oracle makes; compiling the reference's own rbis.cpp against this header pins rbis.cpp:12-304 line by line,
// it does NOT pin eigen_utils itself.
#pragma once
#include <cmath>
#include <cstdint>
#include <iostream>

#include "../mini_eigen.hpp"

#ifndef eigen_dump
#define eigen_dump(MAT) do {} while (0)  // eigen_utils debug print, MSE/mav_state_est.cpp:19-20
#endif

namespace eigen_utils {
struct GVec;
}
namespace Eigen {
template <> struct traits<eigen_utils::GVec> {
  typedef double Scalar;
  enum { Rows = 3, Cols = 1 };
};
}  // namespace Eigen

namespace eigen_utils {

// run-time switchable constants (kept in sync with the oracle's Constants by oracle/ref_capi.cpp)
struct ShimConstants {
  double g_val = 9.8;
  double chi_tol = 1e-6;
  bool ctor_folds_chi = true;
};
inline ShimConstants& shim_constants() {
  static ShimConstants c;
  return c;
}

// g_vec = (0, 0, -g_val): an object that converts to a Vector3d so that `q.inverse() * g_vec` follows a changed g_val
struct GVec : public Eigen::MatrixBase<GVec> {
  int rows_() const { return 3; }
  int cols_() const { return 1; }
  double coeff_(int i, int) const { return i == 2 ? -shim_constants().g_val : 0.0; }
};
static const GVec g_vec = GVec();

template <class V>
Eigen::Matrix3d skewHat(const Eigen::MatrixBase<V>& v) {
  Eigen::Matrix3d m;
  m(0, 0) = 0; m(0, 1) = -v(2); m(0, 2) = v(1);
  m(1, 0) = v(2); m(1, 1) = 0; m(1, 2) = -v(0);
  m(2, 0) = -v(1); m(2, 1) = v(0); m(2, 2) = 0;
  return m;
}

// quaternion of AngleAxis(angle, axis)
inline Eigen::Quaterniond angleAxisToQuat(double angle, const Eigen::Vector3d& axis) {
  const double s = std::sin(0.5 * angle), c = std::cos(0.5 * angle);
  return Eigen::Quaterniond(c, s * axis(0), s * axis(1), s * axis(2));
}
// AngleAxis(q), Eigen >= 3.3 form
inline void quatToAngleAxis(const Eigen::Quaterniond& q, double& angle, Eigen::Vector3d& axis) {
  double n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
  if (n == 0.0) {
    angle = 0;
    axis = Eigen::Vector3d(1, 0, 0);
    return;
  }
  angle = 2.0 * std::atan2(n, std::fabs(q.w()));
  if (q.w() < 0) n = -n;
  axis = Eigen::Vector3d(q.x() / n, q.y() / n, q.z() / n);
}
inline double mod2pi(double a) {  // libbot bot_mod2pi: result in [-pi, pi)
  const double PI = 3.14159265358979323846;
  if (a >= -PI && a < PI) return a;
  a = std::fmod(a + PI, 2 * PI);
  if (a < 0) a += 2 * PI;
  return a - PI;
}
// subtractQuats(q1, q2) = axis * angle of q2^-1 * q1
inline Eigen::Vector3d subtractQuats(const Eigen::Quaterniond& q1, const Eigen::Quaterniond& q2) {
  const Eigen::Quaterniond r = q2.inverse() * q1;
  double angle;
  Eigen::Vector3d axis;
  quatToAngleAxis(r, angle, axis);
  return axis * mod2pi(angle);
}
inline void quaternionToBotDouble(double bot_quat[4], const Eigen::Quaterniond& q) {
  bot_quat[0] = q.w(); bot_quat[1] = q.x(); bot_quat[2] = q.y(); bot_quat[3] = q.z();
}
inline void botDoubleToQuaternion(Eigen::Quaterniond& q, const double bot_quat[4]) {
  q = Eigen::Quaterniond(bot_quat[0], bot_quat[1], bot_quat[2], bot_quat[3]);
}

class RigidBodyState {
 public:
  enum { angular_velocity_ind = 0, velocity_ind = 3, chi_ind = 6, position_ind = 9, acceleration_ind = 12, basic_num_states = 15 };
  typedef Eigen::Block<Eigen::VectorXd, 3, 1> Block3Element;
  typedef const Eigen::Block<const Eigen::VectorXd, 3, 1> ConstBlock3Element;

  Eigen::VectorXd vec;
  Eigen::Quaterniond quat;
  int64_t utime;

  explicit RigidBodyState(int state_dim = basic_num_states) : vec(Eigen::VectorXd::Zero(state_dim)), quat(Eigen::Quaterniond::Identity()), utime(0) {}
  // explicit: with real Eigen the argument of `RBIS(K * z_resid)` is an expression type, so only RBIS(const VectorNd&) is viable
  explicit RigidBodyState(const Eigen::VectorXd& arg_vec) : vec(arg_vec), quat(Eigen::Quaterniond::Identity()), utime(0) {
    if (shim_constants().ctor_folds_chi) this->chiToQuat();
  }
  RigidBodyState(const Eigen::VectorXd& arg_vec, const Eigen::Quaterniond& arg_quat) : vec(arg_vec), quat(arg_quat), utime(0) {}

  Block3Element angularVelocity() { return vec.block<3, 1>(angular_velocity_ind, 0); }
  Block3Element velocity() { return vec.block<3, 1>(velocity_ind, 0); }
  Block3Element chi() { return vec.block<3, 1>(chi_ind, 0); }
  Block3Element position() { return vec.block<3, 1>(position_ind, 0); }
  Block3Element acceleration() { return vec.block<3, 1>(acceleration_ind, 0); }
  ConstBlock3Element angularVelocity() const { return vec.block<3, 1>(angular_velocity_ind, 0); }
  ConstBlock3Element velocity() const { return vec.block<3, 1>(velocity_ind, 0); }
  ConstBlock3Element chi() const { return vec.block<3, 1>(chi_ind, 0); }
  ConstBlock3Element position() const { return vec.block<3, 1>(position_ind, 0); }
  ConstBlock3Element acceleration() const { return vec.block<3, 1>(acceleration_ind, 0); }
  const Eigen::Quaterniond& orientation() const { return quat; }
  Eigen::Quaterniond& orientation() { return quat; }

  static Eigen::Vector3i angularVelocityInds() { return Eigen::Vector3i::LinSpaced(angular_velocity_ind, angular_velocity_ind + 2); }
  static Eigen::Vector3i velocityInds() { return Eigen::Vector3i::LinSpaced(velocity_ind, velocity_ind + 2); }
  static Eigen::Vector3i chiInds() { return Eigen::Vector3i::LinSpaced(chi_ind, chi_ind + 2); }
  static Eigen::Vector3i positionInds() { return Eigen::Vector3i::LinSpaced(position_ind, position_ind + 2); }
  static Eigen::Vector3i accelerationInds() { return Eigen::Vector3i::LinSpaced(acceleration_ind, acceleration_ind + 2); }

  void chiToQuat() {
    const Eigen::Vector3d c = this->chi();
    const double n = c.norm();
    if (n > shim_constants().chi_tol) {
      this->quat = this->quat * angleAxisToQuat(n, c / n);
      this->chi() = Eigen::Vector3d::Zero();
    }
  }
  void quatToChi() {
    this->chi() = subtractQuats(this->quat, Eigen::Quaterniond::Identity());
    this->quat = Eigen::Quaterniond::Identity();
  }
  void addState(const RigidBodyState& rs_to_add) {
    this->vec += rs_to_add.vec;
    this->chiToQuat();
    this->quat = this->quat * rs_to_add.quat;
  }
  void subtractState(const RigidBodyState& rs_to_subtract) {
    this->vec -= rs_to_subtract.vec;
    this->quat = rs_to_subtract.quat.inverse() * this->quat;
  }
};

inline std::ostream& operator<<(std::ostream& os, const RigidBodyState& s) {
  return os << s.vec.transpose() << " | " << s.quat.w() << " " << s.quat.x() << " " << s.quat.y() << " " << s.quat.z();
}


// ---- what state-estimator/src/noise_id/noise_id.cpp uses ([RECALLED] eigen_utils/eigen_select_block.hpp, eigen_numerical.hpp) ----
template <class M, class IR, class IC>
Eigen::MatrixXd selectBlockByIndices(const Eigen::MatrixBase<M>& m, const Eigen::MatrixBase<IR>& rows, const Eigen::MatrixBase<IC>& cols) {
  Eigen::MatrixXd out(rows.rows(), cols.rows());
  for (int i = 0; i < rows.rows(); i++)
    for (int j = 0; j < cols.rows(); j++) out(i, j) = m.coeff(rows.coeff(i, 0), cols.coeff(j, 0));
  return out;
}
template <class M, class IR>
Eigen::VectorXd selectRowsByIndices(const Eigen::MatrixBase<M>& m, const Eigen::MatrixBase<IR>& rows) {
  Eigen::VectorXd out(rows.rows());
  for (int i = 0; i < rows.rows(); i++) out(i) = m.coeff(rows.coeff(i, 0), 0);
  return out;
}
// log of a Gaussian density without its constant: -log det(sigma) - (mu - x)^T sigma^-1 (mu - x)   (same form as MSE/rbis.cpp:142)
template <class X, class MU, class S>
double loglike_normalized(const Eigen::MatrixBase<X>& x, const Eigen::MatrixBase<MU>& mu, const Eigen::MatrixBase<S>& sigma) {
  Eigen::VectorXd diff = mu - x;
  Eigen::MatrixXd sg = sigma;
  Eigen::VectorXd sol = sg.ldlt().solve(diff);
  double q = 0;
  for (int i = 0; i < diff.rows(); i++) q += diff(i) * sol(i);
  return -std::log(sg.determinant()) - q;
}

}  // namespace eigen_utils
