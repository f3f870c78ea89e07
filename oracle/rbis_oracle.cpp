// rbis_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).  PARITY UNPINNED: see the
// header of rbis_oracle.hpp.  Every function cites the reference lines it restates; paths are under
// /root/reference/state-estimator/src/mav_state_est/ unless noted.  The arithmetic is the
// reference's AS WRITTEN (dense Ad*P*Ad^T, dense Wc*Qc*Wc^T*dt, dense (K*C)*cov), so that timing
// this file gives a fair picture of the reference's CPU cost.
#include "rbis_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace rbis_oracle {

Constants& constants() {
  static Constants c;
  return c;
}

// ------------------------------------------------------------------------------------------------
// Eigen::Quaterniond / AngleAxisd semantics
// ------------------------------------------------------------------------------------------------
Quat quatMul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}

Quat quatInverse(const Quat& q) {
  double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
  Quat r;
  if (n2 > 0) {
    r.w = q.w / n2; r.x = -q.x / n2; r.y = -q.y / n2; r.z = -q.z / n2;
  } else {
    r.w = r.x = r.y = r.z = 0;
  }
  return r;
}

void quatRotate(const Quat& q, const double v[3], double out[3]) {
  // uv = 2 * (q.vec x v);  out = v + w*uv + q.vec x uv
  double uvx = q.y * v[2] - q.z * v[1];
  double uvy = q.z * v[0] - q.x * v[2];
  double uvz = q.x * v[1] - q.y * v[0];
  uvx += uvx; uvy += uvy; uvz += uvz;
  out[0] = v[0] + q.w * uvx + (q.y * uvz - q.z * uvy);
  out[1] = v[1] + q.w * uvy + (q.z * uvx - q.x * uvz);
  out[2] = v[2] + q.w * uvz + (q.x * uvy - q.y * uvx);
}

void quatToRotationMatrix(const Quat& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  // column-major: R[r + 3c]
  R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
  R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
  R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

Quat quatFromAngleAxis(double angle, const double axis[3]) {
  Quat q;
  const double h = 0.5 * angle;
  const double s = std::sin(h);
  q.w = std::cos(h);
  q.x = s * axis[0]; q.y = s * axis[1]; q.z = s * axis[2];
  return q;
}

void quatToAngleAxis(const Quat& q, double& angle, double axis[3]) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  if (n != 0.0) {
    angle = 2.0 * std::atan2(n, std::fabs(q.w));
    if (q.w < 0) n = -n;
    axis[0] = q.x / n; axis[1] = q.y / n; axis[2] = q.z / n;
  } else {
    angle = 0; axis[0] = 1; axis[1] = 0; axis[2] = 0;
  }
}

static double bot_mod2pi_positive(double vin) {
  const double q = vin / (2 * M_PI) + 0.5;
  const int qi = (int)q;
  return vin - qi * (2 * M_PI);
}
static double bot_mod2pi(double vin) {  // libbot math_util.h [RECALLED]: map to [-pi, pi]
  return vin < 0 ? -bot_mod2pi_positive(-vin) : bot_mod2pi_positive(vin);
}

void subtractQuats(const Quat& q1, const Quat& q2, double out[3]) {
  const Quat r = quatMul(quatInverse(q2), q1);
  double angle, axis[3];
  quatToAngleAxis(r, angle, axis);
  angle = bot_mod2pi(angle);
  out[0] = axis[0] * angle; out[1] = axis[1] * angle; out[2] = axis[2] * angle;
}

void skewHat(const double v[3], double M[9]) {
  M[0] = 0;     M[3] = -v[2]; M[6] = v[1];
  M[1] = v[2];  M[4] = 0;     M[7] = -v[0];
  M[2] = -v[1]; M[5] = v[0];  M[8] = 0;
}

// ------------------------------------------------------------------------------------------------
// RBIS / RigidBodyState algebra
// ------------------------------------------------------------------------------------------------
void RBIM::setZero() { std::memset(m, 0, sizeof(m)); }

RBIS::RBIS() { std::memset(vec, 0, sizeof(vec)); }
RBIS::RBIS(const double v[N]) {
  std::memcpy(vec, v, sizeof(vec));
  if (constants().ctor_folds_chi) chiToQuat();
}
RBIS::RBIS(const double v[N], const Quat& q) : quat(q) { std::memcpy(vec, v, sizeof(vec)); }

void RBIS::chiToQuat() {
  double* c = chi();
  const double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  if (n > constants().chi_tol) {
    const double axis[3] = {c[0] / n, c[1] / n, c[2] / n};
    quat = quatMul(quat, quatFromAngleAxis(n, axis));
    c[0] = c[1] = c[2] = 0;
  }
}
void RBIS::quatToChi() {
  subtractQuats(quat, Quat(), chi());
  quat = Quat();
}
void RBIS::addState(const RBIS& d) {
  for (int i = 0; i < N; i++) vec[i] += d.vec[i];
  chiToQuat();
  quat = quatMul(quat, d.quat);
}
void RBIS::subtractState(const RBIS& o) {
  for (int i = 0; i < N; i++) vec[i] -= o.vec[i];
  quat = quatMul(quatInverse(o.quat), quat);
}

// small dense helpers (column-major)
static inline void setBlock3(RBIM& A, int r0, int c0, const double B[9], double scale) {
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) A(r0 + r, c0 + c) = scale * B[r + 3 * c];
}

// ------------------------------------------------------------------------------------------------
// rbis.cpp:12-35
// ------------------------------------------------------------------------------------------------
void getIMUProcessLinearizationContinuous(const RBIS& state, RBIM& Ac) {
  Ac.setZero();
  double omega_hat[9], vb_hat[9], Rm[9], g_hat[9];
  skewHat(state.angularVelocity(), omega_hat);
  skewHat(state.velocity(), vb_hat);
  quatToRotationMatrix(state.quat, Rm);
  const double g_vec[3] = {0, 0, -constants().g_val};
  double gb[3];
  quatRotate(quatInverse(state.quat), g_vec, gb);
  skewHat(gb, g_hat);

  setBlock3(Ac, velocity_ind, velocity_ind, omega_hat, -1.0);  // :20
  setBlock3(Ac, velocity_ind, chi_ind, g_hat, 1.0);            // :21
  setBlock3(Ac, chi_ind, chi_ind, omega_hat, -1.0);            // :24
  setBlock3(Ac, position_ind, velocity_ind, Rm, 1.0);          // :27
  double RV[9];                                                // :28  -R * vb_hat
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Rm[r + 3 * k] * vb_hat[k + 3 * c];
      RV[r + 3 * c] = -s;
    }
  setBlock3(Ac, position_ind, chi_ind, RV, 1.0);
  setBlock3(Ac, velocity_ind, gyro_bias_ind, vb_hat, -1.0);    // :31
  for (int i = 0; i < 3; i++) {
    Ac(velocity_ind + i, accel_bias_ind + i) = -1.0;           // :32
    Ac(chi_ind + i, gyro_bias_ind + i) = -1.0;                 // :33
  }
}

// ------------------------------------------------------------------------------------------------
// rbis.cpp:37-75
// ------------------------------------------------------------------------------------------------
void insUpdateState(const double gyro[3], const double accelerometer[3], double dt, RBIS& state) {
  for (int i = 0; i < 3; i++) {
    state.angularVelocity()[i] = gyro[i] - state.gyroBias()[i];          // :50
    state.acceleration()[i] = accelerometer[i] - state.accelBias()[i];   // :51
  }
  RBIS dstate;  // zeros
  const double* w = state.angularVelocity();
  const double* v = state.velocity();
  // :55  dv = -w x v
  dstate.velocity()[0] = -(w[1] * v[2] - w[2] * v[1]);
  dstate.velocity()[1] = -(w[2] * v[0] - w[0] * v[2]);
  dstate.velocity()[2] = -(w[0] * v[1] - w[1] * v[0]);
  // :56  dv += q^-1 * g_vec + a
  const double g_vec[3] = {0, 0, -constants().g_val};
  double gb[3];
  quatRotate(quatInverse(state.quat), g_vec, gb);
  for (int i = 0; i < 3; i++) dstate.velocity()[i] += gb[i] + state.acceleration()[i];
  // :58
  for (int i = 0; i < 3; i++) dstate.chi()[i] = w[i];
  // :59
  quatRotate(state.quat, v, dstate.position());
  // :62-63
  for (int i = 0; i < N; i++) dstate.vec[i] *= dt;
  dstate.chiToQuat();
  // :69
  state.addState(dstate);
}

// ------------------------------------------------------------------------------------------------
// rbis.cpp:77-122 (dense, as written)
// ------------------------------------------------------------------------------------------------
void insUpdateCovariance(double q_gyro, double q_accel, double q_gyro_bias, double q_accel_bias, const RBIS& state,
                         RBIM& cov, double dt) {
  RBIM Ac;
  getIMUProcessLinearizationContinuous(state, Ac);

  const int gyro_ind = 0, accelerometer_ind = 3, gyro_bias_noise_ind = 6, accelerometer_bias_noise_ind = 9;
  const int num_inputs = 12;
  static thread_local double Wc[N * num_inputs];  // 21 x 12 column-major
  std::memset(Wc, 0, sizeof(double) * N * num_inputs);
  double vhat[9];
  skewHat(state.velocity(), vhat);
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) Wc[(velocity_ind + r) + N * (gyro_ind + c)] = vhat[r + 3 * c];  // :93
  for (int i = 0; i < 3; i++) {
    Wc[(velocity_ind + i) + N * (accelerometer_ind + i)] = 1;                  // :94
    Wc[(chi_ind + i) + N * (gyro_ind + i)] = 1;                                // :97
    Wc[(gyro_bias_ind + i) + N * (gyro_bias_noise_ind + i)] = 1;               // :99
    Wc[(accel_bias_ind + i) + N * (accelerometer_bias_noise_ind + i)] = 1;     // :100
  }
  double Qc_vec[num_inputs];
  for (int i = 0; i < 3; i++) {
    Qc_vec[gyro_ind + i] = q_gyro;
    Qc_vec[accelerometer_ind + i] = q_accel;
    Qc_vec[gyro_bias_noise_ind + i] = q_gyro_bias;
    Qc_vec[accelerometer_bias_noise_ind + i] = q_accel_bias;
  }

  // :112-114  Ad = I + Ac*dt
  RBIM Ad;
  for (int i = 0; i < N * N; i++) Ad.m[i] = Ac.m[i] * dt;
  for (int i = 0; i < N; i++) Ad(i, i) += 1.0;

  // :109,116  Qc dense 12x12 (asDiagonal assigned to a Matrix12d);  Qd = Wc * Qc * Wc^T * dt
  // (dense: (Wc*Qc) 21x12x12, then * Wc^T 21x12x21, then * dt)
  double Qc[num_inputs * num_inputs];
  std::memset(Qc, 0, sizeof(Qc));
  for (int i = 0; i < num_inputs; i++) Qc[i + num_inputs * i] = Qc_vec[i];
  static thread_local double WQ[N * num_inputs];
  for (int c = 0; c < num_inputs; c++)
    for (int r = 0; r < N; r++) {
      double s = 0;
      for (int k = 0; k < num_inputs; k++) s += Wc[r + N * k] * Qc[k + num_inputs * c];
      WQ[r + N * c] = s;
    }
  RBIM Qd;
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double s = 0;
      for (int k = 0; k < num_inputs; k++) s += WQ[r + N * k] * Wc[c + N * k];
      Qd(r, c) = s * dt;
    }

  // :118  cov = Ad * cov * Ad^T + Qd   (two dense 21^3 products)
  RBIM T;
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += Ad(r, k) * cov(k, c);
      T(r, c) = s;
    }
  RBIM out;
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += T(r, k) * Ad(c, k);
      out(r, c) = s + Qd(r, c);
    }
  cov = out;

  // :120-121
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) {
      cov(acceleration_ind + r, acceleration_ind + c) = (r == c) ? q_accel : 0.0;
      cov(angular_velocity_ind + r, angular_velocity_ind + c) = (r == c) ? q_gyro : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// dynamic-size helpers standing in for Eigen::LDLT<MatrixXd> and MatrixXd::determinant()
// ------------------------------------------------------------------------------------------------
namespace {

// LDL^T with symmetric diagonal pivoting (largest |diagonal| first), the strategy Eigen::LDLT uses.
// P A P^T = L D L^T.  A is m x m column-major, only read.
struct LDLT {
  int m;
  std::vector<double> L;   // unit lower, column-major
  std::vector<double> D;
  std::vector<int> perm;   // perm[i] = original index at pivoted position i
  explicit LDLT(int m_, const std::vector<double>& A) : m(m_), L(A), D(m_), perm(m_) {
    for (int i = 0; i < m; i++) perm[i] = i;
    std::vector<double>& W = L;  // work in place on a full symmetric copy
    for (int k = 0; k < m; k++) {
      int p = k;
      double best = std::fabs(W[k + m * k]);
      for (int i = k + 1; i < m; i++)
        if (std::fabs(W[i + m * i]) > best) { best = std::fabs(W[i + m * i]); p = i; }
      if (p != k) {  // symmetric row/column swap
        for (int j = 0; j < m; j++) std::swap(W[k + m * j], W[p + m * j]);
        for (int j = 0; j < m; j++) std::swap(W[j + m * k], W[j + m * p]);
        std::swap(perm[k], perm[p]);
      }
      const double d = W[k + m * k];
      D[k] = d;
      if (d == 0.0) continue;
      for (int i = k + 1; i < m; i++) W[i + m * k] /= d;
      for (int j = k + 1; j < m; j++)
        for (int i = k + 1; i < m; i++) W[i + m * j] -= W[i + m * k] * d * W[j + m * k];
    }
    for (int j = 0; j < m; j++)
      for (int i = 0; i < m; i++)
        if (i < j) W[i + m * j] = 0; else if (i == j) W[i + m * j] = 1;
  }
  // solve A X = B for nrhs right-hand sides, B is m x nrhs column-major, overwritten with X
  void solveInPlace(std::vector<double>& B, int nrhs) const {
    std::vector<double> y(m);
    for (int c = 0; c < nrhs; c++) {
      double* b = &B[(size_t)m * c];
      for (int i = 0; i < m; i++) y[i] = b[perm[i]];
      for (int i = 0; i < m; i++)
        for (int k = 0; k < i; k++) y[i] -= L[i + m * k] * y[k];
      for (int i = 0; i < m; i++) y[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
      for (int i = m - 1; i >= 0; i--)
        for (int k = i + 1; k < m; k++) y[i] -= L[k + m * i] * y[k];
      for (int i = 0; i < m; i++) b[perm[i]] = y[i];
    }
  }
};

// determinant through LU with partial pivoting (what MatrixXd::determinant() does for dynamic sizes)
double determinantLU(int m, std::vector<double> A) {
  double det = 1.0;
  for (int k = 0; k < m; k++) {
    int p = k;
    double best = std::fabs(A[k + m * k]);
    for (int i = k + 1; i < m; i++)
      if (std::fabs(A[i + m * k]) > best) { best = std::fabs(A[i + m * k]); p = i; }
    if (best == 0.0) return 0.0;
    if (p != k) {
      for (int j = 0; j < m; j++) std::swap(A[k + m * j], A[p + m * j]);
      det = -det;
    }
    det *= A[k + m * k];
    for (int i = k + 1; i < m; i++) {
      const double f = A[i + m * k] / A[k + m * k];
      for (int j = k + 1; j < m; j++) A[i + m * j] -= f * A[k + m * j];
    }
  }
  return det;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// rbis.cpp:124-143 (dense, as written: C*cov evaluated twice, (K*C)*cov left to right)
// ------------------------------------------------------------------------------------------------
double matrixMeasurementGetKandCovDelta(int m, const std::vector<double>& R, const std::vector<double>& C,
                                        const RBIM& cov, const std::vector<double>& z_resid, RBIM& cov_delta,
                                        std::vector<double>& K) {
  // C * cov  (m x 21)
  std::vector<double> CP((size_t)m * N);
  for (int c = 0; c < N; c++)
    for (int r = 0; r < m; r++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += C[r + m * k] * cov(k, c);
      CP[r + m * c] = s;
    }
  // :134-135  S = R + (C*cov)*C^T
  std::vector<double> S(R);
  for (int c = 0; c < m; c++)
    for (int r = 0; r < m; r++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += CP[r + m * k] * C[c + m * k];
      S[r + m * c] += s;
    }
  // :137
  LDLT Sldlt(m, S);
  // :139  K^T = S^-1 (C*cov)
  std::vector<double> Kt(CP);
  Sldlt.solveInPlace(Kt, N);
  K.assign((size_t)N * m, 0.0);
  for (int r = 0; r < N; r++)
    for (int c = 0; c < m; c++) K[r + N * c] = Kt[c + m * r];
  // :140  cov_delta = (K*C)*cov
  static thread_local double KC[N * N];
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double s = 0;
      for (int k = 0; k < m; k++) s += K[r + N * k] * C[k + m * c];
      KC[r + N * c] = s;
    }
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double s = 0;
      for (int k = 0; k < N; k++) s += KC[r + N * k] * cov(k, c);
      cov_delta(r, c) = s;
    }
  // :142
  std::vector<double> x(z_resid);
  Sldlt.solveInPlace(x, 1);
  double quad = 0;
  for (int i = 0; i < m; i++) quad += z_resid[i] * x[i];
  return -std::log(determinantLU(m, S)) - quad;
}

// rbis.cpp:160-178
double indexedMeasurement(int m, const double* z, const double* R, const int32_t* z_indices, const RBIS& state,
                          const RBIM& cov, RBIS& dstate, RBIM& dcov) {
  std::vector<double> z_resid(m), K, C((size_t)m * N, 0.0), Rv(R, R + (size_t)m * m);
  for (int ii = 0; ii < m; ii++) {
    z_resid[ii] = z[ii] - state.vec[z_indices[ii]];
    C[ii + m * z_indices[ii]] = 1;
  }
  const double loglikelihood = matrixMeasurementGetKandCovDelta(m, Rv, C, cov, z_resid, dcov, K);
  double dx[N];
  for (int r = 0; r < N; r++) {
    double s = 0;
    for (int k = 0; k < m; k++) s += K[r + N * k] * z_resid[k];
    dx[r] = s;
  }
  dstate = RBIS(dx);  // :175
  return loglikelihood;
}

// rbis.cpp:189-217
double indexedPlusOrientationMeasurement(int m, const double* z, const Quat& quat, const double* R,
                                         const int32_t* z_indices, const RBIS& state, const RBIM& cov,
                                         RBIS& dstate, RBIM& dcov) {
  std::vector<double> z_resid(m), K, C((size_t)m * N, 0.0), Rv(R, R + (size_t)m * m);
  double dquat[3];
  subtractQuats(quat, state.quat, dquat);  // :199
  for (int ii = 0; ii < m; ii++) {
    if (z_indices[ii] >= chi_ind && z_indices[ii] <= chi_ind + 2)
      z_resid[ii] = dquat[z_indices[ii] - chi_ind];  // :204, z(ii) ignored
    else
      z_resid[ii] = z[ii] - state.vec[z_indices[ii]];
    C[ii + m * z_indices[ii]] = 1;
  }
  const double loglikelihood = matrixMeasurementGetKandCovDelta(m, Rv, C, cov, z_resid, dcov, K);
  double dx[N];
  for (int r = 0; r < N; r++) {
    double s = 0;
    for (int k = 0; k < m; k++) s += K[r + N * k] * z_resid[k];
    dx[r] = s;
  }
  dstate = RBIS(dx);  // :214
  return loglikelihood;
}

// rbis.cpp:219-227
void rbisApplyDelta(const RBIS& prior_state, const RBIM& prior_cov, const RBIS& dstate, const RBIM& dcov,
                    RBIS& posterior_state, RBIM& posterior_cov) {
  posterior_state = prior_state;
  posterior_cov = prior_cov;
  posterior_state.addState(dstate);
  for (int i = 0; i < N * N; i++) posterior_cov.m[i] -= dcov.m[i];
}

// ------------------------------------------------------------------------------------------------
// rbis_update_interface.cpp:23-107
// ------------------------------------------------------------------------------------------------
void RBISResetUpdate::updateFilter(const RBIS&, const RBIM&, double) {
  posterior_state = reset_state;
  posterior_covariance = reset_cov;
  loglikelihood = 0;
}

RBISIMUProcessStep::RBISIMUProcessStep(const double g[3], const double a[3], double q_gyro_, double q_accel_,
                                       double q_gyro_bias_, double q_accel_bias_, double dt_, int64_t t)
    : RBISUpdateInterface(ins, t), dt(dt_), q_gyro(q_gyro_), q_accel(q_accel_), q_gyro_bias(q_gyro_bias_),
      q_accel_bias(q_accel_bias_) {
  for (int i = 0; i < 3; i++) { gyro[i] = g[i]; accelerometer[i] = a[i]; }
}

void RBISIMUProcessStep::updateFilter(const RBIS& prior_state, const RBIM& prior_cov, double prior_loglikelihood) {
  posterior_state = prior_state;
  posterior_covariance = prior_cov;
  insUpdateState(gyro, accelerometer, dt, posterior_state);
  // NOTE linearised at the PRIOR state (rbis_update_interface.cpp:39)
  insUpdateCovariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, prior_state, posterior_covariance, dt);
  loglikelihood = prior_loglikelihood;
}

RBISIndexedMeasurement::RBISIndexedMeasurement(int m, const int32_t* idx, const double* z, const double* R,
                                               sensor_enum id, int64_t t)
    : RBISUpdateInterface(id, t), index(idx, idx + m), measurement(z, z + m), measurement_cov(R, R + (size_t)m * m) {}

void RBISIndexedMeasurement::updateFilter(const RBIS& prior_state, const RBIM& prior_cov,
                                          double prior_loglikelihood) {
  RBIS dstate;
  RBIM dcov;
  const double cur = indexedMeasurement((int)index.size(), measurement.data(), measurement_cov.data(), index.data(),
                                        prior_state, prior_cov, dstate, dcov);
  rbisApplyDelta(prior_state, prior_cov, dstate, dcov, posterior_state, posterior_covariance);
  loglikelihood = prior_loglikelihood + cur;
}

RBISIndexedPlusOrientationMeasurement::RBISIndexedPlusOrientationMeasurement(int m, const int32_t* idx,
                                                                             const double* z, const double* R,
                                                                             const Quat& q, sensor_enum id, int64_t t)
    : RBISUpdateInterface(id, t), index(idx, idx + m), measurement(z, z + m),
      measurement_cov(R, R + (size_t)m * m), orientation(q) {}

void RBISIndexedPlusOrientationMeasurement::updateFilter(const RBIS& prior_state, const RBIM& prior_cov,
                                                         double prior_loglikelihood) {
  RBIS dstate;
  RBIM dcov;
  const double cur = indexedPlusOrientationMeasurement((int)index.size(), measurement.data(), orientation,
                                                       measurement_cov.data(), index.data(), prior_state, prior_cov,
                                                       dstate, dcov);
  rbisApplyDelta(prior_state, prior_cov, dstate, dcov, posterior_state, posterior_covariance);
  loglikelihood = prior_loglikelihood + cur;
}

// ------------------------------------------------------------------------------------------------
// update_history.cpp:5-54
// ------------------------------------------------------------------------------------------------
updateHistory::updateHistory(RBISUpdateInterface* init) { updateMap.insert({init->utime, init}); }
updateHistory::~updateHistory() {
  for (auto& kv : updateMap) delete kv.second;
  updateMap.clear();
}

updateHistory::historyMapIterator updateHistory::addToHistory(RBISUpdateInterface* rbisu) {
  // hinted insert at end(): equal keys keep arrival order (upper-bound position)  :26
  historyMapIterator it = updateMap.insert(updateMap.end(), {rbisu->utime, rbisu});
  if (it == updateMap.begin()) {  // older than everything in history: discard  :28-39
    delete rbisu;
    updateMap.erase(it);
    return updateMap.end();
  }
  return it;
}

void updateHistory::clearHistoryBeforeUtime(int64_t utime) {
  // stl_utils::stlmultimap_get_lower [RECALLED]: last entry with key <= utime
  historyMapIterator it_before = updateMap.upper_bound(utime);
  if (it_before == updateMap.begin()) return;  // nothing at or before utime
  --it_before;
  if (it_before == updateMap.begin()) return;
  for (historyMapIterator it = updateMap.begin(); it != it_before; ++it) delete it->second;
  updateMap.erase(updateMap.begin(), it_before);
}

// ------------------------------------------------------------------------------------------------
// mav_state_est.cpp:12-96
// ------------------------------------------------------------------------------------------------
MavStateEstimator::MavStateEstimator(RBISResetUpdate* init_state, int64_t span)
    : history(init_state), utime_history_span(span) {
  init_state->updateFilter(RBIS(), RBIM(), 0);  // :16
  init_state->posterior_state.utime = init_state->utime;
  unprocessed_updates_start = history.updateMap.end();
}

void MavStateEstimator::addUpdate(RBISUpdateInterface* update, bool roll_forward) {
  updateHistory::historyMapIterator added_it = history.addToHistory(update);
  // The reference dereferences added_it even when the update was discarded (end()); here a
  // discarded update is simply skipped.
  if (added_it != history.updateMap.end()) {
    if (unprocessed_updates_start == history.updateMap.end() ||
        added_it->first < unprocessed_updates_start->first) {  // :35-38
      unprocessed_updates_start = added_it;
    }
  }
  if (!roll_forward) return;
  if (unprocessed_updates_start == history.updateMap.end()) return;

  updateHistory::historyMapIterator prev_it = unprocessed_updates_start;
  prev_it--;
  updateHistory::historyMapIterator current_it = unprocessed_updates_start;
  while (current_it != history.updateMap.end()) {  // :50-70
    RBISUpdateInterface* current_update = current_it->second;
    RBISUpdateInterface* prev_update = prev_it->second;
    current_update->updateFilter(prev_update->posterior_state, prev_update->posterior_covariance,
                                 prev_update->loglikelihood);
    current_update->posterior_state.utime = current_update->utime;
    n_update_calls++;
    prev_it = current_it;
    current_it++;
  }
  const int64_t newest_utime = prev_it->first;  // :74-77
  history.clearHistoryBeforeUtime(newest_utime - utime_history_span);
  unprocessed_updates_start = history.updateMap.end();
}

void MavStateEstimator::getHeadState(RBIS& head_state, RBIM& head_cov) {
  RBISUpdateInterface* head_update = history.updateMap.rbegin()->second;
  head_state = head_update->posterior_state;
  head_cov = head_update->posterior_covariance;
}

double MavStateEstimator::getMeasurementsLogLikelihood() {
  return history.updateMap.rbegin()->second->loglikelihood;
}

// ------------------------------------------------------------------------------------------------
// IIR notch: estimate_tools/src/estimate_tools/iir_notch.cpp:3-60
// ------------------------------------------------------------------------------------------------
IIRNotch::IIRNotch(double notch_freq, double fs) {
  const double Wo = notch_freq / (fs / 2);  // :6
  const double Bw = Wo;
  secondOrderNotch(Wo, Bw, b, a);
  x[0] = x[1] = y[0] = y[1] = 0;
}
void IIRNotch::secondOrderNotch(double Wo, double BW, double num[3], double den[3]) {
  const double Ab = std::fabs(10 * std::log10(.5));  // :18
  BW = BW * M_PI;
  Wo = Wo * M_PI;
  const double Gb = std::pow(10, -Ab / 20.);
  const double beta = (std::sqrt(1.0 - Gb * Gb) / Gb) * std::tan(BW / 2.0);
  const double gain = 1 / (1 + beta);
  num[0] = gain * 1.0; num[1] = gain * (-2.0 * std::cos(Wo)); num[2] = gain * 1;  // :30
  den[0] = 1.0; den[1] = -2 * gain * std::cos(Wo); den[2] = 2 * gain - 1;          // :31
}
double IIRNotch::processSample(double input) {
  // x_temp = (input, x0, x1), y_temp = (0, y0, y1); output = x_temp.b - y_temp.a   (:37-50)
  const double xb = (input * b[0] + x[0] * b[1]) + x[1] * b[2];
  const double ya = (0 * a[0] + y[0] * a[1]) + y[1] * a[2];
  const double output = xb - ya;
  x[1] = x[0]; x[0] = input;
  y[1] = y[0]; y[0] = output;
  return output;
}

// ------------------------------------------------------------------------------------------------
// EKF smoother: rbis.cpp:234-266 and mav_state_est.cpp:98-189
// ------------------------------------------------------------------------------------------------
void ekfSmoothingStep(const RBIS& next_state_pred, const RBIM& next_cov_pred, const RBIS& next_state, const RBIM& next_cov,
                      double dt, RBIS& cur_state, RBIM& cur_cov) {
  RBIM Ac;
  getIMUProcessLinearizationContinuous(cur_state, Ac);  // :237
  RBIM Ad;
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) Ad(r, c) = (r == c ? 1.0 : 0.0) + Ac(r, c) * dt;  // :238-239
  // :243-250  zero-uncertainty biases: the 3x3 diagonal block is replaced by the identity
  RBIM S = next_cov_pred;
  for (int b0 : {(int)gyro_bias_ind, (int)accel_bias_ind}) {
    bool any = false;
    for (int k = 0; k < 3; k++) any = any || (next_cov_pred(b0 + k, b0 + k) < .00000000001);
    if (any)
      for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) S(b0 + r, b0 + c) = (r == c) ? 1.0 : 0.0;
  }
  // :253-254  L^T = S^-1 (Ad cur_cov)
  std::vector<double> Sv(S.m, S.m + N * N), Lt(N * N);
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double acc = 0;
      for (int k = 0; k < N; k++) acc += Ad(r, k) * cur_cov(k, c);
      Lt[r + N * c] = acc;
    }
  LDLT ldlt(N, Sv);
  ldlt.solveInPlace(Lt, N);
  // :256  cur_cov += L (next_cov - next_cov_pred) L^T, L(i,k) = Lt(k,i)
  RBIM D, T;
  for (int i = 0; i < N * N; i++) D.m[i] = next_cov.m[i] - next_cov_pred.m[i];
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double acc = 0;
      for (int k = 0; k < N; k++) acc += Lt[k + N * r] * D(k, c);
      T(r, c) = acc;  // L * D
    }
  for (int c = 0; c < N; c++)
    for (int r = 0; r < N; r++) {
      double acc = 0;
      for (int k = 0; k < N; k++) acc += T(r, k) * Lt[k + N * c];  // (L D) * L^T, L^T(k,c) = Lt(k,c)
      cur_cov(r, c) += acc;
    }
  // :258-265
  RBIS smooth_resid = next_state;
  smooth_resid.subtractState(next_state_pred);
  smooth_resid.quatToChi();
  double innov[N];
  for (int r = 0; r < N; r++) {
    double acc = 0;
    for (int k = 0; k < N; k++) acc += Lt[k + N * r] * smooth_resid.vec[k];
    innov[r] = acc;
  }
  RBIS smooth_innov(innov);
  cur_state.addState(smooth_innov);
}

// mav_state_est.cpp:98-189, statement by statement (including its iterator arithmetic: after the trailing measurements
// have been rewound the iterator is decremented once more, :130, and `--begin()` serves as the stop mark, :139-140,
// which on libstdc++ is the header node, i.e. end()).
void MavStateEstimator::EKFSmoothBackwardsPass(double dt) {
  updateHistory::historyMapIterator current_it = unprocessed_updates_start;
  current_it--;
  bool measurement_cur_step = false;
  RBIS next_state;
  RBIM next_cov;
  RBISUpdateInterface* current_update = current_it->second;
  while (current_update->sensor_id != RBISUpdateInterface::ins) {  // :116-124
    current_update = current_it->second;
    if (!measurement_cur_step) {
      next_state = current_update->posterior_state;
      next_cov = current_update->posterior_covariance;
      measurement_cur_step = true;
    }
    current_it--;
  }
  RBIS next_state_pred = current_update->posterior_state;
  RBIM next_cov_pred = current_update->posterior_covariance;
  if (!measurement_cur_step) {
    next_state = current_update->posterior_state;
    next_cov = current_update->posterior_covariance;
  }
  current_it--;  // :130
  RBIS cur_state, cur_state_pred;
  RBIM cur_cov, cur_cov_pred;
  measurement_cur_step = false;
  updateHistory::historyMapIterator before_begin = history.updateMap.end();  // `begin()--` of the reference
  while (before_begin != current_it) {
    current_update = current_it->second;
    if (current_update->sensor_id == RBISUpdateInterface::ins) {
      cur_state_pred = current_update->posterior_state;
      cur_cov_pred = current_update->posterior_covariance;
      if (!measurement_cur_step) {
        cur_state = cur_state_pred;
        cur_cov = cur_cov_pred;
      }
      ekfSmoothingStep(next_state_pred, next_cov_pred, next_state, next_cov, dt, cur_state, cur_cov);
      current_update->posterior_covariance = cur_cov;
      current_update->posterior_state = cur_state;
      updateHistory::historyMapIterator forward_it = current_it;
      forward_it++;
      while (forward_it->second->sensor_id != RBISUpdateInterface::ins) {  // :165-169
        forward_it->second->posterior_state = current_update->posterior_state;
        forward_it->second->posterior_covariance = current_update->posterior_covariance;
        forward_it++;
      }
      next_state = cur_state;
      next_cov = cur_cov;
      next_state_pred = cur_state_pred;
      next_cov_pred = cur_cov_pred;
      measurement_cur_step = false;
    } else {
      if (!measurement_cur_step) {
        cur_state = current_update->posterior_state;
        cur_cov = current_update->posterior_covariance;
        measurement_cur_step = true;
      }
    }
    if (current_it == history.updateMap.begin()) current_it = history.updateMap.end();  // `--begin()`
    else current_it--;
  }
}

// ------------------------------------------------------------------------------------------------
// noise identification, state-estimator/src/noise_id/noise_id.cpp:9-65
// ------------------------------------------------------------------------------------------------
void sampleProcessForward(const std::vector<RBIS>& truth_state_history, const std::vector<RBIM>& truth_cov_history, double dt,
                          double q_gyro, double q_accel, int N_window, std::vector<RBIS>& state_errors, std::vector<RBIM>& covs) {
  const double q_gyro_bias = 0, q_accel_bias = 0;                       // noise_id.cpp:13-14
  size_t it = 0;                                                        // truth_state_it / truth_cov_it
  if (truth_state_history.empty()) return;
  while (true) {
    RBIS rolled_state = truth_state_history[it];                        // :19
    RBIM start_window_cov = truth_cov_history[it];                      // :21
    RBIM rolled_covariance = start_window_cov;                          // :22
    for (int ii = 0; ii < N_window; ii++) {
      insUpdateCovariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, rolled_state, rolled_covariance, dt);  // :24
      insUpdateCovariance(0, 0, 0, 0, rolled_state, start_window_cov, dt);                                   // :25
      insUpdateState(truth_state_history[it].angularVelocity(), truth_state_history[it].acceleration(), dt, rolled_state);  // :26
      rolled_state.utime = truth_state_history[it].utime;               // :27
      it++;                                                             // :31-32
      if (it == truth_state_history.size()) return;                     // :33-34
    }
    rolled_state.subtractState(truth_state_history[it]);                // :37
    rolled_state.quatToChi();                                           // :38
    state_errors.push_back(rolled_state);                               // :39
    RBIM d;
    for (int k = 0; k < N * N; k++) d.m[k] = rolled_covariance.m[k] - start_window_cov.m[k];
    covs.push_back(d);                                                  // :40
  }
}

double loglike_normalized(int n, const double* x, const double* mu, const double* sigma) {
  // LDL^T without pivoting of the symmetric positive definite sigma; det = prod d, solve by substitution
  std::vector<double> L((size_t)n * n, 0.0), D((size_t)n), diff((size_t)n), y((size_t)n);
  for (int i = 0; i < n; i++) diff[(size_t)i] = mu[i] - x[i];
  double logdet = 0;
  for (int k = 0; k < n; k++) {
    double d = sigma[k + n * k];
    for (int p = 0; p < k; p++) d -= L[(size_t)(k + n * p)] * L[(size_t)(k + n * p)] * D[(size_t)p];
    D[(size_t)k] = d;
    logdet += std::log(d);
    for (int i = k + 1; i < n; i++) {
      double v = sigma[i + n * k];
      for (int p = 0; p < k; p++) v -= L[(size_t)(i + n * p)] * L[(size_t)(k + n * p)] * D[(size_t)p];
      L[(size_t)(i + n * k)] = v / d;
    }
  }
  double quad = 0;
  for (int i = 0; i < n; i++) {
    double v = diff[(size_t)i];
    for (int k = 0; k < i; k++) v -= L[(size_t)(i + n * k)] * y[(size_t)k];
    y[(size_t)i] = v;
    quad += v * v / D[(size_t)i];
  }
  return -logdet - quad;
}

double negLogLikelihood(const std::vector<RBIS>& state_errors, const std::vector<RBIM>& covs, int n_active, const int32_t* active_inds) {
  double neg_likelihood = 0;                                            // noise_id.cpp:49
  std::vector<double> cov_active((size_t)n_active * n_active), error_active((size_t)n_active), zero((size_t)n_active, 0.0);
  for (size_t w = 0; w < state_errors.size(); w++) {
    for (int a = 0; a < n_active; a++) {
      error_active[(size_t)a] = state_errors[w].vec[active_inds[a]];
      for (int b = 0; b < n_active; b++) cov_active[(size_t)(a + n_active * b)] = covs[w](active_inds[a], active_inds[b]);
    }
    neg_likelihood -= loglike_normalized(n_active, error_active.data(), zero.data(), cov_active.data());  // :58-60
  }
  return neg_likelihood;
}

}  // namespace rbis_oracle
