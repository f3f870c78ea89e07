// rbis_oracle.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// An Eigen-free, single-filter, IEEE-double restatement of the RBIS EKF hot path of
// openhumanoids/pronto.  It exists so the CUDA path has something to be checked against and so
// bench.py has a CPU baseline to time.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may link or call anything in oracle/.
//
// PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path.  The oracle is pinned against the
// reference's OWN sources instead: oracle/Makefile (`make ref`) compiles the unmodified rbis.cpp,
// rbis_update_interface.cpp, update_history.cpp and mav_state_est.cpp from /root/reference into oracle/_ref/librbis_ref.so
// against stand-in headers (oracle/ref_shim/) for the absent Eigen 3 / eigen_utils / LCM / libbot, and
// tests/test_ref_pins_oracle.py + tests/test_smoother.py hold oracle == reference to 1e-12.  What stays RESTATED (and is
// therefore "parity unpinned" in the strict sense) is the inside of eigen_utils: RigidBodyState's algebra, g_val and the
// chiToQuat tolerance follow SURVEY.md section 8c in both the stand-in headers and this file; the two constants that
// cannot be checked in-repo (G_VAL, CHI_TOL) are runtime-settable.
//
// Reference files followed (all under /root/reference/state-estimator/src/mav_state_est/):
//   rbis.hpp:19-146                   types and API names
//   rbis.cpp:12-227                   the math (dense, as written)
//   rbis_update_interface.hpp:8-120   update objects
//   rbis_update_interface.cpp:23-107  updateFilter bodies, log-likelihood bookkeeping
//   mav_state_est.cpp:12-96           addUpdate roll-forward loop
//   rbis.cpp:234-266, mav_state_est.cpp:98-189   EKF smoother step and backwards pass
//   update_history.cpp:5-54           time-ordered multimap history
// Where the reference uses heap-allocated Eigen::MatrixXd this file uses std::vector<double>; where
// it uses fixed-size Eigen matrices this file uses fixed arrays.  Matrices are column-major like
// Eigen's default.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

namespace rbis_oracle {

// ---- constants living in the absent eigen_utils; settable for what-if tests ----
struct Constants {
  double g_val = 9.8;      // |g_vec|, g_vec = (0,0,-g_val)             [RECALLED]
  double chi_tol = 1e-6;   // chiToQuat() folds chi only if norm > tol   [RECALLED]
  bool ctor_folds_chi = true;  // RigidBodyState(VectorXd) ctor calls chiToQuat()   [RECALLED]
};
Constants& constants();

enum {
  angular_velocity_ind = 0, velocity_ind = 3, chi_ind = 6, position_ind = 9, acceleration_ind = 12,
  basic_num_states = 15, gyro_bias_ind = 15, accel_bias_ind = 18, rbis_num_states = 21  // rbis.hpp:22-24
};
constexpr int N = rbis_num_states;

struct Quat {  // Eigen::Quaterniond, stored (w,x,y,z) as in libbot messages
  double w = 1, x = 0, y = 0, z = 0;
};
Quat quatMul(const Quat& a, const Quat& b);            // Eigen operator*
Quat quatInverse(const Quat& q);                       // conj / squaredNorm (no renormalisation)
void quatRotate(const Quat& q, const double v[3], double out[3]);  // Eigen q * v (_transformVector)
void quatToRotationMatrix(const Quat& q, double R[9]);  // column-major 3x3, Eigen toRotationMatrix
Quat quatFromAngleAxis(double angle, const double axis[3]);
void quatToAngleAxis(const Quat& q, double& angle, double axis[3]);  // Eigen >= 3.3 atan2 form
void subtractQuats(const Quat& q1, const Quat& q2, double out[3]);   // Log(q2^-1 * q1), eigen_utils
void skewHat(const double v[3], double M[9]);                        // column-major 3x3

// RBIM: Eigen::Matrix<double,21,21>, column-major (rbis.hpp:30,123)
struct RBIM {
  double m[N * N];
  RBIM() { setZero(); }
  void setZero();
  double& operator()(int r, int c) { return m[r + N * c]; }
  double operator()(int r, int c) const { return m[r + N * c]; }
};

// RBIS: rbis.hpp:19-121 on top of eigen_utils::RigidBodyState (restated)
struct RBIS {
  double vec[N];
  Quat quat;
  int64_t utime = 0;
  RBIS();                                   // zeros + identity (rbis.hpp:40-44)
  explicit RBIS(const double v[N]);         // rbis.hpp:46-50; ctor may fold chi (constants())
  RBIS(const double v[N], const Quat& q);   // rbis.hpp:52-56
  double* angularVelocity() { return vec + angular_velocity_ind; }
  double* velocity() { return vec + velocity_ind; }
  double* chi() { return vec + chi_ind; }
  double* position() { return vec + position_ind; }
  double* acceleration() { return vec + acceleration_ind; }
  double* gyroBias() { return vec + gyro_bias_ind; }
  double* accelBias() { return vec + accel_bias_ind; }
  const double* angularVelocity() const { return vec + angular_velocity_ind; }
  const double* velocity() const { return vec + velocity_ind; }
  const double* chi() const { return vec + chi_ind; }
  const double* acceleration() const { return vec + acceleration_ind; }
  void chiToQuat();
  void quatToChi();
  void addState(const RBIS& d);
  void subtractState(const RBIS& o);
};

// ---- free functions, names and argument order as rbis.hpp:125-146 ----
void getIMUProcessLinearizationContinuous(const RBIS& state, RBIM& Ac);
void insUpdateState(const double gyro[3], const double accelerometer[3], double dt, RBIS& state);
void insUpdateCovariance(double q_gyro, double q_accel, double q_gyro_bias, double q_accel_bias,
                         const RBIS& state, RBIM& cov, double dt);
// R: m x m column-major; C: m x 21 column-major; K: 21 x m column-major (resized)
double matrixMeasurementGetKandCovDelta(int m, const std::vector<double>& R, const std::vector<double>& C,
                                        const RBIM& cov, const std::vector<double>& z_resid, RBIM& dcov,
                                        std::vector<double>& K);
double indexedMeasurement(int m, const double* z, const double* R, const int32_t* z_indices, const RBIS& state,
                          const RBIM& cov, RBIS& dstate, RBIM& dcov);
double indexedPlusOrientationMeasurement(int m, const double* z, const Quat& quat, const double* R,
                                         const int32_t* z_indices, const RBIS& state, const RBIM& cov,
                                         RBIS& dstate, RBIM& dcov);
void rbisApplyDelta(const RBIS& prior_state, const RBIM& prior_cov, const RBIS& dstate, const RBIM& dcov,
                    RBIS& posterior_state, RBIM& posterior_cov);
// EKF / RTS smoothing step, rbis.cpp:234-266 ("next" row 1 of SURVEY.md 8f)
void ekfSmoothingStep(const RBIS& next_state_pred, const RBIM& next_cov_pred, const RBIS& next_state, const RBIM& next_cov,
                      double dt, RBIS& cur_state, RBIM& cur_cov);

// ---- IMU-noise identification, state-estimator/src/noise_id/noise_id.cpp:9-65 ("next" row 2 of SURVEY.md 8f) ----
// sampleProcessForward: windows of N_window IMU steps rolled forward from the truth history with TWO covariance
// propagations per step (the given noise, and zero noise from the same start) -- state_errors / covs per complete window.
void sampleProcessForward(const std::vector<RBIS>& truth_state_history, const std::vector<RBIM>& truth_cov_history, double dt,
                          double q_gyro, double q_accel, int N_window, std::vector<RBIS>& state_errors, std::vector<RBIM>& covs);
// eigen_utils::loglike_normalized [RECALLED]: -log det(sigma) - (mu - x)^T sigma^-1 (mu - x)  (no 1/2, no 2 pi; the same
// form MSE/rbis.cpp:142 writes inline)
double loglike_normalized(int n, const double* x, const double* mu, const double* sigma /* n x n column-major */);
// negLogLikelihood over the active index set (noise_id.cpp:44-65)
double negLogLikelihood(const std::vector<RBIS>& state_errors, const std::vector<RBIM>& covs, int n_active, const int32_t* active_inds);

// ---- IMU conditioning ("next" row 4 of SURVEY.md 8f): second-order IIR notch, estimate_tools/src/estimate_tools/
// iir_notch.cpp:3-60, run as a cascade on the accelerometer channels by InsHandler::doFilter
// (state-estimator/src/mav_state_est/sensor_handlers.cpp:29-41,155-162) ----
struct IIRNotch {
  double b[3], a[3];  // numerator / denominator (iir_notch.cpp:17-33)
  double x[2], y[2];  // carried inputs / outputs
  IIRNotch(double notch_freq, double fs);
  static void secondOrderNotch(double Wo, double BW, double num[3], double den[3]);
  double processSample(double input);  // iir_notch.cpp:35-60
};

// ---- update objects: rbis_update_interface.hpp:8-120 ----
class RBISUpdateInterface {
 public:
  enum sensor_enum {
    ins, gps, vicon, laser, laser_gpf, scan_matcher, optical_flow, reset, invalid, rgbd, fovis, legodo,
    pose_meas, altimeter, airspeed, sideslip, init_message, viewer, yawlock
  };
  int64_t utime;
  RBIS posterior_state;
  RBIM posterior_covariance;
  double loglikelihood = 0;
  sensor_enum sensor_id;
  RBISUpdateInterface(sensor_enum id, int64_t t) : utime(t), sensor_id(id) {}
  virtual ~RBISUpdateInterface() {}
  virtual void updateFilter(const RBIS& prior_state, const RBIM& prior_cov, double prior_loglikelihood) = 0;
};

class RBISResetUpdate : public RBISUpdateInterface {
 public:
  RBIS reset_state;
  RBIM reset_cov;
  RBISResetUpdate(const RBIS& s, const RBIM& c, sensor_enum id, int64_t t)
      : RBISUpdateInterface(id, t), reset_state(s), reset_cov(c) {}
  void updateFilter(const RBIS&, const RBIM&, double) override;
};

class RBISIMUProcessStep : public RBISUpdateInterface {
 public:
  double gyro[3], accelerometer[3];
  double dt, q_gyro, q_accel, q_gyro_bias, q_accel_bias;
  RBISIMUProcessStep(const double g[3], const double a[3], double q_gyro_, double q_accel_, double q_gyro_bias_,
                     double q_accel_bias_, double dt_, int64_t t);
  void updateFilter(const RBIS&, const RBIM&, double) override;
};

class RBISIndexedMeasurement : public RBISUpdateInterface {
 public:
  std::vector<int32_t> index;
  std::vector<double> measurement;
  std::vector<double> measurement_cov;  // m x m column-major
  RBISIndexedMeasurement(int m, const int32_t* idx, const double* z, const double* R, sensor_enum id, int64_t t);
  void updateFilter(const RBIS&, const RBIM&, double) override;
};

class RBISIndexedPlusOrientationMeasurement : public RBISUpdateInterface {
 public:
  std::vector<int32_t> index;
  std::vector<double> measurement;
  std::vector<double> measurement_cov;
  Quat orientation;
  RBISIndexedPlusOrientationMeasurement(int m, const int32_t* idx, const double* z, const double* R, const Quat& q,
                                        sensor_enum id, int64_t t);
  void updateFilter(const RBIS&, const RBIM&, double) override;
};

// ---- history + driver: update_history.cpp:5-54, mav_state_est.cpp:12-96 ----
class updateHistory {
 public:
  typedef std::multimap<int64_t, RBISUpdateInterface*> historyMap;
  typedef historyMap::iterator historyMapIterator;
  historyMap updateMap;
  explicit updateHistory(RBISUpdateInterface* init);
  ~updateHistory();
  historyMapIterator addToHistory(RBISUpdateInterface* rbisu);  // end() if discarded (too old)
  void clearHistoryBeforeUtime(int64_t utime);
};

class MavStateEstimator {
 public:
  MavStateEstimator(RBISResetUpdate* init_state, int64_t utime_history_span);
  updateHistory::historyMapIterator unprocessed_updates_start;
  updateHistory history;
  int64_t utime_history_span;
  int64_t n_update_calls = 0;  // oracle-only instrumentation: number of updateFilter calls made
  void addUpdate(RBISUpdateInterface* update, bool roll_forward);
  void getHeadState(RBIS& head_state, RBIM& head_cov);
  double getMeasurementsLogLikelihood();
  void EKFSmoothBackwardsPass(double dt);  // mav_state_est.cpp:98-189
};

}  // namespace rbis_oracle
