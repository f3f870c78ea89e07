// ref_capi.cpp -- extern "C" face of oracle/_ref/librbis_ref.so: the REFERENCE's own
// state-estimator/src/mav_state_est/rbis.cpp, compiled UNMODIFIED from /root/reference against the stand-in
// headers of oracle/ref_shim/ (Eigen, eigen_utils, LCM and libbot are not installed).  TEST INFRASTRUCTURE:
// it pins the oracle's restatement of rbis.cpp:12-227 line by line (tests/test_ref_pins_oracle.py).  What it
// does NOT pin is eigen_utils itself (RigidBodyState algebra, g_vec, chiToQuat tolerance): those semantics
// are the same [RECALLED] ones in both, see oracle/ref_shim/eigen_utils/eigen_utils.hpp.
// Same symbol names and signatures as oracle_capi.cpp's single-update functions.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include <cmath>
#include <estimate_tools/iir_notch.hpp>  // -I /root/reference/estimate_tools/src
#include <estimate_tools/imu_stream.hpp>
#include <list>
#include <noise_id/noise_id.hpp>          // -I /root/reference/state-estimator/src
#include <mav_est_legodo/rbis_legodo_common.hpp>  // -I /root/reference/motion_estimate/src
#include "mav_state_est.hpp"  // the reference's headers, found with -I /root/reference/state-estimator/src/mav_state_est

extern "C" int64_t rbis_ref_shim_history_span = 10000000;  // what the BotParam stand-in returns (ref_shim/bot_param)

using namespace MavStateEst;

namespace {
RBIS makeState(const double* vec, const double* quat) {
  RBIS::VectorNd v = Eigen::Map<const RBIS::VectorNd>(vec);
  return RBIS(v, Eigen::Quaterniond(quat[0], quat[1], quat[2], quat[3]));
}
void putState(const RBIS& s, double* vec, double* quat) {
  for (int i = 0; i < 21; i++) vec[i] = s.vec(i);
  quat[0] = s.quat.w(); quat[1] = s.quat.x(); quat[2] = s.quat.y(); quat[3] = s.quat.z();
}
RBIM getCov(const double* cov) { return RBIM(Eigen::Map<const RBIM>(cov)); }
void putCov(const RBIM& P, double* cov) { std::memcpy(cov, P.data(), sizeof(double) * 441); }
}  // namespace

extern "C" {

void orc_set_constants(double g_val, double chi_tol, int ctor_folds_chi) {
  eigen_utils::shim_constants().g_val = g_val;
  eigen_utils::shim_constants().chi_tol = chi_tol;
  eigen_utils::shim_constants().ctor_folds_chi = ctor_folds_chi != 0;
}

void orc_linearization(const double* vec, const double* quat, double* Ac) {
  RBIM A;
  getIMUProcessLinearizationContinuous(makeState(vec, quat), A);
  putCov(A, Ac);
}

void orc_ins_update_state(const double* gyro, const double* accel, double dt, double* vec, double* quat) {
  RBIS s = makeState(vec, quat);
  insUpdateState(Eigen::Vector3d(gyro[0], gyro[1], gyro[2]), Eigen::Vector3d(accel[0], accel[1], accel[2]), dt, s);
  putState(s, vec, quat);
}

void orc_ins_update_covariance(double q_gyro, double q_accel, double q_gyro_bias, double q_accel_bias, const double* vec,
                               const double* quat, double* cov, double dt) {
  RBIM P = getCov(cov);
  insUpdateCovariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, makeState(vec, quat), P, dt);
  putCov(P, cov);
}

double orc_indexed_measurement(int m, const double* z, const double* meas_quat, const double* R, const int32_t* idx,
                               const double* vec, const double* quat, const double* cov, double* dvec, double* dquat,
                               double* dcov) {
  Eigen::VectorXd zv = Eigen::Map<const Eigen::VectorXd>(z, m);
  Eigen::MatrixXd Rm = Eigen::Map<const Eigen::MatrixXd>(R, m, m);
  Eigen::VectorXi iv(m);
  for (int i = 0; i < m; i++) iv(i) = idx[i];
  RBIS ds;
  RBIM dP;
  double ll;
  if (meas_quat)
    ll = indexedPlusOrientationMeasurement(zv, Eigen::Quaterniond(meas_quat[0], meas_quat[1], meas_quat[2], meas_quat[3]), Rm, iv,
                                           makeState(vec, quat), getCov(cov), ds, dP);
  else
    ll = indexedMeasurement(zv, Rm, iv, makeState(vec, quat), getCov(cov), ds, dP);
  putState(ds, dvec, dquat);
  putCov(dP, dcov);
  return ll;
}

void orc_apply_delta(const double* vec, const double* quat, const double* cov, const double* dvec, const double* dquat,
                     const double* dcov, double* pvec, double* pquat, double* pcov) {
  RBIS post;
  RBIM Pp;
  rbisApplyDelta(makeState(vec, quat), getCov(cov), makeState(dvec, dquat), getCov(dcov), post, Pp);
  putState(post, pvec, pquat);
  putCov(Pp, pcov);
}

// ---- the reference's own update objects and history driver (rbis_update_interface.cpp:23-107,
// update_history.cpp, mav_state_est.cpp:12-96), one MavStateEstimator per filter; same contract as the oracle's
// orc_run_ensemble.  Too-old updates must not be passed: the reference dereferences the end() iterator that
// addToHistory returns for them (mav_state_est.cpp:33-38).  Returns -1 (the reference does not count calls). ----
typedef struct {
  int32_t m, has_orient, r_mode, sensor_id;
  int32_t idx[9];
  int32_t _pad;
  const double* z;
  const double* quat;
  const double* R;
} orc_stream_t;
typedef struct {
  int32_t kind, stream;
  int64_t row, utime;
  double dt;
} orc_event_t;

static int64_t run_ensemble_impl(int64_t Nf, int n_threads, double* vec, double* quat, double* cov, double* loglik, int64_t utime0,
                         const double* q_gyro, const double* q_accel, const double* q_gyro_bias, const double* q_accel_bias,
                         const double* imu, int n_streams, const orc_stream_t* streams, int64_t n_events,
                         const orc_event_t* events, int64_t history_span, double* trace_vec, double* trace_quat,
                         double* trace_cov, double* trace_loglik, double smooth_dt,
                         double* post_vec, double* post_quat, double* post_cov) {
  (void)n_streams;
  rbis_ref_shim_history_span = history_span;
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const int64_t n = next.fetch_add(1);
      if (n >= Nf) break;
      double v[21], q[4], P[441];
      for (int i = 0; i < 21; i++) v[i] = vec[i * Nf + n];
      for (int i = 0; i < 4; i++) q[i] = quat[i * Nf + n];
      for (int i = 0; i < 441; i++) P[i] = cov[i * Nf + n];
      RBIS s0 = makeState(v, q);
      s0.utime = utime0;
      MavStateEstimator est(new RBISResetUpdate(s0, getCov(P), RBISUpdateInterface::reset, utime0), (BotParam*)0);
      const double ll0 = loglik ? loglik[n] : 0.0;
      for (int64_t e = 0; e < n_events; e++) {
        const orc_event_t& ev = events[e];
        RBISUpdateInterface* u;
        if (ev.kind == 0) {
          Eigen::Vector3d g, a;
          for (int i = 0; i < 3; i++) {
            g(i) = imu[(ev.row * 6 + i) * Nf + n];
            a(i) = imu[(ev.row * 6 + 3 + i) * Nf + n];
          }
          u = new RBISIMUProcessStep(g, a, q_gyro[n], q_accel[n], q_gyro_bias[n], q_accel_bias[n], ev.dt, ev.utime);
        } else {
          const orc_stream_t& st = streams[ev.stream];
          const int m = st.m;
          Eigen::VectorXd z(m);
          Eigen::MatrixXd R = Eigen::MatrixXd::Zero(m, m);
          Eigen::VectorXi iv(m);
          for (int i = 0; i < m; i++) {
            z(i) = st.z[(ev.row * m + i) * Nf + n];
            iv(i) = st.idx[i];
          }
          if (st.r_mode == 0) R = Eigen::Map<const Eigen::MatrixXd>(st.R, m, m);
          else for (int i = 0; i < m; i++) R(i, i) = st.R[i * Nf + n];
          if (st.has_orient) {
            const Eigen::Quaterniond mq(st.quat[(ev.row * 4 + 0) * Nf + n], st.quat[(ev.row * 4 + 1) * Nf + n],
                                        st.quat[(ev.row * 4 + 2) * Nf + n], st.quat[(ev.row * 4 + 3) * Nf + n]);
            u = new RBISIndexedPlusOrientationMeasurement(iv, z, R, mq, (RBISUpdateInterface::sensor_enum)st.sensor_id, ev.utime);
          } else {
            u = new RBISIndexedMeasurement(iv, z, R, (RBISUpdateInterface::sensor_enum)st.sensor_id, ev.utime);
          }
        }
        est.addUpdate(u, true);
        if (trace_vec || trace_quat || trace_cov || trace_loglik) {
          RBIS hs;
          RBIM hP;
          est.getHeadState(hs, hP);
          double ov[21], oq[4];
          putState(hs, ov, oq);
          if (trace_vec) for (int i = 0; i < 21; i++) trace_vec[(e * 21 + i) * Nf + n] = ov[i];
          if (trace_quat) for (int i = 0; i < 4; i++) trace_quat[(e * 4 + i) * Nf + n] = oq[i];
          if (trace_cov) for (int i = 0; i < 441; i++) trace_cov[((int64_t)e * 441 + i) * Nf + n] = hP.data()[i];
          if (trace_loglik) trace_loglik[e * Nf + n] = ll0 + est.getMeasurementsLogLikelihood();
        }
      }
      if (smooth_dt > 0) {  // mav_state_est.cpp:98-189, then every update's posterior in history order (reset excluded)
        est.EKFSmoothBackwardsPass(smooth_dt);
        int64_t e = -1;
        for (auto& kv : est.history.updateMap) {
          if (e >= 0) {
            double pv[21], pq[4];
            putState(kv.second->posterior_state, pv, pq);
            for (int i = 0; i < 21; i++) post_vec[(e * 21 + i) * Nf + n] = pv[i];
            for (int i = 0; i < 4; i++) post_quat[(e * 4 + i) * Nf + n] = pq[i];
            for (int i = 0; i < 441; i++) post_cov[((int64_t)e * 441 + i) * Nf + n] = kv.second->posterior_covariance.data()[i];
          }
          e++;
        }
      }
      RBIS hs;
      RBIM hP;
      est.getHeadState(hs, hP);
      double ov[21], oq[4];
      putState(hs, ov, oq);
      for (int i = 0; i < 21; i++) vec[i * Nf + n] = ov[i];
      for (int i = 0; i < 4; i++) quat[i * Nf + n] = oq[i];
      for (int i = 0; i < 441; i++) cov[i * Nf + n] = hP.data()[i];
      if (loglik) loglik[n] = ll0 + est.getMeasurementsLogLikelihood();
    }
  };
  if (n_threads <= 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return -1;
}

int64_t orc_run_ensemble(int64_t Nf, int n_threads, double* vec, double* quat, double* cov, double* loglik, int64_t utime0,
                         const double* q_gyro, const double* q_accel, const double* q_gyro_bias, const double* q_accel_bias,
                         const double* imu, int n_streams, const orc_stream_t* streams, int64_t n_events,
                         const orc_event_t* events, int64_t history_span, double* trace_vec, double* trace_quat,
                         double* trace_cov, double* trace_loglik) {
  return run_ensemble_impl(Nf, n_threads, vec, quat, cov, loglik, utime0, q_gyro, q_accel, q_gyro_bias, q_accel_bias, imu,
                           n_streams, streams, n_events, events, history_span, trace_vec, trace_quat, trace_cov, trace_loglik,
                           0.0, nullptr, nullptr, nullptr);
}

// Forward pass as orc_run_ensemble, then MavStateEstimator::EKFSmoothBackwardsPass(smooth_dt) (mav_state_est.cpp:98-189);
// post_* [E][k][N] receive every update's posterior AFTER smoothing, in history order.  history_span must keep everything.
int64_t orc_smooth_ensemble(int64_t Nf, int n_threads, double* vec, double* quat, double* cov, int64_t utime0,
                            const double* q_gyro, const double* q_accel, const double* q_gyro_bias, const double* q_accel_bias,
                            const double* imu, int n_streams, const orc_stream_t* streams, int64_t n_events,
                            const orc_event_t* events, int64_t history_span, double smooth_dt, double* post_vec,
                            double* post_quat, double* post_cov) {
  return run_ensemble_impl(Nf, n_threads, vec, quat, cov, nullptr, utime0, q_gyro, q_accel, q_gyro_bias, q_accel_bias, imu,
                           n_streams, streams, n_events, events, history_span, nullptr, nullptr, nullptr, nullptr, smooth_dt,
                           post_vec, post_quat, post_cov);
}

// rbis.cpp:234-266 on one filter; cur_* are updated in place
void orc_ekf_smoothing_step(const double* np_vec, const double* np_quat, const double* np_cov, const double* n_vec,
                            const double* n_quat, const double* n_cov, double dt, double* c_vec, double* c_quat, double* c_cov) {
  RBIS cur = makeState(c_vec, c_quat);
  RBIM P = getCov(c_cov);
  ekfSmoothingStep(makeState(np_vec, np_quat), getCov(np_cov), makeState(n_vec, n_quat), getCov(n_cov), dt, cur, P);
  putState(cur, c_vec, c_quat);
  putCov(P, c_cov);
}

// the reference's OWN IIRNotch (estimate_tools/src/estimate_tools/iir_notch.cpp, compiled unmodified)
// InsHandler::doFilter (sensor_handlers.cpp:155-162) on ONE channel: n_stages notch filters at notch_freq * 2^i, fs,
// in cascade over n samples; state (x0,x1,y0,y1 per stage) in/out so that calls can be chained.  coeffs (optional): b,a per stage.
void orc_notch_cascade(double notch_freq, double fs, int n_stages, int64_t n, const double* in, double* out, double* state,
                       double* coeffs) {
  std::vector<IIRNotch> f;
  for (int i = 0; i < n_stages; i++) {
    f.emplace_back(notch_freq * std::pow(2, i), fs);
    if (state) { f[i].x(0) = state[4 * i]; f[i].x(1) = state[4 * i + 1]; f[i].y(0) = state[4 * i + 2]; f[i].y(1) = state[4 * i + 3]; }
    if (coeffs) for (int k = 0; k < 3; k++) { coeffs[6 * i + k] = f[i].b(k); coeffs[6 * i + 3 + k] = f[i].a(k); }
  }
  for (int64_t k = 0; k < n; k++) {
    double v = in[k];
    for (int i = 0; i < n_stages; i++) v = f[i].processSample(v);
    out[k] = v;
  }
  if (state)
    for (int i = 0; i < n_stages; i++) { state[4 * i] = f[i].x(0); state[4 * i + 1] = f[i].x(1); state[4 * i + 2] = f[i].y(0); state[4 * i + 3] = f[i].y(1); }
}

// ---- the reference's noise identification, state-estimator/src/noise_id/noise_id.cpp:9-65, compiled unmodified ----
// Same signature as oracle_capi.cpp's restatement: truth history [T1][21], [T1][4], [T1][441] (column-major RBIM per row).
double orc_noise_id_neg_loglik(int64_t T1, const double* vec, const double* quat, const double* cov, double dt, double q_gyro,
                               double q_accel, int N_window, int n_active, const int32_t* active, int64_t* n_windows, double* errs) {
  std::list<RBIS> states;
  RBIMList covs;
  for (int64_t t = 0; t < T1; t++) {
    states.push_back(makeState(vec + 21 * t, quat + 4 * t));
    covs.push_back(getCov(cov + 441 * t));
  }
  std::list<RBIS> errors, rolled;
  RBIMList ecovs, rolled_covs;
  sampleProcessForward(states, covs, dt, q_gyro, q_accel, N_window, errors, ecovs, rolled, rolled_covs);
  if (n_windows) *n_windows = (int64_t)errors.size();
  if (errs) {
    size_t w = 0;
    for (const RBIS& e : errors) { for (int i = 0; i < 21; i++) errs[21 * w + i] = e.vec(i); w++; }
  }
  Eigen::VectorXi act(n_active);
  for (int i = 0; i < n_active; i++) act(i) = active[i];
  return negLogLikelihood(errors, ecovs, act);
}

// ---- wire structs: rbisCreateFilterStateMessage (rbis.cpp:268-285) and RBIS(const pronto_filter_state_t*) (rbis.hpp:58-67) ----
// msg_out: utime, quat[4], num_states, state[21], num_cov_elements, cov[441] flattened into doubles
// [0] utime, [1..4] quat, [5] num_states, [6..26] state, [27] num_cov_elements, [28..468] cov
void orc_create_filter_state_message(const double* vec, const double* quat, int64_t utime, const double* cov, double* msg_out) {
  RBIS s = makeState(vec, quat);
  s.utime = utime;
  pronto_filter_state_t* m = rbisCreateFilterStateMessage(s, getCov(cov));
  msg_out[0] = (double)m->utime;
  for (int i = 0; i < 4; i++) msg_out[1 + i] = m->quat[i];
  msg_out[5] = m->num_states;
  for (int i = 0; i < m->num_states; i++) msg_out[6 + i] = m->state[i];
  msg_out[27] = m->num_cov_elements;
  for (int i = 0; i < m->num_cov_elements; i++) msg_out[28 + i] = m->cov[i];
  free(m->state); free(m->cov); free(m);
}
void orc_rbis_from_filter_state(const double* msg, double* vec, double* quat, int64_t* utime) {
  pronto_filter_state_t m;
  m.utime = (int64_t)msg[0];
  for (int i = 0; i < 4; i++) m.quat[i] = msg[1 + i];
  m.num_states = (int32_t)msg[5];
  std::vector<double> st(msg + 6, msg + 27), cv(msg + 28, msg + 469);
  m.state = st.data(); m.num_cov_elements = (int32_t)msg[27]; m.cov = cv.data();
  RBIS s(&m);
  putState(s, vec, quat);
  *utime = s.utime;
}

// ---- IMUStream::convertFromLCMBatch (estimate_tools/src/estimate_tools/imu_stream.cpp:62-97), stateful over calls ----
// packets [n][8]: utime, packet_count, delta_rotation[3], linear_acceleration[3] in MESSAGE order (newest first).
// out_new / out_old [cap][10]: utime_raw, utime_batch, utime, utime_delta, packet_count + 0 pad... see below
static IMUStream* g_stream = nullptr;
void orc_kvh_reset() { delete g_stream; g_stream = new IMUStream(); }
int orc_kvh_decode(int64_t batch_utime, int n, const double* packets, double* out_new, int* n_new, double* out_old, int* n_old) {
  if (!g_stream) g_stream = new IMUStream();
  bot_core::kvh_raw_imu_batch_t msg;
  msg.utime = batch_utime;
  msg.num_packets = n;
  for (int i = 0; i < n; i++) {
    bot_core::kvh_raw_imu_t p;
    p.utime = (int64_t)packets[8 * i]; p.packet_count = (int64_t)packets[8 * i + 1];
    for (int k = 0; k < 3; k++) { p.delta_rotation[k] = packets[8 * i + 2 + k]; p.linear_acceleration[k] = packets[8 * i + 5 + k]; }
    msg.raw_imu.push_back(p);
  }
  IMUBatch b = g_stream->convertFromLCMBatch(&msg);
  auto put = [](const IMUPacket& p, double* o) {
    o[0] = (double)p.utime_raw; o[1] = (double)p.utime_batch; o[2] = (double)p.utime; o[3] = (double)p.utime_delta; o[4] = (double)p.packet_count;
    for (int k = 0; k < 3; k++) { o[5 + k] = p.delta_rotation[k]; o[8 + k] = p.linear_acceleration[k]; }
  };
  *n_new = (int)b.packets.size();
  *n_old = (int)b.packets_old.size();
  for (int i = 0; i < *n_new; i++) put(b.packets[i], out_new + 11 * i);
  for (int i = 0; i < *n_old; i++) put(b.packets_old[i], out_old + 11 * i);
  return g_stream->getNumberBatchSinceReset();
}

}  // extern "C"

// ---- leg-odometry measurement formation: LegOdoCommon::createMeasurement (motion_estimate/src/mav_est_legodo/
// rbis_legodo_common.cpp:110-170 with getCovariance :35-88 and pronto::getDeltaAsVelocity, pronto_conversions_lcm.hpp:38-87),
// compiled unmodified.  The constructor's BotParam reads are served from these process-wide variables. ----
extern "C" {
const char* rbis_ref_shim_legodo_mode = "lin_rate";
double rbis_ref_shim_legodo_r[5] = {0, 0, 0, 0, 0};
char* rbis_ref_shim_param_str(const char* key) {
  (void)key;  // the only string key is state_estimator.legodo.mode
  return strdup(rbis_ref_shim_legodo_mode);
}
double rbis_ref_shim_param_double(const char* key) {
  const char* names[5] = {"state_estimator.legodo.r_xyz", "state_estimator.legodo.r_vxyz", "state_estimator.legodo.r_vang",
                          "state_estimator.legodo.r_vxyz_uncertain", "state_estimator.legodo.r_vang_uncertain"};
  for (int k = 0; k < 5; k++)
    if (!strcmp(key, names[k])) return rbis_ref_shim_legodo_r[k];
  fprintf(stderr, "ref shim: unknown parameter %s\n", key);
  abort();
}
// mode: "lin_rate" / "lin_rot_rate" / "pos_and_lin_rate"; r[5] as above; position / delta: translation xyz and quaternion wxyz.
// Outputs: m, idx[m], z[m], cov[m*m] (column-major); returns the sensor id, or -1 when the reference returned NULL.
int orc_legodo_create_measurement(const char* mode, const double* r, const double* pos_xyz, const double* pos_quat, const double* delta_xyz,
                                  const double* delta_quat, int64_t utime, int64_t prev_utime, int odo_position_status, float odo_delta_status,
                                  int* m_out, int* idx_out, double* z_out, double* cov_out, int64_t* utime_out) {
  rbis_ref_shim_legodo_mode = mode;
  for (int k = 0; k < 5; k++) rbis_ref_shim_legodo_r[k] = r[k];
  std::streambuf* old = std::cout.rdbuf(nullptr);  // the constructor announces its mode on stdout
  LegOdoCommon common(nullptr, nullptr, nullptr);
  std::cout.rdbuf(old);
  BotTrans posT, deltaT;
  for (int k = 0; k < 3; k++) { posT.trans_vec[k] = pos_xyz[k]; deltaT.trans_vec[k] = delta_xyz[k]; }
  for (int k = 0; k < 4; k++) { posT.rot_quat[k] = pos_quat[k]; deltaT.rot_quat[k] = delta_quat[k]; }
  RBISUpdateInterface* u = common.createMeasurement(posT, deltaT, utime, prev_utime, odo_position_status, odo_delta_status);
  if (!u) return -1;
  RBISIndexedMeasurement* im = dynamic_cast<RBISIndexedMeasurement*>(u);
  const int m = (int)im->index.rows();
  *m_out = m;
  for (int a = 0; a < m; a++) { idx_out[a] = im->index(a); z_out[a] = im->measurement(a); }
  for (int b = 0; b < m; b++)
    for (int a = 0; a < m; a++) cov_out[a + m * b] = im->measurement_cov(a, b);
  *utime_out = u->utime;
  const int sensor = (int)u->sensor_id;
  delete u;
  return sensor;
}
}
