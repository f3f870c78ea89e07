// oracle_capi.cpp -- extern "C" face of the CPU ORACLE for ctypes (test infrastructure, NOT product
// code; PARITY UNPINNED, see rbis_oracle.hpp).  Flat double arrays; matrices column-major.
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "rbis_oracle.hpp"

using namespace rbis_oracle;

namespace {
RBIS makeState(const double* vec, const double* quat) {
  Quat q;
  q.w = quat[0]; q.x = quat[1]; q.y = quat[2]; q.z = quat[3];
  return RBIS(vec, q);
}
void putState(const RBIS& s, double* vec, double* quat) {
  std::memcpy(vec, s.vec, sizeof(double) * N);
  quat[0] = s.quat.w; quat[1] = s.quat.x; quat[2] = s.quat.y; quat[3] = s.quat.z;
}
RBIM getCov(const double* cov) {
  RBIM P;
  std::memcpy(P.m, cov, sizeof(P.m));
  return P;
}
void putCov(const RBIM& P, double* cov) { std::memcpy(cov, P.m, sizeof(P.m)); }
}  // namespace

extern "C" {

void orc_set_constants(double g_val, double chi_tol, int ctor_folds_chi) {
  constants().g_val = g_val;
  constants().chi_tol = chi_tol;
  constants().ctor_folds_chi = ctor_folds_chi != 0;
}

void orc_linearization(const double* vec, const double* quat, double* Ac) {
  RBIM A;
  getIMUProcessLinearizationContinuous(makeState(vec, quat), A);
  std::memcpy(Ac, A.m, sizeof(A.m));
}

void orc_ins_update_state(const double* gyro, const double* accel, double dt, double* vec, double* quat) {
  RBIS s = makeState(vec, quat);
  insUpdateState(gyro, accel, dt, s);
  putState(s, vec, quat);
}

void orc_ins_update_covariance(double q_gyro, double q_accel, double q_gyro_bias, double q_accel_bias,
                               const double* vec, const double* quat, double* cov, double dt) {
  RBIM P;
  std::memcpy(P.m, cov, sizeof(P.m));
  insUpdateCovariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, makeState(vec, quat), P, dt);
  std::memcpy(cov, P.m, sizeof(P.m));
}

// indexed (meas_quat == NULL) or indexed-plus-orientation measurement; returns the log-likelihood
// term and writes dstate (vec + quat) and dcov.
double orc_indexed_measurement(int m, const double* z, const double* meas_quat, const double* R, const int32_t* idx,
                               const double* vec, const double* quat, const double* cov, double* dvec,
                               double* dquat, double* dcov) {
  RBIM P, dP;
  std::memcpy(P.m, cov, sizeof(P.m));
  RBIS ds;
  double ll;
  if (meas_quat) {
    Quat q;
    q.w = meas_quat[0]; q.x = meas_quat[1]; q.y = meas_quat[2]; q.z = meas_quat[3];
    ll = indexedPlusOrientationMeasurement(m, z, q, R, idx, makeState(vec, quat), P, ds, dP);
  } else {
    ll = indexedMeasurement(m, z, R, idx, makeState(vec, quat), P, ds, dP);
  }
  putState(ds, dvec, dquat);
  std::memcpy(dcov, dP.m, sizeof(dP.m));
  return ll;
}

void orc_apply_delta(const double* vec, const double* quat, const double* cov, const double* dvec,
                     const double* dquat, const double* dcov, double* pvec, double* pquat, double* pcov) {
  RBIM P, dP, Pp;
  std::memcpy(P.m, cov, sizeof(P.m));
  std::memcpy(dP.m, dcov, sizeof(dP.m));
  RBIS post;
  rbisApplyDelta(makeState(vec, quat), P, makeState(dvec, dquat), dP, post, Pp);
  putState(post, pvec, pquat);
  std::memcpy(pcov, Pp.m, sizeof(Pp.m));
}

void orc_subtract_quats(const double* q1, const double* q2, double* out) {
  Quat a, b;
  a.w = q1[0]; a.x = q1[1]; a.y = q1[2]; a.z = q1[3];
  b.w = q2[0]; b.x = q2[1]; b.y = q2[2]; b.z = q2[3];
  subtractQuats(a, b, out);
}

// state error as SE/noise_id/noise_id.cpp:37-38:  e = est; e.subtractState(truth); e.quatToChi()
void orc_state_error(const double* vec, const double* quat, const double* tvec, const double* tquat, double* err) {
  RBIS e = makeState(vec, quat);
  e.subtractState(makeState(tvec, tquat));
  e.quatToChi();
  std::memcpy(err, e.vec, sizeof(double) * N);
}

// ------------------------------------------------------------------------------------------------
// ensemble runner: every filter replays the same ARRIVAL-ORDERED event list through its own
// MavStateEstimator (multimap history, heap-allocated update objects, as the reference does).
// ------------------------------------------------------------------------------------------------
typedef struct {
  int32_t m;
  int32_t has_orient;
  int32_t r_mode;     // 0: R shared, m x m column-major;  1: per-filter diagonal, [m][N]
  int32_t sensor_id;
  int32_t idx[9];
  int32_t _pad;
  const double* z;     // [rows][m][N]
  const double* quat;  // [rows][4][N] or NULL
  const double* R;
} orc_stream_t;

typedef struct {
  int32_t kind;    // 0 IMU, 1 measurement stream
  int32_t stream;
  int64_t row;
  int64_t utime;
  double dt;       // IMU only
} orc_event_t;

// vec [21][N], quat [4][N], cov [441][N], loglik [N]: in = initial (utime0), out = head after the run.
// qparams: 4 arrays [N] (q_gyro, q_accel, q_gyro_bias, q_accel_bias).  imu: [rows][6][N].
// trace_* (optional, may be NULL): head after EVERY event, [E][k][N].
// Returns total number of updateFilter calls across all filters (replays included).
static int64_t run_ensemble_impl(int64_t Nf, int n_threads, double* vec, double* quat, double* cov, double* loglik,
                         int64_t utime0, const double* q_gyro, const double* q_accel, const double* q_gyro_bias,
                         const double* q_accel_bias, const double* imu, int n_streams, const orc_stream_t* streams,
                         int64_t n_events, const orc_event_t* events, int64_t history_span, double* trace_vec,
                         double* trace_quat, double* trace_cov, double* trace_loglik, double smooth_dt,
                         double* post_vec, double* post_quat, double* post_cov) {
  (void)n_streams;
  std::atomic<int64_t> next(0), calls(0);
  auto worker = [&]() {
    int64_t my_calls = 0;
    for (;;) {
      const int64_t n = next.fetch_add(1);
      if (n >= Nf) break;
      double v[N], q[4];
      RBIM P;
      for (int i = 0; i < N; i++) v[i] = vec[i * Nf + n];
      for (int i = 0; i < 4; i++) q[i] = quat[i * Nf + n];
      for (int i = 0; i < N * N; i++) P.m[i] = cov[i * Nf + n];
      RBIS s0 = makeState(v, q);
      s0.utime = utime0;
      MavStateEstimator est(new RBISResetUpdate(s0, P, RBISUpdateInterface::reset, utime0), history_span);
      // the reset zeroes the log-likelihood (rbis_update_interface.cpp:27); carry an initial offset
      const double ll0 = loglik ? loglik[n] : 0.0;
      for (int64_t e = 0; e < n_events; e++) {
        const orc_event_t& ev = events[e];
        RBISUpdateInterface* u;
        if (ev.kind == 0) {
          double g[3], a[3];
          for (int i = 0; i < 3; i++) {
            g[i] = imu[(ev.row * 6 + i) * Nf + n];
            a[i] = imu[(ev.row * 6 + 3 + i) * Nf + n];
          }
          u = new RBISIMUProcessStep(g, a, q_gyro[n], q_accel[n], q_gyro_bias[n], q_accel_bias[n], ev.dt, ev.utime);
        } else {
          const orc_stream_t& st = streams[ev.stream];
          const int m = st.m;
          double z[9], R[81];
          for (int i = 0; i < m; i++) z[i] = st.z[(ev.row * m + i) * Nf + n];
          if (st.r_mode == 0) {
            std::memcpy(R, st.R, sizeof(double) * m * m);
          } else {
            std::memset(R, 0, sizeof(double) * m * m);
            for (int i = 0; i < m; i++) R[i + m * i] = st.R[i * Nf + n];
          }
          if (st.has_orient) {
            Quat mq;
            mq.w = st.quat[(ev.row * 4 + 0) * Nf + n]; mq.x = st.quat[(ev.row * 4 + 1) * Nf + n];
            mq.y = st.quat[(ev.row * 4 + 2) * Nf + n]; mq.z = st.quat[(ev.row * 4 + 3) * Nf + n];
            u = new RBISIndexedPlusOrientationMeasurement(m, st.idx, z, R, mq,
                                                          (RBISUpdateInterface::sensor_enum)st.sensor_id, ev.utime);
          } else {
            u = new RBISIndexedMeasurement(m, st.idx, z, R, (RBISUpdateInterface::sensor_enum)st.sensor_id, ev.utime);
          }
        }
        est.addUpdate(u, true);
        if (trace_vec || trace_quat || trace_cov || trace_loglik) {
          RBIS hs;
          RBIM hP;
          est.getHeadState(hs, hP);
          if (trace_vec) for (int i = 0; i < N; i++) trace_vec[(e * N + i) * Nf + n] = hs.vec[i];
          if (trace_quat) {
            trace_quat[(e * 4 + 0) * Nf + n] = hs.quat.w; trace_quat[(e * 4 + 1) * Nf + n] = hs.quat.x;
            trace_quat[(e * 4 + 2) * Nf + n] = hs.quat.y; trace_quat[(e * 4 + 3) * Nf + n] = hs.quat.z;
          }
          if (trace_cov) for (int i = 0; i < N * N; i++) trace_cov[((int64_t)e * N * N + i) * Nf + n] = hP.m[i];
          if (trace_loglik) trace_loglik[e * Nf + n] = ll0 + est.getMeasurementsLogLikelihood();
        }
      }
      if (smooth_dt > 0) {  // mav_state_est.cpp:98-189, then every update's posterior in history order (reset excluded)
        est.EKFSmoothBackwardsPass(smooth_dt);
        int64_t e = -1;
        for (auto& kv : est.history.updateMap) {
          if (e >= 0) {
            double pv[N], pq[4];
            putState(kv.second->posterior_state, pv, pq);
            for (int i = 0; i < N; i++) post_vec[(e * N + i) * Nf + n] = pv[i];
            for (int i = 0; i < 4; i++) post_quat[(e * 4 + i) * Nf + n] = pq[i];
            for (int i = 0; i < N * N; i++) post_cov[((int64_t)e * N * N + i) * Nf + n] = kv.second->posterior_covariance.m[i];
          }
          e++;
        }
      }
      RBIS hs;
      RBIM hP;
      est.getHeadState(hs, hP);
      double ov[N], oq[4];
      putState(hs, ov, oq);
      for (int i = 0; i < N; i++) vec[i * Nf + n] = ov[i];
      for (int i = 0; i < 4; i++) quat[i * Nf + n] = oq[i];
      for (int i = 0; i < N * N; i++) cov[i * Nf + n] = hP.m[i];
      if (loglik) loglik[n] = ll0 + est.getMeasurementsLogLikelihood();
      my_calls += est.n_update_calls;
    }
    calls += my_calls;
  };
  if (n_threads <= 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  return calls.load();
}

// noise identification (state-estimator/src/noise_id/noise_id.cpp:9-65): truth history [T1][21], [T1][4], [T1][441]
// (row-major over time), one (q_gyro, q_accel) point -> negative log-likelihood over the complete windows.
// errs/covs (optional): per-window state error [W][21] and active covariance difference [W][n_active^2].
double orc_noise_id_neg_loglik(int64_t T1, const double* vec, const double* quat, const double* cov, double dt, double q_gyro,
                               double q_accel, int N_window, int n_active, const int32_t* active, int64_t* n_windows,
                               double* errs) {
  std::vector<RBIS> states((size_t)T1);
  std::vector<RBIM> covs((size_t)T1);
  for (int64_t t = 0; t < T1; t++) {
    states[(size_t)t] = makeState(vec + 21 * t, quat + 4 * t);
    std::memcpy(covs[(size_t)t].m, cov + 441 * t, sizeof(double) * 441);
  }
  std::vector<RBIS> errors;
  std::vector<RBIM> ecovs;
  sampleProcessForward(states, covs, dt, q_gyro, q_accel, N_window, errors, ecovs);
  if (n_windows) *n_windows = (int64_t)errors.size();
  if (errs)
    for (size_t w = 0; w < errors.size(); w++) std::memcpy(errs + 21 * w, errors[w].vec, sizeof(double) * 21);
  return negLogLikelihood(errors, ecovs, n_active, active);
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

// InsHandler::doFilter (sensor_handlers.cpp:155-162) on ONE channel: n_stages notch filters at notch_freq * 2^i, fs,
// in cascade over n samples; state (x0,x1,y0,y1 per stage) in/out so that calls can be chained.  coeffs (optional): b,a per stage.
void orc_notch_cascade(double notch_freq, double fs, int n_stages, int64_t n, const double* in, double* out, double* state,
                       double* coeffs) {
  std::vector<IIRNotch> f;
  for (int i = 0; i < n_stages; i++) {
    f.emplace_back(notch_freq * std::pow(2, i), fs);
    if (state) { f[i].x[0] = state[4 * i]; f[i].x[1] = state[4 * i + 1]; f[i].y[0] = state[4 * i + 2]; f[i].y[1] = state[4 * i + 3]; }
    if (coeffs) for (int k = 0; k < 3; k++) { coeffs[6 * i + k] = f[i].b[k]; coeffs[6 * i + 3 + k] = f[i].a[k]; }
  }
  for (int64_t k = 0; k < n; k++) {
    double v = in[k];
    for (int i = 0; i < n_stages; i++) v = f[i].processSample(v);
    out[k] = v;
  }
  if (state)
    for (int i = 0; i < n_stages; i++) { state[4 * i] = f[i].x[0]; state[4 * i + 1] = f[i].x[1]; state[4 * i + 2] = f[i].y[0]; state[4 * i + 3] = f[i].y[1]; }
}

}  // extern "C"
