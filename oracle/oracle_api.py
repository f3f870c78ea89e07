"""ctypes face of the CPU ORACLE (oracle/_build/librbis_oracle.so).

Test infrastructure only (how its parity is pinned: see rbis_oracle.hpp): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by pronto_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "librbis_oracle.so")
_lib = None


class _Stream(C.Structure):
    _fields_ = [("m", C.c_int32), ("has_orient", C.c_int32), ("r_mode", C.c_int32), ("sensor_id", C.c_int32),
                ("idx", C.c_int32 * 9), ("_pad", C.c_int32), ("z", C.c_void_p), ("quat", C.c_void_p), ("R", C.c_void_p)]


EVENT_DTYPE = np.dtype([("kind", "<i4"), ("stream", "<i4"), ("row", "<i8"), ("utime", "<i8"), ("dt", "<f8")], align=True)


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, no fast-math)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        vp, d = C.c_void_p, C.c_double
        lib.orc_set_constants.argtypes = [d, d, C.c_int]
        lib.orc_linearization.argtypes = [vp, vp, vp]
        lib.orc_ins_update_state.argtypes = [vp, vp, d, vp, vp]
        lib.orc_ins_update_covariance.argtypes = [d, d, d, d, vp, vp, vp, d]
        lib.orc_indexed_measurement.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.orc_indexed_measurement.restype = d
        lib.orc_apply_delta.argtypes = [vp] * 9
        lib.orc_subtract_quats.argtypes = [vp, vp, vp]
        lib.orc_state_error.argtypes = [vp, vp, vp, vp, vp]
        lib.orc_run_ensemble.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int,
                                         C.POINTER(_Stream), C.c_int64, vp, C.c_int64, vp, vp, vp, vp]
        lib.orc_run_ensemble.restype = C.c_int64
        _lib = lib
    return _lib


REF_LIB_PATH = os.path.join(_HERE, "_ref", "librbis_ref.so")
REF_ROOT = "/root/reference"
_ref_lib = None


def build_ref(force=False):
    """oracle/_ref/librbis_ref.so = the reference's OWN rbis.cpp compiled unmodified against oracle/ref_shim/ (see
    oracle/Makefile).  Needs the reference tree; returns None where neither the tree nor a prebuilt library exists."""
    if os.path.isdir(os.path.join(REF_ROOT, "state-estimator")) and (force or not os.path.exists(REF_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "ref"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return REF_LIB_PATH if os.path.exists(REF_LIB_PATH) else None


class reference:
    """Context manager: inside it the single-update functions of this module (linearization, ins_update_state,
    ins_update_covariance, measurement_update, set_constants) run the REFERENCE's compiled rbis.cpp instead of the
    oracle's restatement."""

    def __enter__(self):
        global _lib, _ref_lib
        if _ref_lib is None:
            path = build_ref()
            if path is None:
                raise FileNotFoundError("oracle/_ref/librbis_ref.so is not built and /root/reference is not mounted")
            lib = C.CDLL(path)
            vp, d = C.c_void_p, C.c_double
            lib.orc_set_constants.argtypes = [d, d, C.c_int]
            lib.orc_linearization.argtypes = [vp, vp, vp]
            lib.orc_ins_update_state.argtypes = [vp, vp, d, vp, vp]
            lib.orc_ins_update_covariance.argtypes = [d, d, d, d, vp, vp, vp, d]
            lib.orc_indexed_measurement.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
            lib.orc_indexed_measurement.restype = d
            lib.orc_apply_delta.argtypes = [vp] * 9
            lib.orc_run_ensemble.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int,
                                             C.POINTER(_Stream), C.c_int64, vp, C.c_int64, vp, vp, vp, vp]
            lib.orc_run_ensemble.restype = C.c_int64
            _ref_lib = lib
        load()
        self._saved = _lib
        _lib = _ref_lib
        return self

    def __exit__(self, *a):
        global _lib
        _lib = self._saved


def _a(x, n=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if n is not None:
        assert x.size == n, (x.shape, n)
    return x


def set_constants(g_val=9.8, chi_tol=1e-6, ctor_folds_chi=True):
    load().orc_set_constants(float(g_val), float(chi_tol), int(bool(ctor_folds_chi)))


def linearization(vec, quat):
    """getIMUProcessLinearizationContinuous -> Ac [21,21]"""
    vec, quat = _a(vec, 21), _a(quat, 4)
    Ac = np.empty(441)
    load().orc_linearization(vec.ctypes.data, quat.ctypes.data, Ac.ctypes.data)
    return Ac.reshape(21, 21).T.copy()  # column-major -> [r, c]


def ins_update_state(gyro, accel, dt, vec, quat):
    gyro, accel = _a(gyro, 3), _a(accel, 3)
    vec, quat = _a(vec, 21).copy(), _a(quat, 4).copy()
    load().orc_ins_update_state(gyro.ctypes.data, accel.ctypes.data, float(dt), vec.ctypes.data, quat.ctypes.data)
    return vec, quat


def ins_update_covariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, vec, quat, cov, dt):
    """cov: [21,21] (row, col) -> new [21,21]"""
    vec, quat = _a(vec, 21), _a(quat, 4)
    P = np.ascontiguousarray(np.asarray(cov, dtype=np.float64).reshape(21, 21).T).reshape(-1).copy()
    load().orc_ins_update_covariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, vec.ctypes.data, quat.ctypes.data,
                                     P.ctypes.data, float(dt))
    return P.reshape(21, 21).T.copy()


def measurement_update(z, R, idx, vec, quat, cov, meas_quat=None):
    """indexed[PlusOrientation]Measurement + rbisApplyDelta -> (vec, quat, cov [21,21], loglik term)"""
    m = len(idx)
    z, vec, quat = _a(z, m), _a(vec, 21), _a(quat, 4)
    Rc = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(m, m).T).reshape(-1)
    idx_a = np.ascontiguousarray(idx, dtype=np.int32)
    P = np.ascontiguousarray(np.asarray(cov, dtype=np.float64).reshape(21, 21).T).reshape(-1)
    dvec, dquat, dcov = np.empty(21), np.empty(4), np.empty(441)
    mq = _a(meas_quat, 4) if meas_quat is not None else None
    ll = load().orc_indexed_measurement(m, z.ctypes.data, mq.ctypes.data if mq is not None else None, Rc.ctypes.data,
                                        idx_a.ctypes.data, vec.ctypes.data, quat.ctypes.data, P.ctypes.data,
                                        dvec.ctypes.data, dquat.ctypes.data, dcov.ctypes.data)
    pvec, pquat, pcov = np.empty(21), np.empty(4), np.empty(441)
    load().orc_apply_delta(vec.ctypes.data, quat.ctypes.data, P.ctypes.data, dvec.ctypes.data, dquat.ctypes.data,
                           dcov.ctypes.data, pvec.ctypes.data, pquat.ctypes.data, pcov.ctypes.data)
    return pvec, pquat, pcov.reshape(21, 21).T.copy(), ll


def subtract_quats(q1, q2):
    q1, q2 = _a(q1, 4), _a(q2, 4)
    out = np.empty(3)
    load().orc_subtract_quats(q1.ctypes.data, q2.ctypes.data, out.ctypes.data)
    return out


def state_error(vec, quat, tvec, tquat):
    vec, quat, tvec, tquat = _a(vec, 21), _a(quat, 4), _a(tvec, 21), _a(tquat, 4)
    out = np.empty(21)
    load().orc_state_error(vec.ctypes.data, quat.ctypes.data, tvec.ctypes.data, tquat.ctypes.data, out.ctypes.data)
    return out


def _marshal(vec, quat, cov, q_params, imu, streams, events):
    vec, quat, cov = _a(vec).copy(), _a(quat).copy(), _a(cov).copy()
    N = vec.shape[1]
    qs = [np.full(N, float(q)) if np.isscalar(q) else _a(q, N) for q in q_params]
    ev = np.zeros(len(events), dtype=EVENT_DTYPE)
    for i, e in enumerate(events):
        ev[i] = tuple(e)
    sarr = (_Stream * max(1, len(streams)))()
    keep = []
    for s, st in enumerate(streams):
        m = len(st["idx"])
        d = sarr[s]
        d.m, d.has_orient = m, int(st.get("quat") is not None)
        d.r_mode = 1 if st.get("per_filter_diag") else 0
        d.sensor_id = int(st.get("sensor_id", 11))  # RBISUpdateInterface::legodo; 0 would be `ins`
        for a, i in enumerate(st["idx"]):
            d.idx[a] = int(i)
        z = _a(st["z"]); keep.append(z); d.z = z.ctypes.data
        if st.get("quat") is not None:
            q = _a(st["quat"]); keep.append(q); d.quat = q.ctypes.data
        if st.get("per_filter_diag"):
            R = _a(st["R"], m * N)
        else:
            R = np.ascontiguousarray(np.asarray(st["R"], dtype=np.float64).reshape(m, m).T).reshape(-1)
        keep.append(R); d.R = R.ctypes.data
    imu_a = _a(imu) if imu is not None else None
    return vec, quat, cov, N, qs, ev, sarr, keep, imu_a


def run_ensemble(vec, quat, cov, loglik, utime0, q_params, imu, streams, events, history_span=10_000_000,
                 n_threads=1, trace=False):
    """Replay an ARRIVAL-ordered event list through one MavStateEstimator per filter.

    vec [21][N], quat [4][N], cov [441][N], loglik [N] (copied; results returned).
    q_params: 4 arrays [N] or scalars.  imu [rows][6][N] or None.
    streams: list of dicts(idx, z [rows][m][N], R (m x m shared, or [m][N] with per_filter_diag=True),
             quat [rows][4][N] or None).
    events: iterable of (kind, stream, row, utime, dt); kind 0 = IMU, 1 = measurement.
    Returns dict(vec, quat, cov, loglik, calls[, trace_vec, trace_quat, trace_cov, trace_loglik])."""
    lib = load()
    vec, quat, cov, N, qs, ev, sarr, keep, imu_a = _marshal(vec, quat, cov, q_params, imu, streams, events)
    loglik = np.zeros(N) if loglik is None else _a(loglik).copy()
    E = len(ev)
    tr = [None] * 4
    if trace:
        tr = [np.empty((E, 21, N)), np.empty((E, 4, N)), np.empty((E, 441, N)), np.empty((E, N))]
    calls = lib.orc_run_ensemble(N, int(n_threads), vec.ctypes.data, quat.ctypes.data, cov.ctypes.data,
                                 loglik.ctypes.data, int(utime0), qs[0].ctypes.data, qs[1].ctypes.data,
                                 qs[2].ctypes.data, qs[3].ctypes.data,
                                 imu_a.ctypes.data if imu_a is not None else None, len(streams), sarr, E,
                                 ev.ctypes.data, int(history_span), *[t.ctypes.data if t is not None else None for t in tr])
    out = dict(vec=vec, quat=quat, cov=cov, loglik=loglik, calls=int(calls))
    if trace:
        out.update(trace_vec=tr[0], trace_quat=tr[1], trace_cov=tr[2], trace_loglik=tr[3])
    return out


def ekf_smoothing_step(next_pred, next, cur, dt):
    """ekfSmoothingStep (MSE/rbis.cpp:234-266).  Each argument is (vec [21], quat [4], cov [21,21]); returns the smoothed
    (vec, quat, cov [21,21]) of `cur`."""
    lib = load()
    lib.orc_ekf_smoothing_step.argtypes = [C.c_void_p] * 6 + [C.c_double] + [C.c_void_p] * 3
    lib.orc_ekf_smoothing_step.restype = None
    def cm(P):
        return np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(21, 21).T).reshape(-1).copy()
    npv, npq, npc = _a(next_pred[0], 21), _a(next_pred[1], 4), cm(next_pred[2])
    nv, nq, nc = _a(next[0], 21), _a(next[1], 4), cm(next[2])
    cv, cq, cc = _a(cur[0], 21).copy(), _a(cur[1], 4).copy(), cm(cur[2])
    lib.orc_ekf_smoothing_step(npv.ctypes.data, npq.ctypes.data, npc.ctypes.data, nv.ctypes.data, nq.ctypes.data, nc.ctypes.data,
                               float(dt), cv.ctypes.data, cq.ctypes.data, cc.ctypes.data)
    return cv, cq, cc.reshape(21, 21).T.copy()


def smooth_ensemble(vec, quat, cov, utime0, q_params, imu, streams, events, smooth_dt, n_threads=1):
    """Forward pass as run_ensemble (in-order events, nothing truncated), then MavStateEstimator::EKFSmoothBackwardsPass
    (MSE/mav_state_est.cpp:98-189).  Returns dict(post_vec [E][21][N], post_quat [E][4][N], post_cov [E][441][N]): every
    update's posterior after smoothing, in history order."""
    lib = load()
    vp = C.c_void_p
    lib.orc_smooth_ensemble.argtypes = [C.c_int64, C.c_int, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int,
                                        C.POINTER(_Stream), C.c_int64, vp, C.c_int64, C.c_double, vp, vp, vp]
    lib.orc_smooth_ensemble.restype = C.c_int64
    vec, quat, cov, N, qs, ev, sarr, keep, imu_a = _marshal(vec, quat, cov, q_params, imu, streams, events)
    E = len(ev)
    pv, pq, pc = np.empty((E, 21, N)), np.empty((E, 4, N)), np.empty((E, 441, N))
    lib.orc_smooth_ensemble(N, int(n_threads), vec.ctypes.data, quat.ctypes.data, cov.ctypes.data, int(utime0),
                            qs[0].ctypes.data, qs[1].ctypes.data, qs[2].ctypes.data, qs[3].ctypes.data,
                            imu_a.ctypes.data if imu_a is not None else None, len(streams), sarr, E, ev.ctypes.data,
                            2_000_000_000, float(smooth_dt), pv.ctypes.data, pq.ctypes.data, pc.ctypes.data)
    return dict(post_vec=pv, post_quat=pq, post_cov=pc)


def notch_cascade(x, notch_freq, fs=1000.0, n_stages=3, state=None):
    """InsHandler::doFilter on one channel (MSE/sensor_handlers.cpp:155-162 over estimate_tools' IIRNotch): n_stages notch
    filters at notch_freq * 2^i in cascade.  Returns (y, state [n_stages][4] = x0,x1,y0,y1 per stage, coeffs [n_stages][6])."""
    lib = load()
    lib.orc_notch_cascade.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_notch_cascade.restype = None
    x = _a(x)
    y = np.empty_like(x)
    st = np.zeros((n_stages, 4)) if state is None else _a(state, 4 * n_stages).reshape(n_stages, 4).copy()
    co = np.zeros((n_stages, 6))
    lib.orc_notch_cascade(float(notch_freq), float(fs), int(n_stages), x.size, x.ctypes.data, y.ctypes.data, st.ctypes.data, co.ctypes.data)
    return y, st, co


def noise_id_neg_loglik(vec, quat, cov, dt, q_gyro, q_accel, n_window, active=(3, 4, 5, 6, 7, 8, 9, 10, 11)):
    """state-estimator/src/noise_id/noise_id.cpp:9-65 for ONE (q_gyro, q_accel): truth history vec [T1][21],
    quat [T1][4], cov [T1][441] (column-major RBIM per row) -> (negative log-likelihood, per-window errors [W][21])."""
    lib = load()
    lib.orc_noise_id_neg_loglik.restype = C.c_double
    lib.orc_noise_id_neg_loglik.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    vec, quat, cov = _a(vec), _a(quat), _a(cov)
    T1 = vec.shape[0]
    act = np.ascontiguousarray(active, dtype=np.int32)
    nw = C.c_int64(0)
    errs = np.zeros((max(1, (T1 - 1) // int(n_window)), 21))
    v = lib.orc_noise_id_neg_loglik(T1, vec.ctypes.data, quat.ctypes.data, cov.ctypes.data, float(dt), float(q_gyro), float(q_accel),
                                    int(n_window), len(act), act.ctypes.data, C.byref(nw), errs.ctypes.data)
    return float(v), errs[:nw.value]
