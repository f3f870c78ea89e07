"""IMU-noise identification on the batch API: the reference's windowed likelihood
(state-estimator/src/noise_id/noise_id.cpp:9-65, driven per parameter point by the Matlab mex
state-estimator/matlab/noiseParamLikelihoodMex.cpp:87-140) for a whole grid of (q_gyro, q_accel) points in
one pass.  SURVEY.md 8f row 2.

Every (parameter point g, window w) is one filter: it starts at the truth state/covariance of the window
start, runs n_window IMU process steps driven by the truth's own angular velocity / acceleration entries
(noise_id.cpp:26) with its point's process noise (bias noise 0, :13-14), and is compared with the truth at
the window end.  All points of a window read the SAME input column (rbis_batch_set_column_map); the zero-noise
covariance of noise_id.cpp:25 is one extra filter per window.  All arithmetic runs in librbis_b200.so.
"""
import numpy as np

from . import capi
from .batch import RBISBatch, make_ops

ACTIVE_DEFAULT = (3, 4, 5, 6, 7, 8, 9, 10, 11)  # velocity, chi, position (SE/noise_id/roll_forward.cpp:54-57)


def neg_log_likelihood(truth_vec, truth_quat, truth_cov, dt, q_gyro, q_accel, n_window, active=ACTIVE_DEFAULT, device=0,
                       return_terms=False):
    """truth_vec [T1][21], truth_quat [T1][4], truth_cov [T1][441] (RBIM column-major per row): the filter-state
    history the reference loads from a log (loadFilterHistory, noise_id.cpp:67-91).
    q_gyro, q_accel: arrays [G] of process-noise variances (as sampleProcessForward takes them).
    Returns nll [G] (negLogLikelihood per parameter point); with return_terms also the [G][W] window terms."""
    import ctypes as C

    truth_vec = np.ascontiguousarray(truth_vec, dtype=np.float64)
    truth_quat = np.ascontiguousarray(truth_quat, dtype=np.float64)
    truth_cov = np.ascontiguousarray(truth_cov, dtype=np.float64)
    q_gyro = np.atleast_1d(np.asarray(q_gyro, dtype=np.float64))
    q_accel = np.atleast_1d(np.asarray(q_accel, dtype=np.float64))
    if q_gyro.shape != q_accel.shape or q_gyro.ndim != 1:
        raise ValueError("q_gyro and q_accel must be 1-d arrays of equal length")
    T1, G, Nw = truth_vec.shape[0], q_gyro.shape[0], int(n_window)
    W = (T1 - 1) // Nw  # complete windows: the state after the last step of the window must exist (noise_id.cpp:31-38)
    if W < 1:
        raise ValueError("truth history shorter than one window")
    starts = np.arange(W) * Nw
    ends = starts + Nw
    # shared inputs: column w, row i = truth (omega, a) at index w*Nw + i   (noise_id.cpp:26)
    idx = starts[None, :] + np.arange(Nw)[:, None]                                  # [Nw][W]
    imu = np.ascontiguousarray(np.concatenate([truth_vec[idx][:, :, 0:3], truth_vec[idx][:, :, 12:15]], axis=2).transpose(0, 2, 1))
    ops = make_ops([(capi.OP_IMU, 0, i, (i + 1) * 1000, float(dt)) for i in range(Nw)])
    sv = np.ascontiguousarray(truth_vec[starts].T)      # [21][W]
    sq = np.ascontiguousarray(truth_quat[starts].T)     # [4][W]
    sc = np.ascontiguousarray(truth_cov[starts].T)      # [441][W]
    # ---- zero-noise covariance per window (noise_id.cpp:25) ----
    with RBISBatch(W, device=device) as b0:
        b0.set_process_noise(0.0, 0.0, 0.0, 0.0)
        b0.set_state(sv, sq, sc)
        b0.run_fused(ops, imu=imu)
        base_cov = b0.get_state()[2]                    # [441][W]
    # ---- the grid: filter n = g * W + w ----
    N = G * W
    wmap = np.tile(np.arange(W, dtype=np.int32), G)
    rep = lambda a: np.ascontiguousarray(a[:, wmap])
    act = np.ascontiguousarray(active, dtype=np.int32)
    out = np.empty(N)
    with RBISBatch(N, device=device) as b:
        b.set_process_noise(np.repeat(q_gyro, W), np.repeat(q_accel, W), np.zeros(N), np.zeros(N))
        b.set_state(rep(sv), rep(sq), rep(sc))
        b.set_column_map(-1, wmap, W)
        b.run_fused(ops, imu=imu)
        tv, tq = rep(np.ascontiguousarray(truth_vec[ends].T)), rep(np.ascontiguousarray(truth_quat[ends].T))
        capi.check(b.lib.rbis_batch_window_neg_loglik(b.h, tv.ctypes.data, tq.ctypes.data, base_cov.ctypes.data, wmap.ctypes.data, W,
                                                      len(act), act.ctypes.data, out.ctypes.data, capi.MEM_HOST))
    terms = out.reshape(G, W)
    nll = np.zeros(G)
    for w in range(W):  # fixed order, as the reference accumulates window by window (noise_id.cpp:52-62)
        nll += terms[:, w]
    return (nll, terms) if return_terms else nll
