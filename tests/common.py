"""Shared scenario builders for the tests (host-side numpy only)."""
import numpy as np

from pronto_b200 import synth


def random_ensemble(N, seed=0, spd_scale=0.01):
    """Random but plausible states and dense SPD covariances, SoA layout."""
    rng = np.random.default_rng(seed)
    vec = rng.normal(size=(21, N)) * 0.3
    vec[6:9] = 0.0
    quat = rng.normal(size=(4, N))
    quat /= np.linalg.norm(quat, axis=0, keepdims=True)
    cov = np.empty((441, N))
    for n in range(N):
        A = rng.normal(size=(21, 21))
        P = A @ A.T * spd_scale / 21 + np.eye(21) * spd_scale
        cov[:, n] = P.T.reshape(-1)  # element r + 21 c
    return np.ascontiguousarray(vec), np.ascontiguousarray(quat), cov


def nominal_q():
    p = synth.NOMINAL
    return (p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])


def scenario(N, T, k0=0, tumbling=False, with_legodo=True, with_pose=True, seed=synth.SEED, truth=None):
    """Initial ensemble + streams + arrival-ordered events for steps k0..k0+T-1 (config-3 schedule)."""
    truth = synth.truth_trajectory(k0 + T, tumbling=tumbling) if truth is None else truth
    tv = np.zeros(21)
    tv[9:12] = (0, 0, 0.85)
    tv[15:18] = synth.NOMINAL["bg"]
    tv[18:21] = synth.NOMINAL["ba"]
    tq = np.array([1.0, 0, 0, 0])
    if k0 > 0:
        tv, tq = synth.truth_state_at(truth, k0 - 1)
    vec, quat, cov = synth.initial_ensemble(N, tv, tq, seed=seed)
    st = synth.make_streams(truth, N, k0, T, with_legodo=with_legodo, with_pose=with_pose, seed=seed)
    return dict(truth=truth, vec=vec, quat=quat, cov=cov, st=st)


def oracle_streams(st):
    return [dict(idx=synth.LEGODO_IDX, z=st["legodo"], R=st["R_legodo"]),
            dict(idx=synth.POSE_IDX, z=st["pose_z"], R=st["R_pose"], quat=st["pose_q"])]


def gpu_streams(st):
    from pronto_b200 import MeasStream

    return [MeasStream(synth.LEGODO_IDX, st["legodo"], st["R_legodo"]),
            MeasStream(synth.POSE_IDX, st["pose_z"], st["R_pose"], quat=st["pose_q"])]
