"""rbis_numpy.py -- second, independent CPU restatement of the RBIS EKF hot path in numpy float64.

CPU ORACLE, test infrastructure only (PARITY UNPINNED: see oracle/rbis_oracle.hpp).  It is written
from the formulas in /root/reference/state-estimator/src/mav_state_est/rbis.cpp, not from
rbis_oracle.cpp, so the two can be cross-checked (known-answer test 9).  Only tests/
may import it.  Single filter, pure-Python loops: small cases only.
"""
import numpy as np

G_VAL = 9.8      # eigen_utils g_val   [RECALLED, SURVEY.md 8c]
CHI_TOL = 1e-6   # chiToQuat tolerance [RECALLED]
W_, V_, CHI_, P_, A_, BG_, BA_ = 0, 3, 6, 9, 12, 15, 18
NS = 21


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
    ])


def qinv(q):
    return np.array([q[0], -q[1], -q[2], -q[3]]) / np.dot(q, q)


def qrot(q, v):
    u = q[1:]
    uv = 2.0 * np.cross(u, v)
    return v + q[0] * uv + np.cross(u, uv)


def qmat(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def qexp(chi):
    n = np.linalg.norm(chi)
    return np.concatenate(([np.cos(0.5 * n)], np.sin(0.5 * n) * chi / n))


def qlog(q):
    """AngleAxis(q) as Eigen >= 3.3, times angle (angle in [0, pi])."""
    n = np.linalg.norm(q[1:])
    if n == 0.0:
        return np.zeros(3)
    ang = 2.0 * np.arctan2(n, abs(q[0]))
    if ang >= np.pi:
        ang -= 2 * np.pi
    return q[1:] / (-n if q[0] < 0 else n) * ang


def subtract_quats(q1, q2):
    return qlog(qmul(qinv(q2), q1))


class State:
    def __init__(self, vec=None, quat=None):
        self.vec = np.zeros(NS) if vec is None else np.array(vec, dtype=np.float64)
        self.quat = np.array([1.0, 0, 0, 0]) if quat is None else np.array(quat, dtype=np.float64)

    def copy(self):
        return State(self.vec.copy(), self.quat.copy())

    def chi_to_quat(self):
        n = np.linalg.norm(self.vec[CHI_:CHI_ + 3])
        if n > CHI_TOL:
            self.quat = qmul(self.quat, qexp(self.vec[CHI_:CHI_ + 3]))
            self.vec[CHI_:CHI_ + 3] = 0.0

    def add_state(self, d):
        self.vec = self.vec + d.vec
        self.chi_to_quat()
        self.quat = qmul(self.quat, d.quat)


def linearization(s):
    """rbis.cpp:12-35"""
    Ac = np.zeros((NS, NS))
    w, v = s.vec[W_:W_ + 3], s.vec[V_:V_ + 3]
    R = qmat(s.quat)
    gb = qrot(qinv(s.quat), np.array([0, 0, -G_VAL]))
    Ac[V_:V_ + 3, V_:V_ + 3] = -skew(w)
    Ac[V_:V_ + 3, CHI_:CHI_ + 3] = skew(gb)
    Ac[CHI_:CHI_ + 3, CHI_:CHI_ + 3] = -skew(w)
    Ac[P_:P_ + 3, V_:V_ + 3] = R
    Ac[P_:P_ + 3, CHI_:CHI_ + 3] = -R @ skew(v)
    Ac[V_:V_ + 3, BG_:BG_ + 3] = -skew(v)
    Ac[V_:V_ + 3, BA_:BA_ + 3] = -np.eye(3)
    Ac[CHI_:CHI_ + 3, BG_:BG_ + 3] = -np.eye(3)
    return Ac


def ins_update_state(gyro, accel, dt, s):
    """rbis.cpp:37-75 (in place)"""
    s.vec[W_:W_ + 3] = gyro - s.vec[BG_:BG_ + 3]
    s.vec[A_:A_ + 3] = accel - s.vec[BA_:BA_ + 3]
    w, v = s.vec[W_:W_ + 3], s.vec[V_:V_ + 3]
    d = State()
    d.vec[V_:V_ + 3] = -np.cross(w, v) + qrot(qinv(s.quat), np.array([0, 0, -G_VAL])) + s.vec[A_:A_ + 3]
    d.vec[CHI_:CHI_ + 3] = w
    d.vec[P_:P_ + 3] = qrot(s.quat, v)
    d.vec *= dt
    d.chi_to_quat()
    s.add_state(d)


def ins_update_covariance(q_gyro, q_accel, q_gyro_bias, q_accel_bias, s, cov, dt):
    """rbis.cpp:77-122; returns the new covariance"""
    Ac = linearization(s)
    Wc = np.zeros((NS, 12))
    Wc[V_:V_ + 3, 0:3] = skew(s.vec[V_:V_ + 3])
    Wc[V_:V_ + 3, 3:6] = np.eye(3)
    Wc[CHI_:CHI_ + 3, 0:3] = np.eye(3)
    Wc[BG_:BG_ + 3, 6:9] = np.eye(3)
    Wc[BA_:BA_ + 3, 9:12] = np.eye(3)
    Qc = np.diag([q_gyro] * 3 + [q_accel] * 3 + [q_gyro_bias] * 3 + [q_accel_bias] * 3)
    Ad = np.eye(NS) + Ac * dt
    Qd = Wc @ Qc @ Wc.T * dt
    out = Ad @ cov @ Ad.T + Qd
    out[A_:A_ + 3, A_:A_ + 3] = q_accel * np.eye(3)
    out[W_:W_ + 3, W_:W_ + 3] = q_gyro * np.eye(3)
    return out


def measurement(z, R, idx, s, cov, meas_quat=None):
    """rbis.cpp:124-217 + rbisApplyDelta :219-227.  Returns (posterior state, posterior cov, loglik term)."""
    m = len(idx)
    R = np.asarray(R, dtype=np.float64).reshape(m, m)
    C = np.zeros((m, NS))
    r = np.zeros(m)
    dq = subtract_quats(meas_quat, s.quat) if meas_quat is not None else None
    for i, j in enumerate(idx):
        if dq is not None and CHI_ <= j <= CHI_ + 2:
            r[i] = dq[j - CHI_]
        else:
            r[i] = z[i] - s.vec[j]
        C[i, j] = 1.0
    S = R + C @ cov @ C.T
    K = np.linalg.solve(S, C @ cov).T
    dcov = K @ C @ cov
    ll = -np.log(np.linalg.det(S)) - r @ np.linalg.solve(S, r)
    d = State(K @ r)
    d.chi_to_quat()
    post = s.copy()
    post.add_state(d)
    return post, cov - dcov, ll


def state_error(est, truth):
    """SE/noise_id/noise_id.cpp:37-38"""
    e = est.vec - truth.vec
    e[CHI_:CHI_ + 3] = qlog(qmul(qinv(truth.quat), est.quat))
    return e


def noise_id_neg_loglik(states, covs, dt, q_gyro, q_accel, n_window, active=(3, 4, 5, 6, 7, 8, 9, 10, 11)):
    """state-estimator/src/noise_id/noise_id.cpp:9-65: states = list of State (truth history), covs = list of 21x21.
    Returns (negative log-likelihood, list of per-window error vectors)."""
    act = list(active)
    it, nll, errs = 0, 0.0, []
    while True:
        rolled = states[it].copy()
        start_cov = covs[it].copy()
        rolled_cov = start_cov.copy()
        for _ in range(n_window):
            rolled_cov = ins_update_covariance(q_gyro, q_accel, 0.0, 0.0, rolled, rolled_cov, dt)
            start_cov = ins_update_covariance(0.0, 0.0, 0.0, 0.0, rolled, start_cov, dt)
            ins_update_state(states[it].vec[W_:W_ + 3].copy(), states[it].vec[A_:A_ + 3].copy(), dt, rolled)
            it += 1
            if it == len(states):
                return nll, errs
        e = state_error(rolled, states[it])
        Cm = (rolled_cov - start_cov)[np.ix_(act, act)]
        ea = e[act]
        nll -= -np.log(np.linalg.det(Cm)) - ea @ np.linalg.solve(Cm, ea)  # loglike_normalized [RECALLED], see rbis_oracle.hpp
        errs.append(e)
