"""The oracle's restatement of rbis.cpp:12-227 against the REFERENCE's own rbis.cpp, compiled unmodified from
/root/reference against the stand-in headers of oracle/ref_shim/ (oracle/_ref/librbis_ref.so, `make -C oracle ref`).
This pins every formula of the hot path (Ac blocks, Wc/Qd, Ad P Ad^T, K / dcov / log-likelihood, residual rules,
rbisApplyDelta) to the reference's source text.  The eigen_utils semantics (RigidBodyState algebra, g_vec, chi
tolerance) are the same recalled ones on both sides and stay unpinned.  CPU only."""
import numpy as np
import pytest

from common import random_ensemble
from oracle import oracle_api

pytestmark = pytest.mark.skipif(oracle_api.build_ref() is None, reason="reference tree not mounted and oracle/_ref not prebuilt")

TOL = 1e-12


def _both(fn):
    a = fn()
    with oracle_api.reference():
        b = fn()
    return a, b


def _rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def test_linearization_and_process_step_match_reference_source():
    vec, quat, cov = random_ensemble(12, seed=5)
    rng = np.random.default_rng(6)
    for n in range(12):
        v, q, P = vec[:, n], quat[:, n], cov[:, n].reshape(21, 21).T
        gyro, acc = rng.normal(size=3) * 0.3, rng.normal(size=3) + np.array([0, 0, 9.8])
        A0, A1 = _both(lambda: oracle_api.linearization(v, q))
        assert np.array_equal(A0 != 0, A1 != 0) and _rel(A0, A1) < TOL                      # rbis.cpp:12-35
        s0, s1 = _both(lambda: oracle_api.ins_update_state(gyro, acc, 1e-3, v, q))
        assert _rel(s0[0], s1[0]) < TOL and _rel(s0[1], s1[1]) < TOL                          # rbis.cpp:37-75
        P0, P1 = _both(lambda: oracle_api.ins_update_covariance(7.6e-5, 1e-2, 3e-10, 1e-6, v, q, P, 1e-3))
        assert _rel(P0, P1) < TOL                                                             # rbis.cpp:77-122
        # the overwrites of rbis.cpp:120-121
        assert np.allclose(P1[12:15, 12:15], 1e-2 * np.eye(3)) and np.allclose(P1[0:3, 0:3], 7.6e-5 * np.eye(3))


@pytest.mark.parametrize("idx,orient", [([3, 4, 5], False), ([9, 10, 11, 6, 7, 8], True), ([8], True), ([17, 8], True),
                                        ([8, 9, 10, 11], False), ([0, 1, 2, 3, 4, 5, 9, 10, 11], False), ([17], False)])
def test_measurement_updates_match_reference_source(idx, orient):
    vec, quat, cov = random_ensemble(6, seed=11)
    rng = np.random.default_rng(12)
    m = len(idx)
    for n in range(6):
        v, q, P = vec[:, n], quat[:, n], cov[:, n].reshape(21, 21).T
        z = rng.normal(size=m)
        A = rng.normal(size=(m, m))
        R = A @ A.T * 1e-3 + np.eye(m) * 1e-3                                                # full (non-diagonal) R
        dq = rng.normal(size=4) * 0.05 + np.array([1, 0, 0, 0]); dq /= np.linalg.norm(dq)
        mq = None
        if orient:
            mq = np.array([q[0] * dq[0] - q[1:] @ dq[1:], *(q[0] * dq[1:] + dq[0] * q[1:] + np.cross(q[1:], dq[1:]))])
        r0, r1 = _both(lambda: oracle_api.measurement_update(z, R, idx, v, q, P, meas_quat=mq))  # rbis.cpp:124-227
        assert _rel(r0[0], r1[0]) < TOL and _rel(r0[1], r1[1]) < TOL and _rel(r0[2], r1[2]) < 1e-11
        assert abs(r0[3] - r1[3]) < 1e-9 * max(1.0, abs(r1[3]))


def test_rolled_sequence_and_constants_match_reference_source():
    """400 alternating updates: the two implementations stay together; flipping the recalled constants moves both."""
    rng = np.random.default_rng(21)
    for consts in (dict(), dict(g_val=9.81, chi_tol=1e-3, ctor_folds_chi=False)):
        def run():
            oracle_api.set_constants(**consts)
            try:
                vec, quat, cov = random_ensemble(1, seed=3)
                v, q, P = vec[:, 0].copy(), quat[:, 0].copy(), cov[:, 0].reshape(21, 21).T.copy()
                r = np.random.default_rng(4)
                for k in range(400):
                    P = oracle_api.ins_update_covariance(7.6e-5, 1e-2, 3e-10, 1e-6, v, q, P, 1e-3)
                    v, q = oracle_api.ins_update_state(r.normal(size=3) * (1e-4 if k % 7 else 0.3), r.normal(size=3) + [0, 0, 9.8], 1e-3, v, q)
                    if k % 2 == 0:
                        v, q, P, _ = oracle_api.measurement_update(v[3:6] + r.normal(size=3) * 0.05, np.eye(3) * 0.01, [3, 4, 5], v, q, P)
                    if k % 50 == 0:
                        v, q, P, _ = oracle_api.measurement_update(np.r_[v[9:12], 0, 0, 0], np.eye(6) * 1e-3, [9, 10, 11, 6, 7, 8], v, q, P, meas_quat=q)
                return v, q, P
            finally:
                oracle_api.set_constants()
        a, b = _both(run)
        assert _rel(a[0], b[0]) < 1e-10 and _rel(a[1], b[1]) < 1e-10 and _rel(a[2], b[2]) < 1e-9
    del rng


def test_update_objects_and_history_driver_match_reference_source():
    """The reference's own RBISIMUProcessStep / RBISIndexed[PlusOrientation]Measurement::updateFilter, updateHistory and
    MavStateEstimator::addUpdate (compiled from /root/reference) against the oracle's restatement: in-order replay with a
    per-event trace, then pose fixes delivered 50 steps late (out-of-order insert + replay, mav_state_est.cpp:35-70) and a
    short history span (truncation, :74-77)."""
    from common import nominal_q, oracle_streams, scenario

    N, T = 3, 260
    sc = scenario(N, T)
    st = sc["st"]
    args = (sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st))
    a, b = _both(lambda: oracle_api.run_ensemble(*args, st["events"], trace=True))
    for k in ("vec", "quat", "cov", "loglik", "trace_vec", "trace_quat", "trace_cov", "trace_loglik"):
        assert _rel(a[k], b[k]) < (1e-9 if "cov" in k else 1e-10), k
    ev = st["events"]
    pose = [e for e in ev if e[0] == 1 and e[1] == 1]
    arrivals, pending = [], list(pose)
    for e in ev:
        if e[0] == 1 and e[1] == 1:
            continue
        arrivals.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + 50_000:
            arrivals.append(pending.pop(0))
    arrivals += pending
    for span in (10_000_000, 80_000):
        a, b = _both(lambda: oracle_api.run_ensemble(*args, arrivals, history_span=span))
        for k in ("vec", "quat", "cov", "loglik"):
            assert _rel(a[k], b[k]) < (1e-9 if k == "cov" else 1e-10), (span, k)
    # equal-utime updates keep their arrival order in both (hinted multimap insert, update_history.cpp:26)
    swapped = list(ev)
    i = next(k for k, e in enumerate(swapped) if e[0] == 1 and e[1] == 0)
    j = next(k for k, e in enumerate(swapped) if e[0] == 1 and e[1] == 1 and e[3] == swapped[i][3])
    swapped[i], swapped[j] = swapped[j], swapped[i]
    a, b = _both(lambda: oracle_api.run_ensemble(*args, swapped))
    assert _rel(a["vec"], b["vec"]) < 1e-10 and _rel(a["cov"], b["cov"]) < 1e-9
