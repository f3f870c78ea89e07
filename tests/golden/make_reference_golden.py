"""Generates tests/golden/rbis_reference_golden.npz by running the REFERENCE's own code: oracle/_ref/librbis_ref.so is
/root/reference/state-estimator/src/mav_state_est/{rbis,rbis_update_interface,update_history,mav_state_est}.cpp compiled
unmodified against the stand-in dependency headers of oracle/ref_shim/ (`make -C oracle ref`).  It can only be run
where /root/reference is mounted; the fixtures travel.  Both the CPU oracle (tests/test_golden_cpu.py) and the CUDA
path (tests/test_gpu_parity.py) must reproduce them.

Run from the repo root:  python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_api  # noqa: E402

from common import nominal_q, oracle_streams, scenario  # noqa: E402


def delayed(ev, latency_us=50_000):
    pose = [e for e in ev if e[0] == 1 and e[1] == 1]
    arrivals, pending = [], list(pose)
    for e in ev:
        if e[0] == 1 and e[1] == 1:
            continue
        arrivals.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + latency_us:
            arrivals.append(pending.pop(0))
    return arrivals + pending


def main():
    assert oracle_api.build_ref(force=True), "needs /root/reference"
    N, T = 4, 400
    blob = {}
    with oracle_api.reference():
        for name, tumbling in (("walk", False), ("tumble", True)):
            sc = scenario(N, T, tumbling=tumbling)
            st = sc["st"]
            args = (sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st))
            out = oracle_api.run_ensemble(*args, st["events"], trace=True)
            E = len(st["events"])
            marks = sorted(set([0, 1, 2, 3, 4, E // 2, E - 1]))
            blob.update({f"{name}_vec": out["vec"], f"{name}_quat": out["quat"], f"{name}_cov": out["cov"], f"{name}_loglik": out["loglik"],
                         f"{name}_marks": np.array(marks), f"{name}_tr_vec": out["trace_vec"][marks], f"{name}_tr_quat": out["trace_quat"][marks],
                         f"{name}_tr_cov": out["trace_cov"][marks], f"{name}_tr_loglik": out["trace_loglik"][marks]})
            late = oracle_api.run_ensemble(*args, delayed(st["events"]))
            blob.update({f"{name}_late_vec": late["vec"], f"{name}_late_quat": late["quat"], f"{name}_late_cov": late["cov"],
                         f"{name}_late_loglik": late["loglik"]})
        # single updates on a random dense-covariance state (same draws as make_golden.py)
        rng = np.random.default_rng(20261018)
        vec = rng.normal(size=21) * 0.3
        vec[6:9] = 0
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        A = rng.normal(size=(21, 21))
        P = A @ A.T * 0.01 / 21 + np.eye(21) * 0.01
        gyro, accel = rng.normal(size=3) * 0.5, rng.normal(size=3) + np.array([0, 0, 9.8])
        v2, q2 = oracle_api.ins_update_state(gyro, accel, 1e-3, vec, q)
        P2 = oracle_api.ins_update_covariance(*nominal_q(), vec, q, P, 1e-3)
        blob.update(op_vec=vec, op_quat=q, op_cov=P, op_gyro=gyro, op_accel=accel, op_ins_vec=v2, op_ins_quat=q2, op_ins_cov=P2,
                    op_Ac=oracle_api.linearization(vec, q))
        cases = [([3, 4, 5], False), ([9, 10, 11, 6, 7, 8], True), ([17], False), ([8, 9, 10, 11], False), ([17, 8], True),
                 ([3, 4, 5, 0, 1, 2], False), ([9, 10, 11, 3, 4, 5, 6, 7, 8], True)]
        for c, (idx, orient) in enumerate(cases):
            m = len(idx)
            z = rng.normal(size=m) * 0.3
            B = rng.normal(size=(m, m))
            R = B @ B.T * 0.01 + np.eye(m) * 0.01
            d = rng.normal(size=3) * 0.05
            a = np.linalg.norm(d)
            dq = np.r_[np.cos(a / 2), np.sin(a / 2) * d / a]
            mq = np.r_[q[0] * dq[0] - q[1:] @ dq[1:], q[0] * dq[1:] + dq[0] * q[1:] + np.cross(q[1:], dq[1:])] if orient else None
            pv, pq, pc, ll = oracle_api.measurement_update(z, R, idx, vec, q, P, meas_quat=mq)
            blob.update({f"m{c}_idx": np.array(idx), f"m{c}_z": z, f"m{c}_R": R, f"m{c}_mq": mq if orient else np.zeros(0),
                         f"m{c}_vec": pv, f"m{c}_quat": pq, f"m{c}_cov": pc, f"m{c}_ll": np.array(ll)})
        blob["n_meas_cases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rbis_reference_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
