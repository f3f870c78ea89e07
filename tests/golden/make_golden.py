"""Generates tests/golden/rbis_golden.npz with the numpy restatement (oracle/rbis_numpy.py).

The reference cannot be built or imported in this image (Eigen, eigen_utils, LCM, libbot are all
absent) and ships no golden vectors for this path, so these fixtures pin the *restated* algorithm:
they are produced by the numpy restatement, which is written from rbis.cpp independently of the C++
oracle, and both the C++ oracle (CPU tests) and the CUDA path (gpu tests) must reproduce them.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import rbis_numpy as rn  # noqa: E402
from pronto_b200 import synth  # noqa: E402

from common import nominal_q, scenario  # noqa: E402


def run_numpy(sc, N):
    st = sc["st"]
    qg, qa, qgb, qab = nominal_q()
    E = len(st["events"])
    out_vec, out_quat, out_cov, out_ll = np.zeros((21, N)), np.zeros((4, N)), np.zeros((441, N)), np.zeros(N)
    marks = sorted(set([0, 1, 2, 3, 4, E // 2, E - 1]))
    tr_vec, tr_quat, tr_cov = np.zeros((len(marks), 21, N)), np.zeros((len(marks), 4, N)), np.zeros((len(marks), 441, N))
    for n in range(N):
        s = rn.State(sc["vec"][:, n], sc["quat"][:, n])
        P = sc["cov"][:, n].reshape(21, 21).T.copy()
        ll = 0.0
        for e, (kind, stream, row, _, dt) in enumerate(st["events"]):
            if kind == 0:
                prior = s.copy()
                rn.ins_update_state(st["imu"][row, 0:3, n], st["imu"][row, 3:6, n], dt, s)
                P = rn.ins_update_covariance(qg, qa, qgb, qab, prior, P, dt)
            elif stream == 0:
                s, P, l = rn.measurement(st["legodo"][row, :, n], st["R_legodo"], synth.LEGODO_IDX, s, P)
                ll += l
            else:
                s, P, l = rn.measurement(st["pose_z"][row, :, n], st["R_pose"], synth.POSE_IDX, s, P, st["pose_q"][row, :, n])
                ll += l
            if e in marks:
                m = marks.index(e)
                tr_vec[m, :, n], tr_quat[m, :, n], tr_cov[m, :, n] = s.vec, s.quat, P.T.reshape(-1)
        out_vec[:, n], out_quat[:, n], out_cov[:, n], out_ll[n] = s.vec, s.quat, P.T.reshape(-1), ll
    return dict(vec=out_vec, quat=out_quat, cov=out_cov, loglik=out_ll, marks=np.array(marks), tr_vec=tr_vec,
                tr_quat=tr_quat, tr_cov=tr_cov)


def main():
    N, T = 4, 400
    blob = {}
    for name, tumbling in (("walk", False), ("tumble", True)):
        sc = scenario(N, T, tumbling=tumbling)
        st = sc["st"]
        out = run_numpy(sc, N)
        blob.update({f"{name}_in_vec": sc["vec"], f"{name}_in_quat": sc["quat"], f"{name}_in_cov": sc["cov"],
                     f"{name}_imu": st["imu"], f"{name}_legodo": st["legodo"], f"{name}_pose_z": st["pose_z"],
                     f"{name}_pose_q": st["pose_q"]})
        blob.update({f"{name}_{k}": v for k, v in out.items()})
    # single-op vectors on a random dense-covariance state
    rng = np.random.default_rng(20261018)
    vec = rng.normal(size=21) * 0.3
    vec[6:9] = 0
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    A = rng.normal(size=(21, 21))
    P = A @ A.T * 0.01 / 21 + np.eye(21) * 0.01
    s = rn.State(vec, q)
    gyro, accel = rng.normal(size=3) * 0.5, rng.normal(size=3) + np.array([0, 0, 9.8])
    s2 = s.copy()
    rn.ins_update_state(gyro, accel, 1e-3, s2)
    P2 = rn.ins_update_covariance(*nominal_q(), s, P, 1e-3)
    blob.update(op_vec=vec, op_quat=q, op_cov=P, op_gyro=gyro, op_accel=accel, op_ins_vec=s2.vec, op_ins_quat=s2.quat,
                op_ins_cov=P2)
    cases = [([3, 4, 5], None), ([9, 10, 11, 6, 7, 8], "q"), ([17], None), ([8, 9, 10, 11], None), ([17, 8], "q"),
             ([3, 4, 5, 0, 1, 2], None), ([9, 10, 11, 3, 4, 5, 6, 7, 8], "q")]
    for c, (idx, oq) in enumerate(cases):
        m = len(idx)
        z = rng.normal(size=m) * 0.3
        B = rng.normal(size=(m, m))
        R = B @ B.T * 0.01 + np.eye(m) * 0.01  # dense SPD: exercises the general path
        mq = rn.qmul(q, rn.qexp(rng.normal(size=3) * 0.05)) if oq else None
        ps, pc, ll = rn.measurement(z, R, idx, s, P, mq)
        blob.update({f"m{c}_idx": np.array(idx), f"m{c}_z": z, f"m{c}_R": R, f"m{c}_mq": mq if oq else np.zeros(0),
                     f"m{c}_vec": ps.vec, f"m{c}_quat": ps.quat, f"m{c}_cov": pc, f"m{c}_ll": np.array(ll)})
    blob["n_meas_cases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rbis_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
