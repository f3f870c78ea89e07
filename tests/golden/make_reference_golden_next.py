"""Generates tests/golden/rbis_reference_golden_next.npz for the "next" rows by running the REFERENCE's own
code (oracle/_ref/librbis_ref.so, see make_reference_golden.py): MavStateEstimator::EKFSmoothBackwardsPass over a replayed
history (MSE/mav_state_est.cpp:98-189, MSE/rbis.cpp:234-266) and the accelerometer notch cascade
(estimate_tools/src/estimate_tools/iir_notch.cpp).  Inputs are the seeded scenarios of tests/common.py, so only outputs
are stored.  Can only be run where /root/reference is mounted; the fixtures travel.

Run from the repo root:  python tests/golden/make_reference_golden_next.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_api  # noqa: E402

from common import nominal_q, oracle_streams, scenario  # noqa: E402

SMOOTH_N, SMOOTH_T = 3, 61


def smoothing_events(trailing):
    sc = scenario(SMOOTH_N, SMOOTH_T, tumbling=True)
    ev = list(sc["st"]["events"])
    if trailing:
        while ev[-1][0] == 0:
            ev.pop()
    return sc, ev


def notch_signal(n=3000, seed=11):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 1000.0
    return rng.normal(size=n) * 0.3 + 2.0 * np.sin(2 * np.pi * 85 * t) + 9.8


def main():
    assert oracle_api.build_ref(force=True), "needs /root/reference"
    blob = {}
    with oracle_api.reference():
        oracle_api.set_constants()
        for name, trailing in (("plain", False), ("trailing", True)):
            sc, ev = smoothing_events(trailing)
            st = sc["st"]
            out = oracle_api.smooth_ensemble(sc["vec"], sc["quat"], sc["cov"], 0, nominal_q(), st["imu"], oracle_streams(st), ev, 1e-3)
            blob.update({f"smooth_{name}_vec": out["post_vec"], f"smooth_{name}_quat": out["post_quat"], f"smooth_{name}_cov": out["post_cov"]})
        y, state, coeffs = oracle_api.notch_cascade(notch_signal(), 85.0, 1000.0, 3)
        blob.update(notch_y=y, notch_state=state, notch_coeffs=coeffs)
    path = os.path.join(ROOT, "tests", "golden", "rbis_reference_golden_next.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
