"""The C++ oracle must reproduce the committed golden vectors (tests/golden/rbis_golden.npz, made by the
independent numpy restatement with tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from pronto_b200 import synth
from pronto_b200.parity import max_errors

from common import nominal_q, scenario

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rbis_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name,tumbling", [("walk", False), ("tumble", True)])
def test_synth_inputs_are_deterministic(golden, name, tumbling):
    sc = scenario(4, 400, tumbling=tumbling)
    assert np.array_equal(sc["vec"], golden[f"{name}_in_vec"])
    assert np.array_equal(sc["quat"], golden[f"{name}_in_quat"])
    assert np.array_equal(sc["st"]["imu"], golden[f"{name}_imu"])
    assert np.array_equal(sc["st"]["legodo"], golden[f"{name}_legodo"])
    assert np.array_equal(sc["st"]["pose_z"], golden[f"{name}_pose_z"])
    assert np.array_equal(sc["st"]["pose_q"], golden[f"{name}_pose_q"])


@pytest.mark.parametrize("name", ["walk", "tumble"])
def test_oracle_trajectory_matches_golden(oracle, golden, name):
    N = 4
    streams = [dict(idx=synth.LEGODO_IDX, z=golden[f"{name}_legodo"], R=np.eye(3) * synth.NOMINAL["r_vxyz"] ** 2),
               dict(idx=synth.POSE_IDX, z=golden[f"{name}_pose_z"], quat=golden[f"{name}_pose_q"],
                    R=np.diag([synth.NOMINAL["r_xyz"] ** 2] * 3 + [synth.NOMINAL["r_chi"] ** 2] * 3))]
    events = scenario(N, 400)["st"]["events"]
    out = oracle.run_ensemble(golden[f"{name}_in_vec"], golden[f"{name}_in_quat"], golden[f"{name}_in_cov"], None, 0,
                              nominal_q(), golden[f"{name}_imu"], streams, events, trace=True)
    e = max_errors(out["vec"], out["quat"], out["cov"], golden[f"{name}_vec"], golden[f"{name}_quat"], golden[f"{name}_cov"])
    assert e["vec"] < 1e-12 and e["quat"] < 1e-12 and e["cov"] < 1e-11, e
    assert np.max(np.abs(out["loglik"] - golden[f"{name}_loglik"]) / np.abs(golden[f"{name}_loglik"])) < 1e-11
    for m, ev in enumerate(golden[f"{name}_marks"]):
        e = max_errors(out["trace_vec"][ev], out["trace_quat"][ev], out["trace_cov"][ev], golden[f"{name}_tr_vec"][m],
                       golden[f"{name}_tr_quat"][m], golden[f"{name}_tr_cov"][m])
        assert e["vec"] < 1e-12 and e["quat"] < 1e-12 and e["cov"] < 1e-11, (ev, e)


def test_oracle_single_ops_match_golden(oracle, golden):
    vec, q, P = golden["op_vec"], golden["op_quat"], golden["op_cov"]
    v2, q2 = oracle.ins_update_state(golden["op_gyro"], golden["op_accel"], 1e-3, vec, q)
    assert np.max(np.abs(v2 - golden["op_ins_vec"])) < 1e-15 and np.max(np.abs(q2 - golden["op_ins_quat"])) < 1e-15
    P2 = oracle.ins_update_covariance(*nominal_q(), vec, q, P, 1e-3)
    assert np.max(np.abs(P2 - golden["op_ins_cov"])) < 1e-16
    for c in range(int(golden["n_meas_cases"])):
        mq = golden[f"m{c}_mq"]
        pv, pq, pc, ll = oracle.measurement_update(golden[f"m{c}_z"], golden[f"m{c}_R"], list(golden[f"m{c}_idx"]), vec,
                                                   q, P, mq if mq.size else None)
        assert np.max(np.abs(pv - golden[f"m{c}_vec"])) < 1e-13, c
        assert np.max(np.abs(pq - golden[f"m{c}_quat"])) < 1e-13, c
        assert np.max(np.abs(pc - golden[f"m{c}_cov"])) < 1e-14, c
        assert abs(ll - float(golden[f"m{c}_ll"])) < 1e-11 * max(1.0, abs(ll)), c
