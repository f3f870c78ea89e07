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


# ------------------------------------------------------------------------------------------------
# fixtures produced by the REFERENCE's own code (tests/golden/make_reference_golden.py, oracle/_ref)
# ------------------------------------------------------------------------------------------------
REF_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rbis_reference_golden.npz")


@pytest.fixture(scope="module")
def ref_golden():
    return np.load(REF_GOLDEN)


def _delayed(ev, latency_us=50_000):
    pose = [e for e in ev if e[0] == 1 and e[1] == 1]
    arrivals, pending = [], list(pose)
    for e in ev:
        if e[0] == 1 and e[1] == 1:
            continue
        arrivals.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + latency_us:
            arrivals.append(pending.pop(0))
    return arrivals + pending


@pytest.mark.parametrize("name,tumbling", [("walk", False), ("tumble", True)])
def test_oracle_matches_reference_golden_trajectories(oracle, ref_golden, name, tumbling):
    from common import oracle_streams

    sc = scenario(4, 400, tumbling=tumbling)
    st = sc["st"]
    args = (sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st))
    out = oracle.run_ensemble(*args, st["events"], trace=True)
    g = ref_golden
    e = max_errors(out["vec"], out["quat"], out["cov"], g[f"{name}_vec"], g[f"{name}_quat"], g[f"{name}_cov"])
    assert e["vec"] < 1e-11 and e["quat"] < 1e-11 and e["cov"] < 1e-9, e
    assert np.max(np.abs(out["loglik"] - g[f"{name}_loglik"]) / np.abs(g[f"{name}_loglik"])) < 1e-11
    for m, ev in enumerate(g[f"{name}_marks"]):
        e = max_errors(out["trace_vec"][ev], out["trace_quat"][ev], out["trace_cov"][ev], g[f"{name}_tr_vec"][m],
                       g[f"{name}_tr_quat"][m], g[f"{name}_tr_cov"][m])
        assert e["vec"] < 1e-11 and e["quat"] < 1e-11 and e["cov"] < 1e-9, (ev, e)
    late = oracle.run_ensemble(*args, _delayed(st["events"]))
    e = max_errors(late["vec"], late["quat"], late["cov"], g[f"{name}_late_vec"], g[f"{name}_late_quat"], g[f"{name}_late_cov"])
    assert e["vec"] < 1e-11 and e["quat"] < 1e-11 and e["cov"] < 1e-9, e


def test_oracle_matches_reference_golden_single_updates(oracle, ref_golden):
    g = ref_golden
    vec, q, P = g["op_vec"], g["op_quat"], g["op_cov"]
    assert np.max(np.abs(oracle.linearization(vec, q) - g["op_Ac"])) < 1e-15
    v2, q2 = oracle.ins_update_state(g["op_gyro"], g["op_accel"], 1e-3, vec, q)
    assert np.max(np.abs(v2 - g["op_ins_vec"])) < 1e-14 and np.max(np.abs(q2 - g["op_ins_quat"])) < 1e-14
    assert np.max(np.abs(oracle.ins_update_covariance(*nominal_q(), vec, q, P, 1e-3) - g["op_ins_cov"])) < 1e-15
    for c in range(int(g["n_meas_cases"])):
        mq = g[f"m{c}_mq"]
        pv, pq, pc, ll = oracle.measurement_update(g[f"m{c}_z"], g[f"m{c}_R"], list(g[f"m{c}_idx"]), vec, q, P, mq if mq.size else None)
        assert np.max(np.abs(pv - g[f"m{c}_vec"])) < 1e-12 and np.max(np.abs(pq - g[f"m{c}_quat"])) < 1e-12, c
        assert np.max(np.abs(pc - g[f"m{c}_cov"])) < 1e-13 and abs(ll - float(g[f"m{c}_ll"])) < 1e-10 * max(1.0, abs(ll)), c
