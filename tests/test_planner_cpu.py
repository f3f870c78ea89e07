"""The delayed-measurement planner (csrc/rbis_planner.cpp) must order work exactly like the
reference's history (MSE/update_history.cpp:16-54, MSE/mav_state_est.cpp:28-80).  Its op programs are
executed here with the oracle's single-update functions and compared, bit for bit, with the oracle's
MavStateEstimator fed the same arrivals.  CPU only."""
import numpy as np
import pytest

from pronto_b200 import capi, synth
from pronto_b200.schedule import Planner, program_from_arrivals

from common import nominal_q, oracle_streams, scenario


def delay_pose(events, latency_steps):
    pose = [e for e in events if e[0] == 1 and e[1] == 1]
    out, pending = [], list(pose)
    for e in events:
        if e[0] == 1 and e[1] == 1:
            continue
        out.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + latency_steps * 1000:
            out.append(pending.pop(0))
    return out + pending


def execute_program(oracle, ops, sc, n):
    """Run an op program for filter n with the oracle's per-update functions."""
    st = sc["st"]
    vec, quat = sc["vec"][:, n].copy(), sc["quat"][:, n].copy()
    P = sc["cov"][:, n].reshape(21, 21).T.copy()
    ll = 0.0
    slots = {}
    q = nominal_q()
    n_applied = 0
    for op in ops:
        kind, stream, row, dt = int(op["kind"]), int(op["stream"]), int(op["row"]), float(op["dt"])
        if kind == capi.OP_IMU:
            P = oracle.ins_update_covariance(*q, vec, quat, P, dt)  # linearised at the prior state
            vec, quat = oracle.ins_update_state(st["imu"][row, 0:3, n], st["imu"][row, 3:6, n], dt, vec, quat)
            n_applied += 1
        elif kind == capi.OP_MEAS:
            if stream == 0:
                vec, quat, P, l = oracle.measurement_update(st["legodo"][row, :, n], st["R_legodo"], synth.LEGODO_IDX, vec, quat, P)
            else:
                vec, quat, P, l = oracle.measurement_update(st["pose_z"][row, :, n], st["R_pose"], synth.POSE_IDX, vec, quat, P,
                                                            st["pose_q"][row, :, n])
            ll += l
            n_applied += 1
        elif kind == capi.OP_SNAPSHOT:
            slots[row] = (vec.copy(), quat.copy(), P.copy(), ll)
        elif kind == capi.OP_RESTORE:
            vec, quat, P, ll = (x.copy() if hasattr(x, "copy") else x for x in slots[row])
        else:
            raise AssertionError(kind)
    return vec, quat, P, ll, n_applied


def oracle_history(oracle, sc, arrivals, n, **kw):
    st = sc["st"]
    sl = slice(n, n + 1)
    streams = [dict(s, z=s["z"][:, :, sl], **({"quat": s["quat"][:, :, sl]} if "quat" in s else {})) for s in oracle_streams(st)]
    return oracle.run_ensemble(sc["vec"][:, sl], sc["quat"][:, sl], sc["cov"][:, sl], None, 0, nominal_q(),
                               st["imu"][:, :, sl], streams, arrivals, **kw)


def check(oracle, sc, arrivals, ops, n=0, **kw):
    ref = oracle_history(oracle, sc, arrivals, n, **kw)
    vec, quat, P, ll, n_applied = execute_program(oracle, ops, sc, n)
    assert np.array_equal(vec, ref["vec"][:, 0]) and np.array_equal(quat, ref["quat"][:, 0])
    assert np.array_equal(P.T.reshape(-1), ref["cov"][:, 0]) and ll == ref["loglik"][0]
    return ref, n_applied


def test_in_order_program_is_the_arrival_list(oracle):
    sc = scenario(2, 250)
    ev = sc["st"]["events"]
    ops, cnt = program_from_arrivals(ev, snapshot_slots=0, snapshot_period_us=0)
    assert len(ops) == len(ev) and cnt["rewinds"] == 0 and cnt["snapshots"] == 0 and cnt["discarded"] == 0
    assert [tuple(o)[:4] for o in ops.tolist()] == [tuple(e)[:4] for e in ev]
    check(oracle, sc, ev, ops)


@pytest.mark.parametrize("latency,period,phase", [(50, 100_000, 1000), (50, 20_000, 0), (7, 100_000, 1000), (130, 50_000, 1000)])
def test_delayed_pose_fixes_rewind_like_the_reference(oracle, latency, period, phase):
    T = 420
    sc = scenario(2, T)
    ev = sc["st"]["events"]
    arrivals = delay_pose(ev, latency)
    ops, cnt = program_from_arrivals(arrivals, snapshot_slots=4, snapshot_period_us=period, snapshot_phase_us=phase)
    n_late = sum(1 for i, e in enumerate(arrivals) if i > 0 and e[3] < max(a[3] for a in arrivals[:i]))
    assert cnt["rewinds"] == n_late and cnt["discarded"] == 0
    ref, n_applied = check(oracle, sc, arrivals, ops, n=1)
    if (period, phase) == (100_000, 1000):
        # snapshots sit exactly at the pose-fix stamps: the program re-applies what the reference re-applies
        assert n_applied == ref["calls"]
    else:
        assert n_applied >= ref["calls"]
    # and both equal plain in-order processing
    inorder = oracle_history(oracle, sc, ev, 1)
    assert np.array_equal(inorder["vec"], ref["vec"]) and np.array_equal(inorder["cov"], ref["cov"])


def test_equal_utime_keeps_arrival_order_and_batched_roll_forward(oracle):
    sc = scenario(1, 120)
    ev = sc["st"]["events"]
    # pose fix of step 0 arrives after step 10, leg odometry rows arrive two steps late, roll forward every 5 arrivals
    arrivals = delay_pose(ev, 10)
    lego = [e for e in arrivals if e[0] == 1 and e[1] == 0]
    rest = [e for e in arrivals if not (e[0] == 1 and e[1] == 0)]
    mixed, pend = [], list(lego)
    for e in rest:
        mixed.append(e)
        while pend and e[0] == 0 and e[3] >= pend[0][3] + 2000:
            mixed.append(pend.pop(0))
    mixed += pend
    ops, cnt = program_from_arrivals(mixed, snapshot_slots=3, snapshot_period_us=10_000, snapshot_phase_us=1000,
                                     roll_forward_every=5)
    assert cnt["rewinds"] > 0
    # the oracle rolls forward on every arrival; batching roll-forwards must not change the result
    check(oracle, sc, mixed, ops)


def test_too_old_updates_are_discarded(oracle):
    sc = scenario(1, 300)
    ev = sc["st"]["events"]
    span = 30_000
    p = Planner(utime0=0, snapshot_slots=4, snapshot_period_us=10_000, snapshot_phase_us=1000, history_span_us=span)
    for e in ev:
        assert p.add_update(*e)
    assert p.counters()["retained"] < 60
    assert not p.add_update(capi.OP_MEAS, 0, 0, 1000, 0.0)          # far older than the retained history
    assert not p.add_update(capi.OP_MEAS, 0, 0, -5, 0.0)            # older than the reset itself
    late_ok = ev[-1][3] - 15_000
    row = [e for e in ev if e[0] == 1 and e[1] == 0 and e[3] <= late_ok][-1]
    assert p.add_update(capi.OP_MEAS, 0, row[2], late_ok, 0.0)     # inside the span: accepted and replayed
    c = p.counters()
    assert c["discarded"] == 2 and c["rewinds"] == 1
    ops = p.take()
    arrivals = list(ev) + [(capi.OP_MEAS, 0, 0, 1000, 0.0), (capi.OP_MEAS, 0, row[2], late_ok, 0.0)]
    check(oracle, sc, arrivals, ops, history_span=span)
    with pytest.raises(capi.RBISError):
        p.add_update(capi.OP_SNAPSHOT, 0, 0, 0, 0.0)
    p.close()


def test_without_snapshots_late_updates_are_dropped():
    p = Planner(snapshot_slots=0, snapshot_period_us=0)
    assert p.add_update(capi.OP_IMU, 0, 0, 1000, 1e-3)
    assert p.add_update(capi.OP_IMU, 0, 1, 2000, 1e-3)
    assert not p.add_update(capi.OP_MEAS, 0, 0, 1000, 0.0)
    assert p.add_update(capi.OP_MEAS, 0, 0, 2000, 0.0)  # equal to the head: in order
    assert len(p.take()) == 3
    p.close()
