"""GPU tests of the warp-group kernels (rbis_batch_config_t::mapping = 2 / 4 / 8 / 16 lanes per filter, rbis_group.cuh) and
of the lane-per-filter kernels at 384 / 256 / 128 filters per CTA.

The mapping is a scheduling decision: every variant must give results BIT-IDENTICAL to the 384-per-CTA lane-per-filter
kernels -- covariance products of MSE/rbis.cpp:113-118 (Ad cov Ad^T + Qd) and :134-140 (S, K, K C cov) split over the
lanes of a group or not -- and all of them are compared with the CPU oracle.  The automatic choice (mapping = 0) is what
every other GPU test of this suite runs through: for their small ensembles that is a warp-group kernel.
"""
import os

import numpy as np
import pytest

from pronto_b200 import MeasStream, RBISBatch, capi, synth
from pronto_b200.parity import max_errors
from pronto_b200.schedule import program_from_arrivals

from common import gpu_streams, nominal_q, oracle_streams, random_ensemble, scenario

pytestmark = pytest.mark.gpu
NTHREADS = min(16, os.cpu_count() or 1)
GROUPS = [2, 4, 8, 16]


def _run(sc, ops, streams=None, imu=None, snapshot_slots=0, **cfg):
    st = sc["st"]
    with RBISBatch(sc["vec"].shape[1], snapshot_slots=snapshot_slots, **cfg) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(ops, imu=st["imu"] if imu is None else imu, streams=gpu_streams(st) if streams is None else streams)
        return b.get_state(), b.last_kernel_variant


def _same(a, b, what=""):
    for k, (x, y) in enumerate(zip(a[:4], b[:4])):
        assert np.array_equal(x, y), (what, k, float(np.max(np.abs(x - y))))


@pytest.mark.parametrize("N", [1, 37, 333])
def test_every_mapping_gives_the_same_bits_and_matches_the_oracle(oracle, N):
    """Config-3 program on a tumbling trajectory (IMU + leg odometry + pose fixes with orientation), decoupled ensemble."""
    T = 230
    sc = scenario(N, T, tumbling=True)
    st = sc["st"]
    ref, v0 = _run(sc, st["events"], mapping=1, lane_filters_per_cta=384)
    assert v0 == 2
    for tpb in (256, 128):
        got, v = _run(sc, st["events"], mapping=1, lane_filters_per_cta=tpb)
        assert v == 2
        _same(got, ref, f"lane {tpb}")
    for G in GROUPS:
        got, v = _run(sc, st["events"], mapping=G)
        assert v == 2 + 16 * G
        _same(got, ref, f"group {G}")
    orc = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st), st["events"],
                              n_threads=NTHREADS)
    e = max_errors(ref[0], ref[1], ref[2], orc["vec"], orc["quat"], orc["cov"])
    assert e["vec"] < 1e-9 and e["quat"] < 1e-9 and e["cov"] < 1e-9, e


@pytest.mark.parametrize("G", GROUPS)
def test_group_kernel_per_step_parity_with_the_oracle(oracle, G):
    """Every event as its own launch against the oracle's trace (<= 1e-12 per IMU step, 1e-9 gate overall), and the
    same program in one launch: same bits."""
    N, T = 45, 60
    sc = scenario(N, T, tumbling=True)
    st = sc["st"]
    ev = st["events"]
    ref = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st), ev,
                              n_threads=NTHREADS, trace=True)
    worst = 0.0
    with RBISBatch(N, mapping=G) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        streams = gpu_streams(st)
        for e, event in enumerate(ev):
            b.run_fused([event], imu=st["imu"], streams=streams)
            gv, gq, gP, gll, _ = b.get_state()
            err = max_errors(gv, gq, gP, ref["trace_vec"][e], ref["trace_quat"][e], ref["trace_cov"][e])
            worst = max(worst, err["vec"], err["quat"], err["cov"])
        step_by_step = (gv, gq, gP, gll)
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(ev, imu=st["imu"], streams=streams)
        fused = b.get_state()
    assert worst <= 1e-11, worst
    _same(step_by_step, fused, "launch fusion")


GENERAL_CASES = [
    ([17], False, "diag"),                            # yaw-bias (rbis_yawlock_update.cpp:79): one-row chunk
    ([17, 8], True, "diag"),                          # yawlock (rbis_yawlock_update.cpp:97-99)
    ([8, 9, 10, 11], False, "dense"),                 # quick-lock (quick_lock.cpp:132): correlated block
    ([3, 4, 5, 0, 1, 2], False, "diag"),              # velocity + angular velocity: couples omega -> dense kernels
    ([9, 10, 11, 3, 4, 5, 6, 7, 8], True, "dense"),   # m = 9 (laser_gpf_lib.cpp:108-110)
    ([9, 10, 11, 8], True, "block"),                  # scan match (sensor_handlers.cpp:653-684): 3-block + 1
]


@pytest.mark.parametrize("idx,orient,rkind", GENERAL_CASES)
@pytest.mark.parametrize("coupled", [False, True])
def test_general_measurement_paths_same_bits_for_every_mapping(oracle, idx, orient, rkind, coupled):
    """One-row chunks, correlated blocks and omega / a indices through IMU steps + updates; decoupled (diagonal P0) and
    coupled (dense SPD P0 -> the dense kernels) ensembles."""
    N, T, m = 70, 12, len(idx)
    sc = scenario(N, T, with_pose=False)
    if coupled:
        _, _, sc["cov"] = random_ensemble(N, seed=m)
    st = sc["st"]
    rng = np.random.default_rng(m * 3 + len(rkind))
    rows = 3
    z = np.ascontiguousarray(rng.normal(size=(rows, m, N)) * 0.1 + sc["vec"][idx, :][None])
    mq = None
    if orient:
        d = rng.normal(size=(rows, 3, N)) * 0.03
        n = np.linalg.norm(d, axis=1, keepdims=True)
        mq = np.ascontiguousarray(np.concatenate([np.cos(n / 2), np.sin(n / 2) * d / n], axis=1))
    if rkind == "diag":
        R = np.diag(np.abs(rng.normal(size=m)) * 0.01 + 0.005)
    elif rkind == "dense":
        B = rng.normal(size=(m, m))
        R = B @ B.T * 0.01 + np.eye(m) * 0.01
    else:
        B = rng.normal(size=(3, 3))
        R = np.zeros((m, m))
        R[:3, :3] = B @ B.T * 0.01 + np.eye(3) * 0.01
        R[3, 3] = 0.02
    ev = []
    for k in range(T):
        ev.append((capi.OP_IMU, 0, k, (k + 1) * 1000, 1e-3))
        if k % 2 == 0:
            ev.append((capi.OP_MEAS, 0, k // 2, (k + 1) * 1000, 0.0))
        if k % 4 == 1:
            ev.append((capi.OP_MEAS, 1, k // 4, (k + 1) * 1000, 0.0))
    streams = lambda: [MeasStream(synth.LEGODO_IDX, st["legodo"], st["R_legodo"]), MeasStream(idx, z, R, quat=mq)]
    ref, v0 = _run(sc, ev, streams=streams(), mapping=1, lane_filters_per_cta=384)
    for G in GROUPS:
        got, v = _run(sc, ev, streams=streams(), mapping=G)
        assert (v & 3) == (v0 & 3) | 1 or (v & 2) == (v0 & 2)
        _same(got, ref, f"group {G}")
    orc = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"],
                              [dict(idx=synth.LEGODO_IDX, z=st["legodo"], R=st["R_legodo"]), dict(idx=idx, z=z, R=R, quat=mq)], ev,
                              n_threads=NTHREADS)
    e = max_errors(ref[0], ref[1], ref[2], orc["vec"], orc["quat"], orc["cov"])
    assert e["vec"] < 1e-9 and e["quat"] < 1e-9 and e["cov"] < 1e-9, e


@pytest.mark.parametrize("G", [4, 8])
def test_rewind_program_with_snapshots_same_bits(oracle, G):
    """Delayed pose fixes: SNAPSHOT / RESTORE ops inside the program (config 5), decoupled ensemble."""
    N, T, LAT = 50, 300, 50
    sc = scenario(N, T)
    st = sc["st"]
    ev = st["events"]
    pose = [e for e in ev if e[0] == 1 and e[1] == 1]
    arrivals, pending = [], list(pose)
    for e in ev:
        if e[0] == 1 and e[1] == 1:
            continue
        arrivals.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + LAT * 1000:
            arrivals.append(pending.pop(0))
    arrivals += pending
    ops, cnt = program_from_arrivals(arrivals, snapshot_slots=3, snapshot_period_us=100_000, snapshot_phase_us=1000)
    assert cnt["rewinds"] == len(pose)
    ref, _ = _run(sc, ops, snapshot_slots=3, mapping=1, lane_filters_per_cta=384)
    got, _ = _run(sc, ops, snapshot_slots=3, mapping=G)
    _same(got, ref, "rewind")
    orc = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st), arrivals,
                              n_threads=NTHREADS)
    e = max_errors(got[0], got[1], got[2], orc["vec"], orc["quat"], orc["cov"])
    assert max(e.values()) < 1e-9, e


def test_snapshot_written_by_one_mapping_restored_by_another():
    """Ring slots are mapping independent: snapshot under the group kernels, restore + continue under the lane kernels."""
    N, T = 90, 40
    sc = scenario(N, T)
    st = sc["st"]
    ev = st["events"]
    half = len(ev) // 2
    prog = list(ev[:half]) + [(capi.OP_SNAPSHOT, 0, 1, ev[half - 1][3], 0.0)] + list(ev[half:]) + \
           [(capi.OP_RESTORE, 0, 1, ev[half - 1][3], 0.0)] + list(ev[half:])
    ref, _ = _run(sc, prog, snapshot_slots=2, mapping=1, lane_filters_per_cta=384)
    for G in (4, 16):
        got, _ = _run(sc, prog, snapshot_slots=2, mapping=G)
        _same(got, ref, f"group {G}")
    straight, _ = _run(sc, ev, mapping=4)
    _same(straight, ref, "restore + replay == straight run")


def test_automatic_mapping_by_ensemble_size():
    """mapping = 0: small ensembles take a warp-group kernel, larger ones the lane-per-filter kernels."""
    sc = scenario(64, 4)
    _, v = _run(sc, sc["st"]["events"])
    assert v >> 4 in GROUPS
    sc = scenario(20_000, 2)
    _, v = _run(sc, sc["st"]["events"])
    assert v >> 4 == 0 and (v & 2)
