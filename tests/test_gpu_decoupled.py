"""GPU tests of the decoupled kernel variant (rbis_batch_config_t::dense_only, rbis_kernels.cuh "decoupled filters").

The variant keeps only the 15x15 block of v, chi, p, b_g, b_a on chip.  It is chosen at run time when every filter's
covariance couplings to the omega / a rows are exactly zero (the reference's diagonal initial covariance,
MSE/rbis_initializer.cpp:85-91) and must then be BIT-IDENTICAL to the dense variant, which carries the whole 21x21
covariance as MSE/rbis.cpp:77-143 does; both are compared with the CPU oracle.
"""
import os

import numpy as np
import pytest

from pronto_b200 import MeasStream, RBISBatch, capi, synth
from pronto_b200.schedule import program_from_arrivals

from common import gpu_streams, nominal_q, oracle_streams, random_ensemble, scenario

pytestmark = pytest.mark.gpu
NTHREADS = min(16, os.cpu_count() or 1)
DENSE, DENSE_GENERAL, DECOUPLED = 0, 1, 2


def _same(a, b):
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)


def _run(sc, ops, dense_only, snapshot_slots=0, launch_groups=0):
    st = sc["st"]
    with RBISBatch(sc["vec"].shape[1], dense_only=dense_only, snapshot_slots=snapshot_slots, launch_groups=launch_groups) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(ops, imu=st["imu"], streams=gpu_streams(st))
        return b.get_state(), (b.last_kernel_variant & 3)


@pytest.mark.parametrize("N", [1, 500, 1000])
def test_decoupled_is_chosen_for_diagonal_p0_and_equals_dense_and_oracle(oracle, N):
    T = 260
    sc = scenario(N, T, tumbling=True)
    st = sc["st"]
    got, variant = _run(sc, st["events"], dense_only=False)
    ref, ref_variant = _run(sc, st["events"], dense_only=True)
    assert variant == DECOUPLED and ref_variant == DENSE
    _same(got, ref)
    cov = got[2].reshape(21, 21, N)
    q = nominal_q()
    pas = [0, 1, 2, 12, 13, 14]
    act = [k for k in range(21) if k not in pas]
    assert np.all(cov[np.ix_(pas, act)] == 0) and np.all(cov[np.ix_(act, pas)] == 0)  # couplings stay exactly zero
    for k in range(3):  # overwrites of rbis.cpp:120-121
        assert np.all(cov[k, k] == q[0]) and np.all(cov[12 + k, 12 + k] == q[1])
    orc = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, q, st["imu"], oracle_streams(st), st["events"],
                              n_threads=NTHREADS)
    assert np.max(np.abs(got[0] - orc["vec"])) < 1e-9 and np.max(np.abs(got[2] - orc["cov"])) < 1e-11


def test_measurement_only_program_keeps_the_overwritten_blocks_as_they_were():
    """Without an IMU step in the program the (omega,omega) / (a,a) blocks must come back untouched."""
    N = 70
    sc = scenario(N, 4)
    sc["cov"] = sc["cov"].copy()
    c = sc["cov"].reshape(21, 21, N)
    rng = np.random.default_rng(3)
    for blk in (slice(0, 3), slice(12, 15)):
        A = rng.normal(size=(3, 3))
        c[blk, blk, :] = (A @ A.T + np.eye(3))[:, :, None] * 1e-3
    ops = [e for e in sc["st"]["events"] if e[0] == capi.OP_MEAS][:2]
    got, variant = _run(sc, ops, dense_only=False)
    ref, _ = _run(sc, ops, dense_only=True)
    assert variant == DECOUPLED
    _same(got, ref)
    assert np.array_equal(got[2].reshape(21, 21, N)[0:3, 0:3], c[0:3, 0:3])


def test_coupled_covariance_takes_the_dense_kernel_and_diagonal_reset_returns_to_decoupled(oracle):
    N, T = 300, 40
    sc = scenario(N, T)
    st = sc["st"]
    vec, quat, cov = random_ensemble(N, seed=4)  # dense SPD covariances
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(vec, quat, cov)
        b.run_fused(st["events"], imu=st["imu"], streams=gpu_streams(st))
        assert (b.last_kernel_variant & 3) == DENSE
        got = b.get_state()
        b.run_fused(st["events"][:5], imu=st["imu"], streams=gpu_streams(st))
        assert (b.last_kernel_variant & 3) == DENSE  # remembered, no re-check needed
        b.set_state(sc["vec"], sc["quat"], sc["cov"])  # RBISResetUpdate with a diagonal covariance
        b.run_fused(st["events"], imu=st["imu"], streams=gpu_streams(st))
        assert (b.last_kernel_variant & 3) == DECOUPLED
        # one filter replaced by a coupled one: the ensemble is no longer decoupled
        b.set_filter(7, vec[:, 7], quat[:, 7], cov[:, 7], 0.0)
        b.run_fused(st["events"][:5], imu=st["imu"], streams=gpu_streams(st))
        assert (b.last_kernel_variant & 3) == DENSE
    orc = oracle.run_ensemble(vec, quat, cov, None, 0, nominal_q(), st["imu"], oracle_streams(st), st["events"],
                              n_threads=NTHREADS)
    assert np.max(np.abs(got[0] - orc["vec"])) < 1e-9 and np.max(np.abs(got[2] - orc["cov"])) < 1e-11


def test_measuring_omega_couples_the_filters_and_later_launches_go_dense():
    N, T = 200, 30
    sc = scenario(N, T)
    st = sc["st"]
    rng = np.random.default_rng(8)
    idx = [3, 4, 5, 0, 1, 2]  # velocity + angular velocity (rbis_legodo_common.cpp:59-67)
    z = np.ascontiguousarray(rng.normal(size=(6, N)) * 0.1)
    B = rng.normal(size=(6, 6))
    R = B @ B.T * 0.01 + np.eye(6) * 0.01  # correlated noise: the update couples omega to the velocity
    ev = st["events"]
    half = len(ev) // 2
    out = []
    for dense_only in (False, True):
        with RBISBatch(N, dense_only=dense_only) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            b.run_fused(ev[:half], imu=st["imu"], streams=gpu_streams(st))
            v0 = (b.last_kernel_variant & 3)
            b.indexed_update(idx, z, R, utime=ev[half - 1][3])
            v1 = (b.last_kernel_variant & 3)
            b.run_fused(ev[half:], imu=st["imu"], streams=gpu_streams(st))
            v2 = (b.last_kernel_variant & 3)
            out.append((b.get_state(), (v0, v1, v2)))
    assert out[0][1] == (DECOUPLED, DENSE_GENERAL, DENSE) and out[1][1] == (DENSE, DENSE_GENERAL, DENSE)
    _same(out[0][0], out[1][0])
    c = out[0][0][2].reshape(21, 21, N)
    assert np.any(c[0:3, 3:6] != 0)  # the couplings did become non-zero


def test_scalar_updates_stay_on_the_decoupled_kernel_and_match_dense_and_oracle(oracle):
    """One-row chunks (yaw lock [17, 8] with orientation, rbis_yawlock_update.cpp:97-99; a plain scalar on the yaw bias,
    :79; an unaligned diagonal triple) take meas1 and do not need the general path."""
    N, T = 260, 60
    sc = scenario(N, T, tumbling=True)
    st = sc["st"]
    rng = np.random.default_rng(12)
    n_extra = T // 5
    yaw_z = np.ascontiguousarray(rng.normal(size=(n_extra, 2, N)) * 1e-3)
    yaw_q = np.ascontiguousarray(st["pose_q"][:1].repeat(n_extra, axis=0))
    bias_z = np.ascontiguousarray(rng.normal(size=(n_extra, 1, N)) * 1e-3)
    odd_z = np.ascontiguousarray(rng.normal(size=(n_extra, 3, N)) * 0.05)
    extra = [dict(idx=[17, 8], z=yaw_z, R=np.diag([1e-4, 1e-2]), quat=yaw_q),
             dict(idx=[17], z=bias_z, R=np.array([[1e-4]])),
             dict(idx=[4, 5, 9], z=odd_z, R=np.diag([0.02, 0.03, 0.04]))]
    ev, k = [], 0
    for e in st["events"]:
        ev.append(e)
        if e[0] == capi.OP_IMU and (e[2] % 5) == 4 and k < n_extra:
            for sidx in (2, 3, 4):
                ev.append((capi.OP_MEAS, sidx, k, e[3], 0.0))
            k += 1
    gs = gpu_streams(st) + [MeasStream(x["idx"], x["z"], x["R"], quat=x.get("quat")) for x in extra]
    out = []
    for dense_only in (False, True):
        with RBISBatch(N, dense_only=dense_only) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            b.run_fused(ev, imu=st["imu"], streams=gs)
            out.append((b.get_state(), (b.last_kernel_variant & 3)))
    assert out[0][1] == DECOUPLED + 1 and out[1][1] == DENSE + 1  # + 1: the instantiations with meas1 / meas_block
    _same(out[0][0], out[1][0])
    orc = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st) + extra, ev,
                              n_threads=NTHREADS)
    got = out[0][0]
    assert np.max(np.abs(got[0] - orc["vec"])) < 1e-9 and np.max(np.abs(got[2] - orc["cov"])) < 1e-11
    assert np.max(np.abs(got[3] - orc["loglik"]) / np.maximum(1.0, np.abs(orc["loglik"]))) < 1e-9


def _delayed_program(sc, n_slots=3, lat=50):
    ev = sc["st"]["events"]
    pose = [e for e in ev if e[0] == 1 and e[1] == 1]
    arrivals, pending = [], list(pose)
    for e in ev:
        if e[0] == 1 and e[1] == 1:
            continue
        arrivals.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + lat * 1000:
            arrivals.append(pending.pop(0))
    arrivals += pending
    ops, cnt = program_from_arrivals(arrivals, snapshot_slots=n_slots, snapshot_period_us=100_000, snapshot_phase_us=1000)
    assert cnt["rewinds"] == len(pose)
    return ops


def test_rewind_program_decoupled_equals_dense():
    N, T = 450, 400
    sc = scenario(N, T)
    ops = _delayed_program(sc)
    got, variant = _run(sc, ops, dense_only=False, snapshot_slots=3)
    ref, _ = _run(sc, ops, dense_only=True, snapshot_slots=3)
    assert variant == DECOUPLED
    _same(got, ref)


def test_snapshots_cross_launches_and_variants():
    """A slot written by a decoupled launch may be restored by a later decoupled launch; a slot written while the
    ensemble was coupled forces the restoring launch onto the dense kernel."""
    N, T = 130, 60
    sc = scenario(N, T)
    st = sc["st"]
    ev = st["events"]
    cut = len(ev) // 2
    snap = (capi.OP_SNAPSHOT, 0, 0, ev[cut - 1][3], 0.0)
    restore = (capi.OP_RESTORE, 0, 0, ev[cut - 1][3], 0.0)
    vec, quat, cov = random_ensemble(N, seed=11)
    out = []
    for dense_only in (False, True):
        with RBISBatch(N, dense_only=dense_only, snapshot_slots=2) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            b.run_fused(list(ev[:cut]) + [snap], imu=st["imu"], streams=gpu_streams(st))
            va = (b.last_kernel_variant & 3)
            b.run_fused(ev[cut:], imu=st["imu"], streams=gpu_streams(st))
            b.run_fused([restore] + list(ev[cut:]), imu=st["imu"], streams=gpu_streams(st))
            vb = (b.last_kernel_variant & 3)
            first = b.get_state()
            # now a coupled ensemble writes slot 1 ...
            b.set_state(vec, quat, cov)
            b.run_fused(list(ev[:cut]) + [(capi.OP_SNAPSHOT, 0, 1, ev[cut - 1][3], 0.0)], imu=st["imu"], streams=gpu_streams(st))
            # ... and a decoupled ensemble restores it
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            b.run_fused([(capi.OP_RESTORE, 0, 1, ev[cut - 1][3], 0.0)] + list(ev[cut:]), imu=st["imu"], streams=gpu_streams(st))
            vc = (b.last_kernel_variant & 3)
            second = b.get_state()
            b.run_fused(ev[:4], imu=st["imu"], streams=gpu_streams(st))
            vd = (b.last_kernel_variant & 3)
            out.append((first, second, (va, vb, vc, vd)))
    assert out[0][2] == (DECOUPLED, DECOUPLED, DENSE, DENSE) and out[1][2] == (DENSE,) * 4
    _same(out[0][0], out[1][0])
    _same(out[0][1], out[1][1])


def test_launch_groups_with_the_decoupled_grid():
    N, T = 384 * 5 + 17, 50
    sc = scenario(N, T)
    ev = sc["st"]["events"]
    a, va = _run(sc, ev, dense_only=False, launch_groups=3)
    b, vb = _run(sc, ev, dense_only=False, launch_groups=1)
    assert va == DECOUPLED and vb == DECOUPLED
    _same(a, b)


def test_variant_flip_between_grouped_launches_is_race_free():
    """Launch groups cut the ensemble into per-group filter ranges of (CTAs per group) x (filters per CTA), and the filters per
    CTA differ between the dense (256) and decoupled (384) kernels: when the variant flips between two overlapped launches,
    every group must wait for ALL groups of the previous launch.  65,536 filters, 6 groups: decoupled, then a launch that
    measures omega (dense), then decoupled-eligible programs again; results must equal launch_groups = 1 bit for bit."""
    import torch

    N, T = 65_536, 24
    base = scenario(256, T)
    rep = lambda a: np.ascontiguousarray(np.tile(a, (1,) * (a.ndim - 1) + (N // 256,)))
    st = base["st"]
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(rep(a)).to(dev)
    imu, lego, pz, pq = t(st["imu"]), t(st["legodo"]), t(st["pose_z"]), t(st["pose_q"])
    vec, quat, cov = t(base["vec"]), t(base["quat"]), t(base["cov"])
    rng = np.random.default_rng(11)
    zw = torch.from_numpy(rng.normal(size=(T, 6, N)) * 0.05).to(dev)   # velocity + angular velocity rows (rbis_legodo_common.cpp:59-67)
    Rw = np.diag([0.01] * 3 + [0.02] * 3)
    ev = st["events"]
    third = len(ev) // 3
    progs = [list(ev[:third]),
             list(ev[third:2 * third]) + [(capi.OP_MEAS, 2, 0, ev[2 * third - 1][3], 0.0)],
             list(ev[2 * third:])]
    out = []
    for groups in (1, 6):
        with RBISBatch(N, launch_groups=groups, mapping=1, lane_filters_per_cta=384) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(vec, quat, cov)
            variants = []
            for k, prog in enumerate(progs):
                streams = [MeasStream(synth.LEGODO_IDX, lego, st["R_legodo"]), MeasStream(synth.POSE_IDX, pz, st["R_pose"], quat=pq)]
                if k == 1:
                    streams.append(MeasStream([3, 4, 5, 0, 1, 2], zw, Rw))
                b.run_fused(prog, imu=imu, streams=streams)
                variants.append(b.last_kernel_variant & 3)
            gv = torch.empty((21, N), dtype=torch.float64, device=dev)
            gc = torch.empty((441, N), dtype=torch.float64, device=dev)
            gl = torch.empty((N,), dtype=torch.float64, device=dev)
            b.get_state_into(vec=gv, cov=gc, loglik=gl)
            b.synchronize()
            out.append((gv.cpu().numpy(), gc.cpu().numpy(), gl.cpu().numpy(), variants))
    assert out[0][3] == [DECOUPLED, DENSE, DENSE] and out[1][3] == out[0][3]
    for a, c in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(a, c)
