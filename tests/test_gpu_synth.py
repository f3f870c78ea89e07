"""On-device synthesis of the per-filter sensor streams (rbis_batch_synthesize / rbis_batch_run_fused_synth).

mode 0 is the splitmix64 -> Box-Muller generator of pronto_b200/synth.py:normal: the counters and the hash are integers (bit
exact), the three libm calls differ from numpy's by an ulp or two.  A fused run over synthesised inputs must equal, bit for
bit, a fused run over the same rows materialised with rbis_batch_synthesize -- which is also how the CPU oracle is fed
exactly what the device drew."""
import os

import numpy as np
import pytest

from pronto_b200 import MeasStream, RBISBatch, SynthSpec, synth
from pronto_b200.parity import max_errors

from common import nominal_q, oracle_streams, scenario

pytestmark = pytest.mark.gpu
NTHREADS = min(16, os.cpu_count() or 1)


def _spec(truth, k0, T, mode=0, first_filter=0, **kw):
    d = synth.synth_spec_inputs(truth, k0, T)
    return SynthSpec(synth.SEED, d["imu_mean"], d["imu_step"], d["streams"], mode=mode, first_filter=first_filter, **kw)


def _streams(st):
    return [MeasStream(synth.LEGODO_IDX, None, st["R_legodo"]), MeasStream(synth.POSE_IDX, None, st["R_pose"], quat=True)]


def test_exact_mode_reproduces_the_numpy_generator():
    N, T = 700, 230
    sc = scenario(N, T)
    st = sc["st"]
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        got = b.synthesize(_spec(sc["truth"], 0, T))
    for name, g, ref in (("imu", got["imu"], st["imu"]), ("legodo", got["z"][0], st["legodo"]), ("pose_z", got["z"][1], st["pose_z"]),
                         ("pose_q", got["quat"][1], st["pose_q"])):
        g = g.cpu().numpy()
        assert g.shape == ref.shape, name
        assert np.max(np.abs(g - ref)) <= 2e-13 * max(1.0, np.max(np.abs(ref))), (name, np.max(np.abs(g - ref)))
    # explicit sigmas instead of sqrt(q / dt) give the same rows
    p = synth.NOMINAL
    with RBISBatch(N) as b:
        again = b.synthesize(_spec(sc["truth"], 0, T, sigma_gyro=np.sqrt(p["q_gyro"] / p["dt"]), sigma_accel=np.sqrt(p["q_accel"] / p["dt"])))
    assert np.max(np.abs(again["imu"].cpu().numpy() - got["imu"].cpu().numpy())) < 1e-15


@pytest.mark.parametrize("mode,mapping,materialize", [(0, 0, False), (1, 1, False), (1, 4, False), (1, 8, False), (1, 1, True), (1, 0, False)])
def test_fused_run_over_synthesised_inputs_equals_run_over_materialised_rows(oracle, mode, mapping, materialize):
    """mode 1 on a decoupled ensemble draws the rows INSIDE the fused kernel (SYN instantiations of the lane-per-filter and
    warp-group kernels) unless synth_materialize is set; mode 0 always materialises them first.  Same bits every way."""
    N, T = 300, 220
    sc = scenario(N, T, tumbling=True)
    st = sc["st"]
    with RBISBatch(N, mapping=mapping, synth_materialize=materialize) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        spec = _spec(sc["truth"], 0, T, mode=mode)
        l0 = b.launch_count
        b.run_fused_synth(st["events"], _streams(st), spec)
        # fused synthesis = the coupling check + the fused kernel(s) and nothing else; materialising adds three generator kernels
        n_launched = b.launch_count - l0
        assert (n_launched <= 2) == (mode == 1 and not materialize), n_launched
        assert bool(b.last_kernel_variant & 4) == (mode == 1 and not materialize)   # the SYN kernel instantiation ran
        if mapping > 1:
            assert b.last_kernel_variant >> 4 == mapping
        a = b.get_state()
        rows = b.synthesize(spec)
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(st["events"], imu=rows["imu"], streams=[MeasStream(synth.LEGODO_IDX, rows["z"][0], st["R_legodo"]),
                                                            MeasStream(synth.POSE_IDX, rows["z"][1], st["R_pose"], quat=rows["quat"][1])])
        c = b.get_state()
    for x, y in zip(a[:4], c[:4]):
        assert np.array_equal(x, y)
    # ... and the CPU oracle fed with what the device drew agrees to the per-step gate
    orc = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), rows["imu"].cpu().numpy(),
                              [dict(idx=synth.LEGODO_IDX, z=rows["z"][0].cpu().numpy(), R=st["R_legodo"]),
                               dict(idx=synth.POSE_IDX, z=rows["z"][1].cpu().numpy(), R=st["R_pose"], quat=rows["quat"][1].cpu().numpy())],
                              st["events"], n_threads=NTHREADS)
    e = max_errors(a[0], a[1], a[2], orc["vec"], orc["quat"], orc["cov"])
    assert max(e.values()) < 1e-9, e
    if mode == 1:  # the fast generator draws unit-variance noise too
        nz = (rows["z"][0].cpu().numpy() - sc["truth"]["v"][::2][:, :, None]) / synth.NOMINAL["r_vxyz"]
        assert abs(nz.std() - 1.0) < 0.02 and abs(nz.mean()) < 0.02 and np.abs(nz).max() < 6.0


def test_shards_draw_the_noise_of_their_global_filter_indices():
    N, T = 512, 60
    sc = scenario(N, T)
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        whole = b.synthesize(_spec(sc["truth"], 0, T, mode=1))["imu"].cpu().numpy()
    for lo in (0, 256):
        with RBISBatch(256) as b:
            b.set_process_noise(*nominal_q())
            part = b.synthesize(_spec(sc["truth"], 0, T, mode=1, first_filter=lo))["imu"].cpu().numpy()
        assert np.array_equal(part, whole[:, :, lo:lo + 256])


def test_consecutive_synth_launches_overlap_safely():
    """Chunk after chunk through the double-buffered staging slots (with launch groups): same bits as one long program."""
    N, T, CH = 4000, 120, 40
    sc = scenario(N, T)
    st = sc["st"]
    ev = st["events"]
    out = []
    for groups, piece in ((1, 0), (4, 0), (4, 16)):
        with RBISBatch(N, launch_groups=groups, mapping=1, piece_ops=piece) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            for k0 in range(0, T, CH):
                sub = synth.make_streams(sc["truth"], 1, k0, CH)  # chunk-relative rows of the schedule
                b.run_fused_synth(sub["events"], _streams(st), _spec(sc["truth"], k0, CH, mode=1))
            out.append(b.get_state())
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused_synth(ev, _streams(st), _spec(sc["truth"], 0, T, mode=1))
        out.append(b.get_state())
    for o in out[1:]:
        for x, y in zip(out[0][:4], o[:4]):
            assert np.array_equal(x, y)


def test_full_size_monte_carlo_ensemble_matches_oracle_on_scattered_filters(oracle):
    """BASELINE configs[2] at full size: 65,536 DISTINCT filters (counter-based initial ensemble and sensor rows, drawn inside the
    fused lane-per-filter kernel) through 300 steps of IMU + leg odometry + pose fixes; 384 filters scattered over the ensemble are
    replayed by the CPU oracle from the rows the device drew for exactly those filters."""
    import torch

    N, T, S = 65536, 300, 384
    truth = synth.truth_trajectory(T)
    tv0 = np.zeros(21)
    tv0[9:12] = (0, 0, 0.85); tv0[15:18] = synth.NOMINAL["bg"]; tv0[18:21] = synth.NOMINAL["ba"]
    vec0, quat0, cov0 = synth.initial_ensemble(N, tv0, truth["quat"][0])
    st1 = synth.make_streams(truth, 1, 0, T)
    ev = st1["events"]
    streams = [MeasStream(synth.LEGODO_IDX, None, st1["R_legodo"]), MeasStream(synth.POSE_IDX, None, st1["R_pose"], quat=True)]
    pick = np.sort(np.random.default_rng(5).choice(N, size=S, replace=False))
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(vec0, quat0, cov0)
        spec = _spec(truth, 0, T, mode=1)
        b.run_fused_synth(ev, streams, spec)
        assert b.last_kernel_variant == 2 + 4      # decoupled lane-per-filter kernel, SYN instantiation
        gv, gq, gP, gll, _ = b.get_state()
        rows = b.synthesize(spec)
        sel = torch.from_numpy(pick).to(rows["imu"].device)
        take = lambda t: t.index_select(t.ndim - 1, sel).cpu().numpy()
        imu, lz, pz, pq = take(rows["imu"]), take(rows["z"][0]), take(rows["z"][1]), take(rows["quat"][1])
    assert np.isfinite(gv).all() and np.isfinite(gP).all()
    sub = lambda a: np.ascontiguousarray(a[..., pick])
    orc = oracle.run_ensemble(sub(vec0), sub(quat0), sub(cov0), None, 0, nominal_q(), imu,
                              [dict(idx=synth.LEGODO_IDX, z=lz, R=st1["R_legodo"]), dict(idx=synth.POSE_IDX, z=pz, R=st1["R_pose"], quat=pq)],
                              ev, n_threads=NTHREADS)
    e = max_errors(sub(gv), sub(gq), sub(gP), orc["vec"], orc["quat"], orc["cov"])
    assert max(e.values()) < 1e-9, e
    assert np.max(np.abs(sub(gll) - orc["loglik"]) / np.maximum(1.0, np.abs(orc["loglik"]))) < 1e-9
    # the ensemble is an ensemble: distinct filters, consistent with the truth
    assert np.unique(gv[3]).size > 0.99 * N
    tvT, _ = synth.truth_state_at(truth, T - 1)
    assert np.sqrt(np.mean((gv[9:12] - tvT[9:12, None]) ** 2)) < 0.05
