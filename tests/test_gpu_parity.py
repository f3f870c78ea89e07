"""Parity of the CUDA path (through the C ABI, librbis_b200.so) against the CPU oracle and the golden
vectors.  Gates (BASELINE.json north_star / SURVEY.md 8d): <= 1e-9 per step, <= 1e-6 after a 60 s
1 kHz trajectory, ensemble statistics bit-exact for any sharding.  All tests need a B200."""
import os

import numpy as np
import pytest

from pronto_b200 import MeasStream, RBISBatch, capi, reduce_chunks, synth
from pronto_b200.parity import max_errors

from common import gpu_streams, nominal_q, oracle_streams, random_ensemble, scenario

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rbis_golden.npz")
STEP_TOL = 1e-9     # per-step gate
TRAJ_TOL = 1e-6     # after 60 s
NTHREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _assert_close(got, ref, tol, what=""):
    e = max_errors(got[0], got[1], got[2], ref[0], ref[1], ref[2])
    assert e["vec"] <= tol and e["quat"] <= tol and e["cov"] <= tol, (what, e)
    return e


def _rel_ll(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


# ------------------------------------------------------------------------------------------------
# single ops
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 37, 300])
def test_ins_step_matches_oracle(oracle, N):
    vec, quat, cov = random_ensemble(N, seed=N)
    rng = np.random.default_rng(100 + N)
    gyro = rng.normal(size=(3, N)) * 0.5
    accel = rng.normal(size=(3, N)) + np.array([[0], [0], [9.8]])
    qs = [np.abs(rng.normal(size=N)) * s + s for s in (1e-4, 1e-2, 1e-9, 1e-6)]
    with RBISBatch(N) as b:
        b.set_state(vec, quat, cov)
        b.set_process_noise(*qs)
        b.ins_step(gyro, accel, 2e-3, utime=2000)
        gv, gq, gP, gll, ut = b.get_state()
    assert ut == 2000 and np.all(gll == 0)
    rv, rq, rP = np.empty_like(vec), np.empty_like(quat), np.empty_like(cov)
    for n in range(N):
        rv[:, n], rq[:, n] = oracle.ins_update_state(gyro[:, n], accel[:, n], 2e-3, vec[:, n], quat[:, n])
        P = oracle.ins_update_covariance(qs[0][n], qs[1][n], qs[2][n], qs[3][n], vec[:, n], quat[:, n],
                                         cov[:, n].reshape(21, 21).T, 2e-3)
        rP[:, n] = P.T.reshape(-1)
    _assert_close((gv, gq, gP), (rv, rq, rP), 1e-12, "ins_step")


MEAS_CASES = [
    ([3, 4, 5], False, "diag"),            # leg-odometry velocity (rbis_legodo_common.cpp:69-73)
    ([9, 10, 11], False, "diag"),          # pose position (pose_meas.cpp:47,79)
    ([9, 10, 11, 6, 7, 8], True, "diag"),  # pose fix with orientation (pose_meas.cpp:39-52)
    ([17], False, "diag"),                 # yaw-bias (rbis_yawlock_update.cpp:79)
    ([17, 8], True, "diag"),               # yawlock (rbis_yawlock_update.cpp:97-99)
    ([8, 9, 10, 11], False, "dense"),      # quick-lock (quick_lock.cpp:132), dense R -> general path
    ([3, 4, 5, 0, 1, 2], False, "diag"),   # vel + angular velocity (rbis_legodo_common.cpp:59-67)
    ([9, 10, 11, 3, 4, 5, 6, 7, 8], True, "dense"),  # m = 9 (laser_gpf_lib.cpp:108-110)
    ([9, 10, 11, 8], True, "block"),       # scan match (sensor_handlers.cpp:653-684): 3-block + 1
    ([6, 7, 8], False, "diag"),            # plain indexed update ON chi indices (z is used)
]


@pytest.mark.parametrize("idx,orient,rkind", MEAS_CASES)
def test_indexed_update_matches_oracle(oracle, idx, orient, rkind):
    N, m = 130, len(idx)
    vec, quat, cov = random_ensemble(N, seed=m * 7 + len(rkind))
    rng = np.random.default_rng(5 + m)
    # leave a below-tolerance residual chi in a third of the filters
    vec[6:9, ::3] = rng.normal(size=(3, len(range(0, N, 3)))) * 2e-7
    z = np.ascontiguousarray(vec[idx, :] + rng.normal(size=(m, N)) * 0.1)
    mq = None
    if orient:
        d = rng.normal(size=(3, N)) * 0.05
        n = np.linalg.norm(d, axis=0)
        dq = np.stack([np.cos(n / 2), *(np.sin(n / 2) * d / n)])
        w0, x0, y0, z0 = quat
        w1, x1, y1, z1 = dq
        mq = np.ascontiguousarray(np.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                                            w0 * y1 + y0 * w1 + z0 * x1 - x0 * z1, w0 * z1 + z0 * w1 + x0 * y1 - y0 * x1]))
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
    with RBISBatch(N) as b:
        b.set_state(vec, quat, cov, loglik=np.full(N, 1.5))
        b.indexed_update(idx, z, R, utime=77, quat=mq)
        gv, gq, gP, gll, ut = b.get_state()
    assert ut == 77
    rv, rq, rP, rll = np.empty_like(vec), np.empty_like(quat), np.empty_like(cov), np.empty(N)
    for n in range(N):
        pv, pq, pc, ll = oracle.measurement_update(z[:, n], R, idx, vec[:, n], quat[:, n], cov[:, n].reshape(21, 21).T,
                                                   mq[:, n] if orient else None)
        rv[:, n], rq[:, n], rP[:, n], rll[n] = pv, pq, pc.T.reshape(-1), 1.5 + ll
    _assert_close((gv, gq, gP), (rv, rq, rP), 1e-11, f"meas {idx}")
    assert _rel_ll(gll, rll) < 1e-11


def test_per_filter_diagonal_R(oracle):
    N, idx = 70, [3, 4, 5]
    vec, quat, cov = random_ensemble(N, seed=9)
    rng = np.random.default_rng(9)
    z = np.ascontiguousarray(vec[idx, :] + rng.normal(size=(3, N)) * 0.1)
    Rd = np.abs(rng.normal(size=(3, N))) * 0.01 + 0.001
    with RBISBatch(N) as b:
        b.set_state(vec, quat, cov)
        b.indexed_update(idx, z, Rd, per_filter_diag=True)
        gv, gq, gP, gll, _ = b.get_state()
    for n in range(N):
        pv, pq, pc, ll = oracle.measurement_update(z[:, n], np.diag(Rd[:, n]), idx, vec[:, n], quat[:, n],
                                                   cov[:, n].reshape(21, 21).T)
        assert np.max(np.abs(gv[:, n] - pv)) < 1e-12 and np.max(np.abs(gP[:, n] - pc.T.reshape(-1))) < 1e-13
        assert abs(gll[n] - ll) < 1e-11 * max(1, abs(ll))


def test_single_ops_match_golden(golden):
    N = 33
    vec = np.repeat(golden["op_vec"][:, None], N, axis=1)
    quat = np.repeat(golden["op_quat"][:, None], N, axis=1)
    cov = np.repeat(golden["op_cov"].T.reshape(-1)[:, None], N, axis=1)
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(vec, quat, cov)
        b.ins_step(np.repeat(golden["op_gyro"][:, None], N, axis=1), np.repeat(golden["op_accel"][:, None], N, axis=1), 1e-3)
        gv, gq, gP, _, _ = b.get_state()
        assert np.max(np.abs(gv - golden["op_ins_vec"][:, None])) < 1e-14
        assert np.max(np.abs(gq - golden["op_ins_quat"][:, None])) < 1e-14
        assert np.max(np.abs(gP - golden["op_ins_cov"].T.reshape(-1)[:, None])) < 1e-15
        for c in range(int(golden["n_meas_cases"])):
            b.set_state(vec, quat, cov)
            mq = golden[f"m{c}_mq"]
            idx = [int(i) for i in golden[f"m{c}_idx"]]
            b.indexed_update(idx, np.repeat(golden[f"m{c}_z"][:, None], N, axis=1), golden[f"m{c}_R"],
                             quat=np.repeat(mq[:, None], N, axis=1) if mq.size else None)
            gv, gq, gP, gll, _ = b.get_state()
            assert np.max(np.abs(gv - golden[f"m{c}_vec"][:, None])) < 1e-12, c
            assert np.max(np.abs(gq - golden[f"m{c}_quat"][:, None])) < 1e-12, c
            assert np.max(np.abs(gP - golden[f"m{c}_cov"].T.reshape(-1)[:, None])) < 1e-13, c
            assert np.max(np.abs(gll - float(golden[f"m{c}_ll"]))) < 1e-10, c
            assert np.all(gv == gv[:, :1]) and np.all(gP == gP[:, :1])  # every lane computes the same bits


# ------------------------------------------------------------------------------------------------
# fused programs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["walk", "tumble"])
def test_fused_trajectory_matches_golden(golden, name):
    N = 4
    st = scenario(N, 400)["st"]
    streams = [MeasStream(synth.LEGODO_IDX, golden[f"{name}_legodo"], st["R_legodo"]),
               MeasStream(synth.POSE_IDX, golden[f"{name}_pose_z"], st["R_pose"], quat=golden[f"{name}_pose_q"])]
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(golden[f"{name}_in_vec"], golden[f"{name}_in_quat"], golden[f"{name}_in_cov"])
        b.run_fused(st["events"], imu=golden[f"{name}_imu"], streams=streams)
        gv, gq, gP, gll, ut = b.get_state()
    assert ut == st["events"][-1][3]
    _assert_close((gv, gq, gP), (golden[f"{name}_vec"], golden[f"{name}_quat"], golden[f"{name}_cov"]), STEP_TOL, name)
    assert _rel_ll(gll, golden[f"{name}_loglik"]) < STEP_TOL


@pytest.mark.parametrize("tumbling", [False, True])
def test_per_step_parity_and_launch_fusion(oracle, tumbling):
    """Every event applied as its own launch is compared with the oracle's head after the same event
    (<= 1e-9 per step); the same program in ONE launch must give the same bits."""
    N, T = 150, 220
    sc = scenario(N, T, tumbling=tumbling)
    st = sc["st"]
    ev = st["events"]
    ref = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st), ev,
                              n_threads=NTHREADS, trace=True)
    worst = dict(vec=0.0, quat=0.0, cov=0.0, ll=0.0)
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        streams = gpu_streams(st)
        for e, event in enumerate(ev):
            b.run_fused([event], imu=st["imu"], streams=streams)
            gv, gq, gP, gll, _ = b.get_state()
            err = max_errors(gv, gq, gP, ref["trace_vec"][e], ref["trace_quat"][e], ref["trace_cov"][e])
            err["ll"] = _rel_ll(gll, ref["trace_loglik"][e])
            for k in worst:
                worst[k] = max(worst[k], err[k])
        step_by_step = (gv, gq, gP, gll)
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(ev, imu=st["imu"], streams=streams)
        fused = b.get_state()
    assert max(worst.values()) <= STEP_TOL, worst
    for a, c in zip(step_by_step, fused[:4]):
        assert np.array_equal(a, c)


def test_device_resident_inputs_match_host_inputs():
    import torch

    N, T = 260, 64
    sc = scenario(N, T)
    st = sc["st"]
    dev = torch.device("cuda:0")
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(st["events"], imu=st["imu"], streams=gpu_streams(st))
        host = b.get_state()
        t = {k: torch.from_numpy(st[k]).to(dev) for k in ("imu", "legodo", "pose_z", "pose_q")}
        b.set_state(torch.from_numpy(sc["vec"]).to(dev), torch.from_numpy(sc["quat"]).to(dev), torch.from_numpy(sc["cov"]).to(dev))
        torch.cuda.synchronize()
        b.run_fused(st["events"], imu=t["imu"],
                    streams=[MeasStream(synth.LEGODO_IDX, t["legodo"], st["R_legodo"]),
                             MeasStream(synth.POSE_IDX, t["pose_z"], st["R_pose"], quat=t["pose_q"])])
        dvec = torch.empty(21, N, dtype=torch.float64, device=dev)
        dcov = torch.empty(441, N, dtype=torch.float64, device=dev)
        b.get_state_into(vec=dvec, cov=dcov)
        b.synchronize()
        assert np.array_equal(dvec.cpu().numpy(), host[0]) and np.array_equal(dcov.cpu().numpy(), host[2])


def test_60s_trajectory_config1(oracle):
    """BASELINE.json config 1 shape (60 s, 1 kHz IMU + 500 Hz leg odometry) plus 10 Hz pose fixes, run
    in time chunks of 2,000 steps; <= 1e-6 on state and covariance at the end (north_star)."""
    N, T, CH = 16, 60_000, 2_000
    truth = synth.truth_trajectory(T)
    sc0 = scenario(N, 1, truth=truth)
    vec, quat, cov = sc0["vec"], sc0["quat"], sc0["cov"]
    rv, rq, rP, rll = vec.copy(), quat.copy(), cov.copy(), np.zeros(N)
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(vec, quat, cov)
        for k0 in range(0, T, CH):
            st = synth.make_streams(truth, N, k0, CH)
            b.run_fused(st["events"], imu=st["imu"], streams=gpu_streams(st))
            ref = oracle.run_ensemble(rv, rq, rP, rll, k0 * 1000, nominal_q(), st["imu"], oracle_streams(st), st["events"],
                                      n_threads=NTHREADS)
            rv, rq, rP, rll = ref["vec"], ref["quat"], ref["cov"], ref["loglik"]
            b.synchronize()
        gv, gq, gP, gll, ut = b.get_state()
    assert ut == T * 1000
    e = _assert_close((gv, gq, gP), (rv, rq, rP), TRAJ_TOL, "60 s")
    assert _rel_ll(gll, rll) < TRAJ_TOL
    print("60 s parity:", e, "loglik", _rel_ll(gll, rll))
    # the filter tracked the truth
    tv, tq = synth.truth_state_at(truth, T - 1)
    assert np.max(np.abs(gv[9:12] - tv[9:12, None])) < 0.5


# ------------------------------------------------------------------------------------------------
# delayed measurements: history rewind (config 5)
# ------------------------------------------------------------------------------------------------
def test_delayed_pose_fix_rewind_matches_oracle_history(oracle):
    """Pose fixes stamped at step k arrive 50 steps late.  Oracle: the reference's multimap insert +
    replay (mav_state_est.cpp:35-70).  GPU: snapshot at k, restore + update + replay in the op program."""
    N, T, LAT = 96, 400, 50
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
    ref = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st),
                              arrivals, n_threads=NTHREADS)
    from pronto_b200.schedule import program_from_arrivals

    n_slots = 3
    ops, cnt = program_from_arrivals(arrivals, snapshot_slots=n_slots, snapshot_period_us=100_000, snapshot_phase_us=1000)
    assert cnt["rewinds"] == len(pose) and cnt["discarded"] == 0
    with RBISBatch(N, snapshot_slots=n_slots) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(ops, imu=st["imu"], streams=gpu_streams(st))
        gv, gq, gP, gll, _ = b.get_state()
    _assert_close((gv, gq, gP), (ref["vec"], ref["quat"], ref["cov"]), STEP_TOL, "rewind")
    assert _rel_ll(gll, ref["loglik"]) < STEP_TOL
    n_applied = int(np.sum((ops["kind"] == capi.OP_IMU) | (ops["kind"] == capi.OP_MEAS)))
    assert n_applied * N == ref["calls"]  # the device re-applies exactly what the reference's replay re-applies


def test_two_arrival_schedules_as_two_handles(oracle):
    """Filters whose measurements arrive on different schedules live in different handles, one planner each (rbis_batch.h,
    "one schedule per handle"): sub-ensemble A gets its pose fixes 50 steps late, B 20 steps late and its leg odometry 3 steps
    late; the two handles are driven chunk by chunk, interleaved on the device, and each must match the oracle's per-filter
    multimap history fed with ITS arrivals."""
    from pronto_b200.schedule import program_from_arrivals

    N, T = 64, 300
    sc = scenario(N, T)
    st = sc["st"]
    ev = st["events"]

    def delayed(lat_pose, lat_lego):
        out, pend = [], []
        for e in ev:
            lat = 0 if e[0] == 0 else (lat_pose if e[1] == 1 else lat_lego)
            if e[0] == 0 or lat == 0:
                out.append(e)
            else:
                pend.append((e[3] + lat * 1000, e))
            if e[0] == 0:
                due = [q for q in pend if q[0] <= e[3]]
                pend = [q for q in pend if q[0] > e[3]]
                out += [q[1] for q in due]
        return out + [q[1] for q in pend]

    halves = {"A": (0, N // 2, delayed(50, 0)), "B": (N // 2, N, delayed(20, 3))}
    sub = lambda a, lo, hi: np.ascontiguousarray(a[..., lo:hi])
    handles, progs = {}, {}
    try:
        for name, (lo, hi, arr) in halves.items():
            ops, cnt = program_from_arrivals(arr, snapshot_slots=4, snapshot_period_us=25_000, snapshot_phase_us=1000)
            assert cnt["discarded"] == 0 and cnt["rewinds"] > 0
            b = RBISBatch(hi - lo, snapshot_slots=4)
            b.set_process_noise(*nominal_q())
            b.set_state(sub(sc["vec"], lo, hi), sub(sc["quat"], lo, hi), sub(sc["cov"], lo, hi))
            handles[name], progs[name] = b, ops
        # interleaved launches of the two programs, cut at arbitrary places (pieces of one handle run in order on its stream)
        cuts = {k: np.linspace(0, len(v), 5).astype(int) for k, v in progs.items()}
        for i in range(4):
            for name, (lo, hi, _) in halves.items():
                piece = progs[name][cuts[name][i]:cuts[name][i + 1]]
                handles[name].run_fused(piece, imu=sub(st["imu"], lo, hi),
                                        streams=[MeasStream(synth.LEGODO_IDX, sub(st["legodo"], lo, hi), st["R_legodo"]),
                                                 MeasStream(synth.POSE_IDX, sub(st["pose_z"], lo, hi), st["R_pose"], quat=sub(st["pose_q"], lo, hi))])
        for name, (lo, hi, arr) in halves.items():
            gv, gq, gP, gll, _ = handles[name].get_state()
            ost = [dict(idx=synth.LEGODO_IDX, z=sub(st["legodo"], lo, hi), R=st["R_legodo"]),
                   dict(idx=synth.POSE_IDX, z=sub(st["pose_z"], lo, hi), R=st["R_pose"], quat=sub(st["pose_q"], lo, hi))]
            ref = oracle.run_ensemble(sub(sc["vec"], lo, hi), sub(sc["quat"], lo, hi), sub(sc["cov"], lo, hi), None, 0, nominal_q(),
                                      sub(st["imu"], lo, hi), ost, arr, n_threads=NTHREADS)
            _assert_close((gv, gq, gP), (ref["vec"], ref["quat"], ref["cov"]), STEP_TOL, f"schedule {name}")
            assert _rel_ll(gll, ref["loglik"]) < STEP_TOL
    finally:
        for b in handles.values():
            b.close()


# ------------------------------------------------------------------------------------------------
# statistics
# ------------------------------------------------------------------------------------------------
def test_stats_match_oracle_and_are_sharding_invariant(oracle):
    N, T, CH = 3000, 40, 256
    sc = scenario(N, T)
    st = sc["st"]
    tv, tq = synth.truth_state_at(sc["truth"], T - 1)

    def run(lo, hi):
        n = hi - lo
        sub = lambda a: np.ascontiguousarray(a[..., lo:hi])
        with RBISBatch(n) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sub(sc["vec"]), sub(sc["quat"]), sub(sc["cov"]))
            b.run_fused(st["events"], imu=sub(st["imu"]),
                        streams=[MeasStream(synth.LEGODO_IDX, sub(st["legodo"]), st["R_legodo"]),
                                 MeasStream(synth.POSE_IDX, sub(st["pose_z"]), st["R_pose"], quat=sub(st["pose_q"]))])
            chunks, pf = b.stats(tv, tq, chunk=CH, want_per_filter=True)
            state = b.get_state()
        return chunks, pf, state

    chunks, pf, state = run(0, N)
    assert chunks.shape == ((N + CH - 1) // CH, 96)
    total = reduce_chunks(chunks)
    assert total[46] == N and total[45] == 0
    # per-filter error / NEES against the oracle's definition (noise_id.cpp:37-38)
    gv, gq, gP = state[0], state[1], state[2]
    for n in range(0, N, 97):
        e = oracle.state_error(gv[:, n], gq[:, n], tv, tq)
        assert np.max(np.abs(pf[:21, n] - e)) < 1e-13
        P = gP[:, n].reshape(21, 21).T[3:12, 3:12]
        nees = e[3:12] @ np.linalg.solve(P, e[3:12])
        assert abs(pf[21, n] - nees) < 1e-9 * max(1.0, nees)
    assert np.array_equal(pf[22], state[3])
    assert 4.0 < total[42] / N < 16.0  # mean NEES plausible
    # host replay of the fixed reduction tree gives the same bits
    def tree(v):
        v = np.concatenate([v, np.zeros(CH - len(v))])
        s = CH // 2
        while s > 0:
            v = v[:s] + v[s:2 * s]
            s //= 2
        return v[0]

    for c in (0, chunks.shape[0] - 1):
        sl = slice(c * CH, min(N, (c + 1) * CH))
        assert tree(pf[3, sl]) == chunks[c, 3]
        assert tree(pf[21, sl]) == chunks[c, 42]
        assert tree(pf[5, sl] * pf[5, sl]) == chunks[c, 21 + 5]
    # shard the ensemble 2 and 4 ways on chunk boundaries: identical chunk partials => identical totals
    for G in (2, 4):
        per = ((N + CH - 1) // CH + G - 1) // G * CH
        parts = [run(g * per, min(N, (g + 1) * per))[0] for g in range(G) if g * per < N]
        merged = np.concatenate(parts)
        assert np.array_equal(merged, chunks)
        assert np.array_equal(reduce_chunks(merged), total)


def test_nonfinite_filters_are_counted_not_fatal():
    N = 64
    vec, quat, cov = random_ensemble(N, seed=3)
    vec[4, 10] = np.nan
    cov[0, 20] = np.inf
    with RBISBatch(N) as b:
        b.set_state(vec, quat, cov)
        b.set_process_noise(*nominal_q())
        g = np.zeros((3, N)); a = np.zeros((3, N)); a[2] = 9.8
        b.ins_step(g, a, 1e-3)
        chunks, _ = b.stats(np.zeros(21), np.array([1.0, 0, 0, 0]), chunk=64)
    tot = reduce_chunks(chunks)
    assert tot[46] == N and tot[45] >= 1 and np.isfinite(tot[:45]).all()


# ------------------------------------------------------------------------------------------------
# API behaviour
# ------------------------------------------------------------------------------------------------
def test_state_roundtrip_and_filter_access():
    N = 200
    vec, quat, cov = random_ensemble(N, seed=11)
    ll = np.arange(N) * 0.5
    with RBISBatch(N) as b:
        b.set_state(vec, quat, cov, loglik=ll, utime=123)
        gv, gq, gP, gll, ut = b.get_state()
        assert ut == 123 and np.array_equal(gv, vec) and np.array_equal(gq, quat) and np.array_equal(gll, ll)
        # cov comes back symmetrised from its upper triangle
        full = cov.reshape(21, 21, N)  # [c][r][N]
        up = np.where((np.arange(21)[:, None] >= np.arange(21)[None, :])[:, :, None], full, full.transpose(1, 0, 2))
        assert np.array_equal(gP.reshape(21, 21, N), up)
        v, q, P, l = b.get_filter(57)
        assert np.array_equal(v, vec[:, 57]) and np.array_equal(q, quat[:, 57]) and l == ll[57]
        assert np.array_equal(P, gP[:, 57])
        b.set_filter(3, v, q, P, 9.0)
        v3, q3, P3, l3 = b.get_filter(3)
        assert np.array_equal(v3, v) and np.array_equal(P3, P) and l3 == 9.0
        assert b.launch_count > 0


def test_errors():
    N = 8
    with RBISBatch(N, snapshot_slots=2) as b:
        imu = np.zeros((2, 6, N))
        with pytest.raises(capi.RBISError, match="out of range"):
            b.run_fused([(capi.OP_IMU, 0, 5, 0, 1e-3)], imu=imu)
        with pytest.raises(capi.RBISError, match="dt must be positive"):
            b.run_fused([(capi.OP_IMU, 0, 0, 0, 0.0)], imu=imu)
        with pytest.raises(capi.RBISError, match="imu is NULL"):
            b.run_fused([(capi.OP_IMU, 0, 0, 0, 1e-3)])
        with pytest.raises(capi.RBISError, match="stream"):
            b.run_fused([(capi.OP_MEAS, 0, 0, 0, 0.0)])
        with pytest.raises(capi.RBISError, match="empty snapshot"):
            b.run_fused([(capi.OP_RESTORE, 0, 1, 0, 0.0)])
        with pytest.raises(capi.RBISError, match="slot"):
            b.run_fused([(capi.OP_SNAPSHOT, 0, 2, 0, 0.0)])
        with pytest.raises(capi.RBISError, match="index"):
            b.indexed_update([21], np.zeros((1, N)), np.eye(1))
        with pytest.raises(capi.RBISError):
            b.indexed_update(list(range(10)), np.zeros((10, N)), np.eye(10))
        b.run_fused([])  # empty program is a no-op
        b.run_fused([(capi.OP_SNAPSHOT, 0, 1, 0, 0.0), (capi.OP_RESTORE, 0, 1, 5, 0.0)])
        assert b.get_state(cov=False)[4] == 5


# ------------------------------------------------------------------------------------------------
# full-size property checks (BASELINE.json sizes; the oracle only sees a slice)
# ------------------------------------------------------------------------------------------------
def test_full_size_ensemble_replication_property(oracle):
    """65,536 filters = 256 distinct filters tiled 256 times: every replica must produce the same bits
    wherever it sits in the grid, and the distinct 256 must match the oracle."""
    N, D, T = 65_536, 256, 100
    sc = scenario(D, T)
    st = sc["st"]
    tile = lambda a: np.ascontiguousarray(np.tile(a, (1,) * (a.ndim - 1) + (N // D,)))
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(tile(sc["vec"]), tile(sc["quat"]), tile(sc["cov"]))
        b.run_fused(st["events"], imu=tile(st["imu"]),
                    streams=[MeasStream(synth.LEGODO_IDX, tile(st["legodo"]), st["R_legodo"]),
                             MeasStream(synth.POSE_IDX, tile(st["pose_z"]), st["R_pose"], quat=tile(st["pose_q"]))])
        gv, gq, gP, gll, _ = b.get_state()
    for a in (gv, gq, gP, gll[None]):
        r = a.reshape(a.shape[0], N // D, D)
        assert np.array_equal(r, np.broadcast_to(r[:, :1], r.shape))
    ref = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st),
                              st["events"], n_threads=NTHREADS)
    _assert_close((gv[:, :D], gq[:, :D], gP[:, :D]), (ref["vec"], ref["quat"], ref["cov"]), STEP_TOL, "65536")
    # covariance stays symmetric positive definite on the measured block
    P = gP[:, 5].reshape(21, 21)
    assert np.array_equal(P, P.T) and np.all(np.linalg.eigvalsh(P[3:12, 3:12]) > 0)


# ------------------------------------------------------------------------------------------------
# launch groups (CTA ranges on separate streams, consecutive launches overlap)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_launch_groups_do_not_change_results_and_keep_call_order(oracle):
    from pronto_b200.batch import make_ops

    """A >148-CTA ensemble run as 3 consecutive fused launches with launch_groups = 1 / automatic / 8 gives
    bit-identical states, also when reads (get_state, stats) and writes (set_process_noise) sit between the
    launches -- the main stream must join the group streams and the next launch must wait for it."""
    N, T = 148 * 256 + 700, 12
    sc = scenario(64, 3 * T)
    st = sc["st"]
    rep = lambda a: np.ascontiguousarray(np.tile(a, (1,) * (a.ndim - 1) + (N // 64 + 1,))[..., :N])
    vec, quat, cov = rep(sc["vec"]), rep(sc["quat"]), rep(sc["cov"])
    imu, lego, pz, pq = rep(st["imu"]), rep(st["legodo"]), rep(st["pose_z"]), rep(st["pose_q"])
    ev = st["events"]
    cuts = [0, len(ev) // 3, 2 * len(ev) // 3, len(ev)]
    outs = []
    for groups, piece in ((1, 0), (0, 0), (8, 0), (6, 8)):   # the last: every call cut into pieces of ~8 ops (host inputs staged once)
        with RBISBatch(N, launch_groups=groups, piece_ops=piece) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(vec, quat, cov)
            mids = []
            for k in range(3):
                ops = make_ops(ev[cuts[k]:cuts[k + 1]])
                b.run_fused(ops, imu=imu, streams=[MeasStream(synth.LEGODO_IDX, lego, st["R_legodo"]),
                                                   MeasStream(synth.POSE_IDX, pz, st["R_pose"], quat=pq)])
                if k == 0:
                    mids.append(b.get_state(cov=False)[0].copy())      # read between launches
                if k == 1:
                    q = nominal_q()
                    b.set_process_noise(q[0] * 1.5, q[1], q[2], q[3])  # write between launches
            outs.append((b.get_state(), mids[0]))
    (g1, m1), (g0, m0), (g8, m8), (gp, mp) = outs
    for a, c in ((g1, g0), (g1, g8), (g1, gp)):
        for x, y in zip(a[:4], c[:4]):
            assert np.array_equal(x, y)
    assert np.array_equal(m1, m0) and np.array_equal(m1, m8) and np.array_equal(m1, mp)
    # and the replicated filters agree with the oracle on the first 64
    ref = oracle.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), st["imu"], oracle_streams(st),
                              ev[:cuts[2]], n_threads=NTHREADS)
    q = nominal_q()
    ref = oracle.run_ensemble(ref["vec"], ref["quat"], ref["cov"], ref["loglik"], ev[cuts[2] - 1][3], (q[0] * 1.5, q[1], q[2], q[3]),
                              st["imu"], oracle_streams(st), ev[cuts[2]:], n_threads=NTHREADS)
    _assert_close((g1[0][:, :64], g1[1][:, :64], g1[2][:, :64]), (ref["vec"], ref["quat"], ref["cov"]), STEP_TOL, "groups")


# ------------------------------------------------------------------------------------------------
# shared input columns (parameter sweeps over the same data, BASELINE config 4)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_column_maps_equal_expanded_inputs_and_oracle(oracle):
    """N = 12 parameter points x 16 noise realisations: every filter reads one of 16 shared input columns
    (rbis_batch_set_column_map) and has its own process noise and per-filter diagonal leg-odometry R.  Must be
    bit-identical to the same run with the columns expanded to per-filter arrays, and match the oracle."""
    from pronto_b200.batch import make_ops

    C_, G_, T = 16, 12, 60
    N = C_ * G_
    sc = scenario(C_, T)
    st = sc["st"]
    cmap = (np.arange(N) % C_).astype(np.int32)
    grid = np.repeat(np.linspace(0.5, 2.0, G_), C_)
    q = nominal_q()
    qs = (q[0] * grid, q[1] * grid[::-1].copy(), np.full(N, q[2]), np.full(N, q[3]))
    r_lego = np.ascontiguousarray(np.tile((synth.NOMINAL["r_vxyz"] ** 2) * grid, (3, 1)))
    ex = lambda a: np.ascontiguousarray(a[..., cmap])
    vec, quat, cov = ex(sc["vec"]), ex(sc["quat"]), ex(sc["cov"])
    ops = make_ops(st["events"])
    outs = []
    for mapped in (True, False):
        with RBISBatch(N) as b:
            b.set_process_noise(*[np.ascontiguousarray(a) for a in qs])
            b.set_state(vec, quat, cov)
            if mapped:
                b.set_column_map(-1, cmap, C_); b.set_column_map(0, cmap, C_); b.set_column_map(1, cmap, C_)
                imu, lego, pz, pq = st["imu"], st["legodo"], st["pose_z"], st["pose_q"]
            else:
                imu, lego, pz, pq = ex(st["imu"]), ex(st["legodo"]), ex(st["pose_z"]), ex(st["pose_q"])
            streams = [MeasStream(synth.LEGODO_IDX, lego, r_lego, per_filter_diag=True),
                       MeasStream(synth.POSE_IDX, pz, st["R_pose"], quat=pq)]
            b.run_fused(ops, imu=imu, streams=streams)
            if mapped:  # the one-op entry points ignore the maps ([3][N] inputs)
                b.ins_step(ex(st["imu"][0, 0:3]), ex(st["imu"][0, 3:6]), 1e-3, utime=10 ** 9)
                b.set_column_map(-1, None)
                with pytest.raises(ValueError):
                    b.run_fused(ops[:1], imu=st["imu"])  # identity again: [rows][6][16] is the wrong shape
            else:
                b.ins_step(ex(st["imu"][0, 0:3]), ex(st["imu"][0, 3:6]), 1e-3, utime=10 ** 9)
            outs.append(b.get_state())
    for x, y in zip(outs[0][:4], outs[1][:4]):
        assert np.array_equal(x, y)
    ev = list(st["events"]) + [(0, 0, 0, 10 ** 9, 1e-3)]
    ref = oracle.run_ensemble(vec, quat, cov, None, 0, qs, ex(st["imu"]),
                              [dict(idx=synth.LEGODO_IDX, z=ex(st["legodo"]), R=r_lego, per_filter_diag=True),
                               dict(idx=synth.POSE_IDX, z=ex(st["pose_z"]), R=st["R_pose"], quat=ex(st["pose_q"]))],
                              ev, n_threads=NTHREADS)
    _assert_close(outs[0][:3], (ref["vec"], ref["quat"], ref["cov"]), STEP_TOL, "column maps")
    assert _rel_ll(outs[0][3], ref["loglik"]) < STEP_TOL


# ------------------------------------------------------------------------------------------------
# against fixtures produced by the REFERENCE's own code (tests/golden/rbis_reference_golden.npz, made by
# tests/golden/make_reference_golden.py from oracle/_ref = the reference's rbis.cpp, rbis_update_interface.cpp,
# update_history.cpp, mav_state_est.cpp compiled unmodified against stand-in dependency headers)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rbis_reference_golden.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,tumbling", [("walk", False), ("tumble", True)])
def test_fused_and_rewind_match_reference_golden(ref_golden, name, tumbling):
    from pronto_b200.schedule import program_from_arrivals

    g = ref_golden
    sc = scenario(4, 400, tumbling=tumbling)
    st = sc["st"]
    with RBISBatch(4, snapshot_slots=3) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(st["events"], imu=st["imu"], streams=gpu_streams(st))
        gv, gq, gP, gll, _ = b.get_state()
        _assert_close((gv, gq, gP), (g[f"{name}_vec"], g[f"{name}_quat"], g[f"{name}_cov"]), STEP_TOL, name)
        assert _rel_ll(gll, g[f"{name}_loglik"]) < STEP_TOL
        # pose fixes 50 steps late: the reference's multimap insert + replay vs snapshot / restore / replay on the device
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
        ops, cnt = program_from_arrivals(arrivals, snapshot_slots=3, snapshot_period_us=100_000, snapshot_phase_us=1000)
        assert cnt["rewinds"] == len(pose)
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(ops, imu=st["imu"], streams=gpu_streams(st))
        gv, gq, gP, gll, _ = b.get_state()
        _assert_close((gv, gq, gP), (g[f"{name}_late_vec"], g[f"{name}_late_quat"], g[f"{name}_late_cov"]), STEP_TOL, name + " late")
        assert _rel_ll(gll, g[f"{name}_late_loglik"]) < STEP_TOL


@pytest.mark.gpu
def test_single_updates_match_reference_golden(ref_golden):
    g = ref_golden
    N = 5
    rep = lambda a: np.repeat(np.asarray(a)[:, None], N, axis=1)
    vec, quat, cov = rep(g["op_vec"]), rep(g["op_quat"]), rep(g["op_cov"].T.reshape(-1))
    with RBISBatch(N) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(vec, quat, cov)
        b.ins_step(rep(g["op_gyro"]), rep(g["op_accel"]), 1e-3)
        gv, gq, gP, _, _ = b.get_state()
        assert np.max(np.abs(gv - g["op_ins_vec"][:, None])) < 1e-13 and np.max(np.abs(gq - g["op_ins_quat"][:, None])) < 1e-13
        assert np.max(np.abs(gP - g["op_ins_cov"].T.reshape(-1)[:, None])) < 1e-14
        for c in range(int(g["n_meas_cases"])):
            b.set_state(vec, quat, cov)
            mq = g[f"m{c}_mq"]
            b.indexed_update([int(i) for i in g[f"m{c}_idx"]], rep(g[f"m{c}_z"]), g[f"m{c}_R"], quat=rep(mq) if mq.size else None)
            gv, gq, gP, gll, _ = b.get_state()
            assert np.max(np.abs(gv - g[f"m{c}_vec"][:, None])) < 1e-11, c
            assert np.max(np.abs(gq - g[f"m{c}_quat"][:, None])) < 1e-11, c
            assert np.max(np.abs(gP - g[f"m{c}_cov"].T.reshape(-1)[:, None])) < 1e-12, c
            assert np.max(np.abs(gll - float(g[f"m{c}_ll"]))) < 1e-9 * max(1.0, abs(float(g[f"m{c}_ll"]))), c


def test_snapshot_statistics_on_the_side_stream_equal_live_statistics():
    """rbis_batch_stats_snapshot_enqueue: the statistics of the ensemble as the program's final SNAPSHOT left it, computed on a side
    stream while later launches already run -- same bits as rbis_batch_stats on the live state at that point, for launch groups
    on and off, and a program that re-snapshots into the slot waits for the reader."""
    import torch

    N, T, CH = 3000, 60, 256
    sc = scenario(N, T)
    st = sc["st"]
    ev = list(st["events"])
    half = len(ev) // 2
    tv, tq = synth.truth_state_at(sc["truth"], T - 1)
    n_ch = (N + CH - 1) // CH
    for groups in (1, 4):
        with RBISBatch(N, snapshot_slots=2, launch_groups=groups, mapping=1) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            b.run_fused(ev[:half], imu=st["imu"], streams=gpu_streams(st))
            want_a, _ = b.stats(tv, tq, chunk=CH)
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            outs = [torch.empty((n_ch, capi.NUM_STATS), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
            b.run_fused(ev[:half] + [(capi.OP_SNAPSHOT, 0, 0, ev[half - 1][3], 0.0)], imu=st["imu"], streams=gpu_streams(st))
            t0 = b.stats_snapshot_enqueue(0, tv, tq, outs[0], chunk=CH)
            # the ensemble moves on at once; the second half ends with a snapshot into slot 1, a third launch re-snapshots slot 0
            b.run_fused(ev[half:] + [(capi.OP_SNAPSHOT, 0, 1, ev[-1][3], 0.0)], imu=st["imu"], streams=gpu_streams(st))
            t1 = b.stats_snapshot_enqueue(1, tv, tq, outs[1], chunk=CH)
            b.run_fused([(capi.OP_SNAPSHOT, 0, 0, ev[-1][3], 0.0)], imu=st["imu"], streams=gpu_streams(st))
            t2 = b.stats_snapshot_enqueue(0, tv, tq, outs[2], chunk=CH)
            for t in (t0, t1, t2):
                b.wait(t)
            want_b, _ = b.stats(tv, tq, chunk=CH)
        assert np.array_equal(outs[0], want_a), groups
        assert np.array_equal(outs[1], want_b) and np.array_equal(outs[2], want_b), groups
        assert not np.array_equal(want_a, want_b)
    # argument checking: slots out of range / never written, a handle without a ring
    from pronto_b200.capi import RBISError

    with RBISBatch(64, snapshot_slots=2) as b:
        out = torch.empty((1, capi.NUM_STATS), dtype=torch.float64, pin_memory=True).numpy()
        for bad in (-1, 2, 1):   # 1 is in range but empty
            with pytest.raises(RBISError):
                b.stats_snapshot_enqueue(bad, tv, tq, out, chunk=64)
    with RBISBatch(64) as b:
        with pytest.raises(RBISError):
            b.stats_snapshot_enqueue(0, tv, tq, out, chunk=64)


@pytest.mark.parametrize("where", ["host", "device"])
def test_float32_rows_are_widened_exactly(where):
    """RBIS_MEM_F32_ROWS: sensor rows given as float arrays (host: half the PCIe bytes) give the same bits as the same values given as
    doubles, with and without launch groups."""
    import torch

    N, T = 1500, 40
    sc = scenario(N, T)
    st = sc["st"]
    f32 = {k: np.ascontiguousarray(st[k].astype(np.float32)) for k in ("imu", "legodo", "pose_z", "pose_q")}
    f64 = {k: np.ascontiguousarray(v.astype(np.float64)) for k, v in f32.items()}
    put = (lambda a: torch.from_numpy(a).cuda()) if where == "device" else (lambda a: a)
    out = []
    for rows, groups in ((f64, 1), (f32, 1), (f32, 3)):
        with RBISBatch(N, launch_groups=groups, mapping=1) as b:
            b.set_process_noise(*nominal_q())
            b.set_state(sc["vec"], sc["quat"], sc["cov"])
            r = {k: put(v) for k, v in rows.items()}
            ev = st["events"]
            h = len(ev) // 2
            for part in (ev[:h], ev[h:]):   # two calls: the float scratch of the staging slots is reused
                b.run_fused(part, imu=r["imu"], streams=[MeasStream(synth.LEGODO_IDX, r["legodo"], st["R_legodo"]),
                                                         MeasStream(synth.POSE_IDX, r["pose_z"], st["R_pose"], quat=r["pose_q"])])
            out.append(b.get_state())
    for o in out[1:]:
        for x, y in zip(out[0][:4], o[:4]):
            assert np.array_equal(x, y)
    assert np.isfinite(out[0][0]).all()
