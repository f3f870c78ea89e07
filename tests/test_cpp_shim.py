"""include/rbis_batch.hpp (the C++ mirror of the reference's RBISUpdateInterface / MavStateEstimator API):
compiles with g++ against the C ABI (CPU), refuses to run without a GPU (CPU), and on a B200 replays a
delayed-arrival update list to the same head state as the CPU oracle's multimap history (gpu)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from common import nominal_q, oracle_streams, scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "shim_replay")


@pytest.fixture(scope="module")
def shim_exe(rbis_lib):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "shim_replay.cpp")
    libdir = os.path.join(ROOT, "pronto_b200", "lib")
    deps = [src, os.path.join(ROOT, "include", "rbis_batch.hpp"), os.path.join(ROOT, "include", "rbis_batch.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-o", EXE, src, f"-L{libdir}", "-lrbis_b200",
                               f"-Wl,-rpath,{libdir}"])
    return EXE


def delayed_arrivals(ev, latency_steps):
    pose = [e for e in ev if e[0] == 1 and e[1] == 1]
    arrivals, pending = [], list(pose)
    for e in ev:
        if e[0] == 1 and e[1] == 1:
            continue
        arrivals.append(e)
        while pending and e[0] == 0 and e[3] >= pending[0][3] + latency_steps * 1000:
            arrivals.append(pending.pop(0))
    return arrivals + pending


def write_case(path, sc, arrivals, span, slots, period):
    from pronto_b200 import synth

    st = sc["st"]
    N = sc["vec"].shape[1]
    q = nominal_q()
    with open(path, "wb") as f:
        f.write(struct.pack("<6q", N, 0, span, slots, period, len(arrivals)))
        for a in (sc["vec"], sc["quat"], sc["cov"]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        for kind, stream, row, utime, dt in arrivals:
            if kind == 0:
                f.write(struct.pack("<2q", 0, utime))
                f.write(struct.pack("<5d", dt, *q))
                f.write(np.ascontiguousarray(st["imu"][row]).tobytes())  # [6][N] = gyro then accel
            else:
                idx = synth.LEGODO_IDX if stream == 0 else synth.POSE_IDX
                R = np.asarray(st["R_legodo"] if stream == 0 else st["R_pose"], dtype=np.float64)
                z = st["legodo"][row] if stream == 0 else st["pose_z"][row]
                f.write(struct.pack("<2q", 1 if stream == 0 else 2, utime))
                f.write(struct.pack(f"<{1 + len(idx)}q", len(idx), *idx))
                f.write(np.ascontiguousarray(R.T).tobytes())  # column-major
                f.write(np.ascontiguousarray(z).tobytes())
                if stream == 1:
                    f.write(np.ascontiguousarray(st["pose_q"][row]).tobytes())


def test_shim_compiles_and_fails_loudly_without_a_gpu(shim_exe, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    sc = scenario(4, 4)
    p_in, p_out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_case(p_in, sc, sc["st"]["events"], 10_000_000, 2, 50_000)
    r = subprocess.run([shim_exe, p_in, p_out], capture_output=True, text=True)
    assert r.returncode == 5 and "no CPU path" in r.stderr  # RBIS_ERR_CUDA through MavStateEst::batch::Error
    assert not os.path.exists(p_out)


@pytest.mark.gpu
def test_shim_replay_matches_oracle_history(shim_exe, tmp_path):
    from oracle import oracle_api

    N, T, LAT = 80, 300, 50
    sc = scenario(N, T)
    arrivals = delayed_arrivals(sc["st"]["events"], LAT)
    p_in, p_out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_case(p_in, sc, arrivals, 10_000_000, 3, 100_000)
    r = subprocess.run([shim_exe, p_in, p_out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(p_out, dtype=np.float64)
    n0 = 21 * N + 4 * N + 441 * N + N
    gv, gq, gP, gll = np.split(raw[:n0], [21 * N, 25 * N, 466 * N])
    launches, dropped = np.frombuffer(raw[n0:n0 + 2].tobytes(), dtype=np.int64)
    ref = oracle_api.run_ensemble(sc["vec"], sc["quat"], sc["cov"], None, 0, nominal_q(), sc["st"]["imu"],
                                  oracle_streams(sc["st"]), arrivals, n_threads=os.cpu_count() or 1)
    from pronto_b200.parity import max_errors

    e = max_errors(gv.reshape(21, N), gq.reshape(4, N), gP.reshape(441, N), ref["vec"], ref["quat"], ref["cov"])
    assert e["vec"] < 1e-9 and e["quat"] < 1e-9 and e["cov"] < 1e-9, e
    assert np.max(np.abs(gll - ref["loglik"]) / np.maximum(1.0, np.abs(ref["loglik"]))) < 1e-9
    assert dropped == 0 and launches == len(arrivals)  # one fused launch per addUpdate(roll_forward=true)


def test_legodo_measurement_formation(rbis_lib, tmp_path):
    """MavStateEst::batch::LegOdoCommon against the formulas of rbis_legodo_common.cpp:35-170 /
    pronto_conversions_lcm.hpp:38-87 (velocity = delta / elapsed; index set and R by mode; uncertain R when
    odo_delta_status >= 0.5; fall back to lin_rate without a valid position).  CPU only: no filter math."""
    exe = os.path.join(BUILD, "legodo_check")
    os.makedirs(BUILD, exist_ok=True)
    libdir = os.path.join(ROOT, "pronto_b200", "lib")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-o", exe, os.path.join(ROOT, "tests", "cpp", "legodo_check.cpp"),
                           f"-L{libdir}", "-lrbis_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    rows = {}
    for line in out:
        tag, rest = line.split(" ", 1)
        f = dict(kv.split("=", 1) for kv in rest.split(" "))
        rows[tag] = dict(m=int(f["m"]), idx=[int(v) for v in f["idx"].split(",") if v], R=[float(v) for v in f["Rdiag"].split(",") if v],
                         off=float(f["offdiag"]), z=np.array([float(v) for v in f["z"].split(",") if v]), utime=int(f["utime"]),
                         sensor=int(f["sensor"]))
    dxyz = np.array([[0.002, 0.004], [-0.001, 0.003], [0.0005, 0.0015]])
    pos = np.array([[1.0, 2], [3, 4], [5, 6]])
    v2 = (dxyz / 0.002).reshape(-1)
    r = rows["lin_certain"]
    assert r["idx"] == [3, 4, 5] and r["R"] == [0.1 ** 2] * 3 and r["off"] == 0 and r["sensor"] == 11 and r["utime"] == 1002000
    assert np.allclose(r["z"], v2, rtol=1e-15)
    assert rows["lin_uncertain"]["R"] == [0.5 ** 2] * 3 and np.array_equal(rows["lin_uncertain"]["z"], r["z"])
    p = rows["posvel"]
    assert p["idx"] == [9, 10, 11, 3, 4, 5] and p["R"] == [0.01 ** 2] * 3 + [0.1 ** 2] * 3
    assert np.allclose(p["z"], np.concatenate([pos.reshape(-1), v2]), rtol=1e-15)
    assert rows["posvel_nopos"]["idx"] == [3, 4, 5] and np.array_equal(rows["posvel_nopos"]["z"], r["z"])
    lr = rows["linrot"]
    assert lr["idx"] == [3, 4, 5, 0, 1, 2] and lr["R"] == [0.5 ** 2] * 3 + [0.9 ** 2] * 3
    q = np.array([[0.99999, 0.99998], [0.002, -0.001], [0.001, 0.003], [0.004, 0.002]])
    rpy = np.stack([np.arctan2(2 * (q[0] * q[1] + q[2] * q[3]), 1 - 2 * (q[1] ** 2 + q[2] ** 2)),
                    np.arcsin(2 * (q[0] * q[2] - q[3] * q[1])),
                    np.arctan2(2 * (q[0] * q[3] + q[1] * q[2]), 1 - 2 * (q[2] ** 2 + q[3] ** 2))])
    assert np.allclose(lr["z"], np.concatenate([(dxyz / 0.004).reshape(-1), (rpy / 0.004).reshape(-1)]), rtol=1e-14)


def test_legodo_measurement_formation_matches_the_reference_compiled(rbis_lib):
    """The same class against the reference's own LegOdoCommon::createMeasurement (rbis_legodo_common.cpp:35-170 with
    pronto::getDeltaAsVelocity, pronto_conversions_lcm.hpp:38-87, and pronto_math.cpp), compiled unmodified into oracle/_ref:
    index set, measurement covariance, measurement vector, utime and sensor id for random poses, all three modes, certain /
    uncertain deltas, valid / invalid positions."""
    import ctypes as C

    from oracle import oracle_api

    path = oracle_api.build_ref()
    if path is None:
        pytest.skip("oracle/_ref/librbis_ref.so is not built and /root/reference is not mounted")
    ref = C.CDLL(path)
    if not hasattr(ref, "orc_legodo_create_measurement"):
        pytest.skip("oracle/_ref/librbis_ref.so predates the leg-odometry export")
    vp = C.c_void_p
    ref.orc_legodo_create_measurement.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_float, vp, vp, vp, vp, vp]
    ref.orc_legodo_create_measurement.restype = C.c_int
    exe = os.path.join(BUILD, "legodo_check")
    os.makedirs(BUILD, exist_ok=True)
    libdir = os.path.join(ROOT, "pronto_b200", "lib")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-o", exe, os.path.join(ROOT, "tests", "cpp", "legodo_check.cpp"),
                           f"-L{libdir}", "-lrbis_b200", f"-Wl,-rpath,{libdir}"])
    rng = np.random.default_rng(11)
    modes = {0: b"lin_rate", 2: b"lin_rot_rate", 3: b"pos_and_lin_rate"}
    worst = 0.0
    for case in range(60):
        mode = [0, 2, 3][case % 3]
        r = np.abs(rng.normal(size=5)) * 0.3 + 0.01
        pos = rng.normal(size=3)
        pq = rng.normal(size=4); pq /= np.linalg.norm(pq)
        dxyz = rng.normal(size=3) * 0.01
        ang = rng.normal(size=3) * (0.02 if case % 4 else 0.8)   # mostly small increments, some large rotations
        n = np.linalg.norm(ang)
        dq = np.concatenate([[np.cos(n / 2)], np.sin(n / 2) * ang / n])
        prev = 1_000_000 + int(rng.integers(0, 1000))
        ut = prev + int(rng.integers(500, 20000))
        pos_ok, delta_status = int(rng.integers(0, 2)), float(rng.choice([0.0, 0.2, 0.5, 0.7, 1.0]))
        m, ut_out = C.c_int(0), C.c_int64(0)
        idx, z, cov = np.zeros(6, dtype=np.int32), np.zeros(6), np.zeros(36)
        sensor = ref.orc_legodo_create_measurement(modes[mode], r.ctypes.data, pos.ctypes.data, pq.ctypes.data, dxyz.ctypes.data, dq.ctypes.data,
                                                   ut, prev, pos_ok, delta_status, C.byref(m), idx.ctypes.data, z.ctypes.data, cov.ctypes.data,
                                                   C.byref(ut_out))
        args = [str(mode)] + [repr(float(v)) for v in r] + [str(ut), str(prev), str(pos_ok), repr(delta_status)] + \
               [repr(float(v)) for v in np.concatenate([pos, dxyz, dq])]
        line = subprocess.run([exe] + args, capture_output=True, text=True, check=True).stdout.strip()
        f = dict(kv.split("=", 1) for kv in line.split(" ", 1)[1].split(" "))
        k = m.value
        assert int(f["m"]) == k and [int(v) for v in f["idx"].split(",") if v] == list(idx[:k]), (case, line)
        assert int(f["sensor"]) == sensor and int(f["utime"]) == ut_out.value == ut
        R_ref = cov[:k * k].reshape(k, k)
        assert np.array_equal(np.diag(R_ref), [float(v) for v in f["Rdiag"].split(",") if v]) and float(f["offdiag"]) == 0.0
        assert np.count_nonzero(R_ref - np.diag(np.diag(R_ref))) == 0
        zz = np.array([float(v) for v in f["z"].split(",") if v])
        err = np.max(np.abs(zz - z[:k]) / np.maximum(1.0, np.abs(z[:k])))
        worst = max(worst, err)
    assert worst <= 1e-12, worst
