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
