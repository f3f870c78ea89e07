#!/usr/bin/env python
"""dev/config5_bench.py -- BASELINE configs[4]: 65,536 filters, pose fixes delivered 50 steps late (history rewind +
replay inside the fused program).  Prints filter-steps/s counting ONLY trajectory steps (replayed steps are overhead)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from pronto_b200 import MeasStream, RBISBatch, capi, synth
from pronto_b200.batch import make_ops
from pronto_b200.schedule import program_from_arrivals

N, Tc, LAT, CH = 65536, 200, 50, 6
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
truth = synth.truth_trajectory(CH * Tc)
vec0, quat0, cov0 = bench.initial_state(N, gen, dev)
p = synth.NOMINAL
R_lego = np.eye(3) * p["r_vxyz"] ** 2
R_pose = np.diag([p["r_xyz"] ** 2] * 3 + [p["r_chi"] ** 2] * 3)
chunks = [bench.device_chunk(truth, c * Tc, Tc, N, gen, dev) for c in range(CH)]
# one program for the whole run with GLOBAL rows; inputs concatenated
cat = {k: torch.cat([c[k] for c in chunks]).contiguous() for k in chunks[0]}
ev, li, pi = [], 0, 0
for k in range(CH * Tc):
    ut = (k + 1) * 1000
    ev.append((capi.OP_IMU, 0, k, ut, 1e-3))
    if k % 2 == 0: ev.append((capi.OP_MEAS, 0, li, ut, 0.0)); li += 1
    if k % 100 == 0: ev.append((capi.OP_MEAS, 1, pi, ut, 0.0)); pi += 1
pose = [e for e in ev if e[0] == 1 and e[1] == 1]
arr, pend = [], list(pose)
for e in ev:
    if e[0] == 1 and e[1] == 1: continue
    arr.append(e)
    while pend and e[0] == 0 and e[3] >= pend[0][3] + LAT * 1000: arr.append(pend.pop(0))
arr += pend
for label, arrivals in (("in-order (config 3)", ev), ("50-step-late pose fixes (config 5)", arr)):
    ops, cnt = program_from_arrivals(arrivals, snapshot_slots=3, snapshot_period_us=100_000, snapshot_phase_us=1000)
    with RBISBatch(N, snapshot_slots=3) as b:
        b.set_process_noise(p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
        streams = [MeasStream(synth.LEGODO_IDX, cat["legodo"], R_lego), MeasStream(synth.POSE_IDX, cat["pose_z"], R_pose, quat=cat["pose_q"])]
        best = 1e9
        for rep in range(3):
            b.set_state(vec0, quat0, cov0); b.synchronize()
            t0 = time.perf_counter()
            b.run_fused(ops, imu=cat["imu"], streams=streams)
            b.synchronize()
            best = min(best, time.perf_counter() - t0)
        n_applied = int(np.sum((ops["kind"] == capi.OP_IMU) | (ops["kind"] == capi.OP_MEAS)))
        print(f"{label}: {len(ops)} ops ({n_applied} updates applied, {cnt}), {best * 1e3:.2f} ms, {N * CH * Tc / best / 1e9:.3f} G trajectory filter-steps/s")
