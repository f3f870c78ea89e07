#!/usr/bin/env python
"""dev/smoother_bench.py -- EKF smoother over an ensemble: forward pass that snapshots every update's posterior, then the
backward pass (rbis_batch_smooth_backward).  (The CPU oracle's backward pass was timed beside it once, profiles/r1_smoother_bench.txt;
only tests/, smoke() and bench.py may call into oracle/, so that leg is not kept here.)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from pronto_b200 import MeasStream, RBISBatch, capi, smoother, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
truth = synth.truth_trajectory(T)
vec0, quat0, cov0 = bench.initial_state(N, gen, dev)
ch = bench.device_chunk(truth, 0, T, N, gen, dev)
ev, _, _ = bench.chunk_events(T, 0)
p = synth.NOMINAL
R_lego = np.eye(3) * p["r_vxyz"] ** 2
R_pose = np.diag([p["r_xyz"] ** 2] * 3 + [p["r_chi"] ** 2] * 3)
ops, is_ins, slot = smoother.forward_program(ev)
np_slot, n_slot, steps, alias = smoother.plan(is_ins, slot)
print(f"N={N} T={T}: {len(ev)} updates, {len(is_ins)} ring slots = {len(is_ins) * 257 * 8 * N / 1e9:.2f} GB, {len(steps)} smoothing steps")
with RBISBatch(N, snapshot_slots=len(is_ins)) as b:
    b.set_process_noise(p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
    streams = [MeasStream(synth.LEGODO_IDX, ch["legodo"], R_lego), MeasStream(synth.POSE_IDX, ch["pose_z"], R_pose, quat=ch["pose_q"])]
    for rep in range(2):
        b.set_state(vec0, quat0, cov0); b.synchronize()
        t0 = time.perf_counter()
        b.run_fused(ops, imu=ch["imu"], streams=streams)
        b.synchronize()
        t1 = time.perf_counter()
        b.smooth_backward(np_slot, n_slot, steps, 1e-3)
        b.synchronize()
        t2 = time.perf_counter()
        print(f"  forward + snapshots {1e3 * (t1 - t0):.2f} ms ({N * T / (t1 - t0) / 1e9:.3f} G filter-steps/s), "
              f"backward {1e3 * (t2 - t1):.2f} ms = {N * len(steps) / (t2 - t1) / 1e6:.1f} M smoothing steps/s, "
              f"{N * len(steps) * 3 * 257 * 8 / (t2 - t1) / 1e9:.0f} GB/s of ring traffic (2 reads + 1 write of 257 doubles)")
