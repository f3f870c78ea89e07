#!/usr/bin/env python
"""dev/smoother_bench.py -- EKF smoother over an ensemble: forward pass that snapshots every update's posterior, then the
backward pass (rbis_batch_smooth_backward); the CPU oracle's backward pass timed beside it on a small sample."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from pronto_b200 import MeasStream, RBISBatch, capi, smoother, synth
from oracle import oracle_api

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
    gv, gq, gP, _ = b.get_snapshot(int(alias[1]))
# CPU oracle: forward + backward on a few filters, all threads
nf = 2 * (os.cpu_count() or 1)
sub = lambda a: np.ascontiguousarray(a[..., :nf].cpu().numpy())
st = [dict(idx=synth.LEGODO_IDX, z=sub(ch["legodo"]), R=R_lego), dict(idx=synth.POSE_IDX, z=sub(ch["pose_z"]), R=R_pose, quat=sub(ch["pose_q"]))]
oracle_api.load(); oracle_api.set_constants()
q = (p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
t0 = time.perf_counter()
oracle_api.run_ensemble(sub(vec0), sub(quat0), sub(cov0), None, 0, q, sub(ch["imu"]), st, ev, n_threads=os.cpu_count())
t1 = time.perf_counter()
ref = oracle_api.smooth_ensemble(sub(vec0), sub(quat0), sub(cov0), 0, q, sub(ch["imu"]), st, ev, 1e-3, n_threads=os.cpu_count())
t2 = time.perf_counter()
back = (t2 - t1) - (t1 - t0)
print(f"  CPU oracle ({os.cpu_count()} threads, {nf} filters): backward pass ~{1e3 * back:.1f} ms = {nf * len(steps) / max(back, 1e-9) / 1e6:.3f} M smoothing steps/s")
print("  parity of update 1 (first IMU step) on the sample: vec %.2e cov %.2e" % (
    np.max(np.abs(gv[:, :nf] - ref["post_vec"][0])), np.max(np.abs(gP[:, :nf] - ref["post_cov"][0]))))
