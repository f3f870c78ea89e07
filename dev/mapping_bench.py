#!/usr/bin/env python
"""dev/mapping_bench.py -- small-ensemble timings of the three mappings through the C ABI: device-resident inputs, back-to-back launches."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pronto_b200 import MeasStream, RBISBatch, synth
from pronto_b200.batch import make_ops

dev = torch.device("cuda", 0)
p = synth.NOMINAL
R_lego = np.eye(3) * p["r_vxyz"] ** 2
R_pose = np.diag([p["r_xyz"] ** 2] * 3 + [p["r_chi"] ** 2] * 3)
Tc, K = 200, 8
truth = synth.truth_trajectory(K * Tc)
cases = [(int(a), b) for a, b in (x.split(":") for x in (sys.argv[1:] or ["4096:imu", "8192:cfg3", "2048:cfg3", "16384:cfg3"]))]
for N, kind in cases:
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    vec0, quat0, cov0 = bench.initial_state(N, gen, dev)
    chunks = [bench.device_chunk(truth, c * Tc, Tc, N, gen, dev) for c in range(K)]
    ref = None
    for mapping in [int(m) for m in os.environ.get("MAPPINGS", "1,4,8").split(",")]:
        with RBISBatch(N, mapping=mapping) as b:
            b.set_process_noise(p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
            b.set_state(vec0, quat0, cov0)
            stream = torch.cuda.ExternalStream(b.cuda_stream, device=dev)
            if kind == "imu":
                progs = [make_ops([e for e in bench.chunk_events(Tc, c * Tc)[0] if e[0] == 0]) for c in range(K)]
                preps = [b.prepare_fused(progs[c], imu=chunks[c]["imu"], streams=[]) for c in range(K)]
            else:
                progs = [make_ops(bench.chunk_events(Tc, c * Tc)[0]) for c in range(K)]
                preps = [b.prepare_fused(progs[c], imu=chunks[c]["imu"], streams=[MeasStream(synth.LEGODO_IDX, chunks[c]["legodo"], R_lego),
                                                                                 MeasStream(synth.POSE_IDX, chunks[c]["pose_z"], R_pose, quat=chunks[c]["pose_q"])]) for c in range(K)]
            best = 1e9
            for rep in range(3):
                b.set_state(vec0, quat0, cov0); b.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for c in range(K):
                    b.run_prepared(preps[c])
                b.record(); e1.record(stream)
                b.synchronize(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
            st = b.get_state()
            same = "" if ref is None else ("same bits" if all(np.array_equal(x, y) for x, y in zip(st[:4], ref[:4])) else "DIFFERENT")
            if ref is None:
                ref = st
            print(f"N={N} {kind} mapping={mapping}: {best:.3f} ms per {Tc}-step launch = {N * Tc / best / 1e6:.3f} G filter-steps/s  variant={b.last_kernel_variant} {same}", flush=True)
