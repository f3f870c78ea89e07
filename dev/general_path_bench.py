#!/usr/bin/env python
"""dev/general_path_bench.py -- cost of measurement chunks that are not aligned index triples (they take meas_general in the
dense kernel): 65,536 filters x 200 IMU steps + leg odometry every 2nd step, with and without a scalar yaw update
(m = 1, index 8, rbis_yawlock-style) every 10th step."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pronto_b200 import MeasStream, RBISBatch, capi, synth
from pronto_b200.batch import make_ops

N, T = 65536, 200
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
truth = synth.truth_trajectory(T)
vec0, quat0, cov0 = bench.initial_state(N, gen, dev)
ch = bench.device_chunk(truth, 0, T, N, gen, dev)
p = synth.NOMINAL
R_lego = np.eye(3) * p["r_vxyz"] ** 2
yaw_z = torch.zeros((T // 10, 1, N), dtype=torch.float64, device=dev)
ql_z = torch.zeros((T // 10, 4, N), dtype=torch.float64, device=dev)
Bq = np.random.default_rng(3).normal(size=(4, 4))
R_ql = Bq @ Bq.T * 1e-3 + np.eye(4) * 1e-2   # correlated 4x4: quick-lock style [8, 9, 10, 11] (quick_lock.cpp:132) -> meas_general
for label, with_yaw in (("IMU + leg odometry", False), ("IMU + leg odometry + scalar yaw update every 10th step", True),
                        ("IMU + leg odometry + correlated 4-row update [8,9,10,11] every 10th step (general path)", 2)):
    ev, li, yi = [], 0, 0
    for k in range(T):
        ut = (k + 1) * 1000
        ev.append((capi.OP_IMU, 0, k, ut, 1e-3))
        if k % 2 == 0: ev.append((capi.OP_MEAS, 0, li, ut, 0.0)); li += 1
        if with_yaw and k % 10 == 0: ev.append((capi.OP_MEAS, 1, yi, ut, 0.0)); yi += 1
    ops = make_ops(ev)
    streams = [MeasStream(synth.LEGODO_IDX, ch["legodo"], R_lego)]
    if with_yaw == 2:
        streams.append(MeasStream([8, 9, 10, 11], ql_z, R_ql))
    elif with_yaw:
        streams.append(MeasStream([8], yaw_z, np.array([[0.01]])))
    with RBISBatch(N) as b:
        b.set_process_noise(p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
        best = 1e9
        for rep in range(3):
            b.set_state(vec0, quat0, cov0); b.synchronize()
            for _ in range(2):
                b.run_fused(ops, imu=ch["imu"], streams=streams)
            b.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                b.run_fused(ops, imu=ch["imu"], streams=streams)
            b.synchronize()
            best = min(best, (time.perf_counter() - t0) / 5)
        print(f"{label}: kernel variant {b.last_kernel_variant}, {best * 1e3:.2f} ms per 200-step launch, {N * T / best / 1e9:.3f} G filter-steps/s")
