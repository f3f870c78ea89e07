#!/usr/bin/env python
"""dev/synth_bench.py -- timings of the on-device stream synthesis: the materialising kernels alone (both generator modes) and
back-to-back rbis_batch_run_fused_synth calls against rbis_batch_run_fused over resident inputs."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pronto_b200 import MeasStream, RBISBatch, SynthSpec, synth
from pronto_b200.batch import make_ops

N, Tc, K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 200, 12
dev = torch.device("cuda", 0)
p = synth.NOMINAL
R_lego = np.eye(3) * p["r_vxyz"] ** 2
R_pose = np.diag([p["r_xyz"] ** 2] * 3 + [p["r_chi"] ** 2] * 3)
truth = synth.truth_trajectory(K * Tc)
gen = torch.Generator(device=dev); gen.manual_seed(1)
vec0, quat0, cov0 = bench.initial_state(N, gen, dev)
with RBISBatch(N) as b:
    b.set_process_noise(p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
    b.set_state(vec0, quat0, cov0)
    stream = torch.cuda.ExternalStream(b.cuda_stream, device=dev)
    for mode in (0, 1):
        d = synth.synth_spec_inputs(truth, 0, Tc)
        spec = SynthSpec(synth.SEED, d["imu_mean"], d["imu_step"], d["streams"], mode=mode)
        rows = b.synthesize(spec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import ctypes as C
        from pronto_b200 import capi
        zp = (C.c_void_p * 2)(rows["z"][0].data_ptr(), rows["z"][1].data_ptr())
        qp = (C.c_void_p * 2)(None, rows["quat"][1].data_ptr())
        e0.record(stream)
        for _ in range(5):
            capi.check(b.lib.rbis_batch_synthesize(b.h, C.byref(spec.c), rows["imu"].data_ptr(), zp, qp))
        e1.record(stream)
        b.synchronize(); torch.cuda.synchronize()
        nbytes = sum(t.numel() * 8 for t in [rows["imu"], rows["z"][0], rows["z"][1], rows["quat"][1]])
        ms = e0.elapsed_time(e1) / 5
        print(f"synthesize mode {mode}: {ms:.3f} ms per {Tc}-step chunk of {N} filters ({nbytes / 1e6:.0f} MB written, {nbytes / ms / 1e6:.0f} GB/s)")
    specs, progs = [], []
    for c in range(K):
        d = synth.synth_spec_inputs(truth, c * Tc, Tc)
        specs.append(SynthSpec(synth.SEED, d["imu_mean"], d["imu_step"], d["streams"], mode=1))
        progs.append(make_ops(bench.chunk_events(Tc, c * Tc)[0]))
    sst = [MeasStream(synth.LEGODO_IDX, None, R_lego), MeasStream(synth.POSE_IDX, None, R_pose, quat=True)]
    for rep in range(2):
        b.set_state(vec0, quat0, cov0); b.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for c in range(K):
            b.run_fused_synth(progs[c], sst, specs[c])
        t1 = time.perf_counter()
        b.record(); e1.record(stream)
        b.synchronize(); torch.cuda.synchronize()
        print(f"run_fused_synth x {K}: {e0.elapsed_time(e1) / K:.3f} ms per call on the device, host enqueue {1e3 * (t1 - t0) / K:.3f} ms per call")
    chunks = [bench.device_chunk(truth, c * Tc, Tc, N, gen, dev) for c in range(K)]
    preps = [b.prepare_fused(progs[c], imu=chunks[c]["imu"], streams=[MeasStream(synth.LEGODO_IDX, chunks[c]["legodo"], R_lego),
                                                                      MeasStream(synth.POSE_IDX, chunks[c]["pose_z"], R_pose, quat=chunks[c]["pose_q"])]) for c in range(K)]
    for rep in range(2):
        b.set_state(vec0, quat0, cov0); b.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for c in range(K):
            b.run_prepared(preps[c])
        t1 = time.perf_counter()
        b.record(); e1.record(stream)
        b.synchronize(); torch.cuda.synchronize()
        print(f"run_fused (resident inputs) x {K}: {e0.elapsed_time(e1) / K:.3f} ms per call on the device, host enqueue {1e3 * (t1 - t0) / K:.3f} ms per call")

# ---- the bench's e2e_synth loop: one launch + statistics read-back per step, host one step ahead ----
from pronto_b200 import capi
with RBISBatch(N) as b:
    b.set_process_noise(p["q_gyro"], p["q_accel"], p["q_gyro_bias"], p["q_accel_bias"])
    b.set_state(vec0, quat0, cov0)
    stream = torch.cuda.ExternalStream(b.cuda_stream, device=dev)
    tv, tq = synth.truth_state_at(truth, K * Tc - 1)
    n_local = (N + 1023) // 1024
    res = [torch.empty((n_local, capi.NUM_STATS), dtype=torch.float64, pin_memory=True).numpy() for _ in range(2)]
    for variant in ("synth+stats+wait", "synth+wait", "synth+stats, sync at end"):
        for rep in range(2):
            b.set_state(vec0, quat0, cov0); b.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            tick = None
            t_enq = t_wait = 0.0
            for c in range(K):
                t0 = time.perf_counter()
                b.run_fused_synth(progs[c], sst, specs[c])
                if "stats" in variant:
                    b.stats_enqueue(tv, tq, res[c % 2], chunk=1024)
                t = b.record()
                t1 = time.perf_counter()
                if "wait" in variant and tick is not None:
                    b.wait(tick)
                t2 = time.perf_counter()
                tick = t
                t_enq += t1 - t0; t_wait += t2 - t1
            b.wait(tick)
            e1.record(stream)
            b.synchronize(); torch.cuda.synchronize()
        print(f"{variant}: {e0.elapsed_time(e1) / K:.3f} ms per step; host enqueue {1e3 * t_enq / K:.3f} ms, host wait {1e3 * t_wait / K:.3f} ms per step")
