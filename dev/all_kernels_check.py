#!/usr/bin/env python
"""dev/all_kernels_check.py -- one small pass through every fused-kernel family with a bit-identity check (compute-sanitizer is closed on the GPU pool):
lane-per-filter decoupled + dense, warp-group 4 / 8 lanes, SYN instantiations, snapshot / restore, snapshot statistics."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from common import gpu_streams, nominal_q, scenario, random_ensemble
from pronto_b200 import MeasStream, RBISBatch, SynthSpec, capi, synth

N, T = 37, 24
sc = scenario(N, T, tumbling=True)
st = sc["st"]
ev = list(st["events"])
half = len(ev) // 2
prog = ev[:half] + [(capi.OP_SNAPSHOT, 0, 0, ev[half - 1][3], 0.0)] + ev[half:] + [(capi.OP_RESTORE, 0, 0, ev[half - 1][3], 0.0)] + ev[half:]
tv, tq = synth.truth_state_at(sc["truth"], T - 1)
ref = None
for cfg in (dict(mapping=1, lane_filters_per_cta=384), dict(mapping=1, lane_filters_per_cta=128), dict(mapping=1, dense_only=True),
            dict(mapping=4), dict(mapping=8), dict(mapping=4, dense_only=True)):
    with RBISBatch(N, snapshot_slots=2, **cfg) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused(prog, imu=st["imu"], streams=gpu_streams(st))
        out = torch.empty((1, capi.NUM_STATS), dtype=torch.float64, pin_memory=True).numpy()
        b.run_fused([(capi.OP_SNAPSHOT, 0, 1, ev[-1][3], 0.0)], imu=st["imu"], streams=gpu_streams(st))
        b.wait(b.stats_snapshot_enqueue(1, tv, tq, out, chunk=64))
        got = b.get_state()
    if ref is None:
        ref = got
    same = all(np.array_equal(x, y) for x, y in zip(got[:4], ref[:4]))
    print(cfg, "variant", "same bits" if same else "DIFFERENT", flush=True)
    assert same
d = synth.synth_spec_inputs(sc["truth"], 0, T)
for cfg in (dict(mapping=1), dict(mapping=4), dict(mapping=8), dict(mapping=1, synth_materialize=True)):
    with RBISBatch(N, **cfg) as b:
        b.set_process_noise(*nominal_q())
        b.set_state(sc["vec"], sc["quat"], sc["cov"])
        b.run_fused_synth(ev, [MeasStream(synth.LEGODO_IDX, None, st["R_legodo"]), MeasStream(synth.POSE_IDX, None, st["R_pose"], quat=True)],
                          SynthSpec(synth.SEED, d["imu_mean"], d["imu_step"], d["streams"], mode=1))
        got = b.get_state()
    if cfg == dict(mapping=1):
        ref = got
    same = all(np.array_equal(x, y) for x, y in zip(got[:4], ref[:4]))
    print("synth", cfg, "same bits" if same else "DIFFERENT", flush=True)
    assert same
print("ok")
