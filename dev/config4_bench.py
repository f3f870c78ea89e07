#!/usr/bin/env python
"""dev/config4_bench.py -- BASELINE configs[3]: a 1M-filter Monte-Carlo noise-parameter sweep sharded over the GPUs of one
box (torchrun, one rank per GPU), NCCL all-reduce of the error / NEES statistics at the end.

Grid: per-filter (q_gyro, q_accel, r_vxyz) on a 128 x 128 x 64 log-spaced grid spanning 1/3 .. 3 x nominal =
1,048,576 filters, contiguous shards by filter index.  Every filter of the grid replays the SAME noisy log (one shared
input column, rbis_batch_set_column_map), which is what a parameter sweep over a recorded log does; inputs come from
pinned HOST buffers through rbis_batch_run_fused, statistics are read back at the end.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 dev/config4_bench.py [launches]"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pronto_b200 import MeasStream, RBISBatch, capi, synth
from pronto_b200.batch import make_ops, reduce_chunks
from pronto_b200.ensemble import allreduce_chunks, summarize

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bench.bind_to_gpu_numa_node(local, rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
G1, G2, G3 = 128, 128, 64
NT = G1 * G2 * G3
N = NT // world
lo = rank * N
Tc, CHUNK, C_ = 200, 1024, 1
p = synth.NOMINAL
idx = np.arange(lo, lo + N)
ax = lambda n: np.exp(np.linspace(np.log(1 / 3), np.log(3.0), n))
q_gyro = p["q_gyro"] * ax(G1)[idx // (G2 * G3)]
q_accel = p["q_accel"] * ax(G2)[(idx // G3) % G2]
r_v = np.ascontiguousarray(np.tile((p["r_vxyz"] ** 2) * ax(G3)[idx % G3], (3, 1)))
gen = torch.Generator(device=dev); gen.manual_seed(7)  # same seed on every rank: the shared log
truth = synth.truth_trajectory((K + 2) * Tc)
vec0, quat0, cov0 = bench.initial_state(C_, gen, dev)
vec0, quat0, cov0 = (t.expand(-1, N).contiguous() for t in (vec0, quat0, cov0))
R_pose = np.diag([p["r_xyz"] ** 2] * 3 + [p["r_chi"] ** 2] * 3)
host, progs = [], []
for c in range(K + 2):
    ch = bench.device_chunk(truth, c * Tc, Tc, C_, gen, dev)
    host.append({k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v).numpy() for k, v in ch.items()})
    progs.append(make_ops(bench.chunk_events(Tc, c * Tc)[0]))
torch.cuda.synchronize()
tv, tq = synth.truth_state_at(truth, (K + 2) * Tc - 1)
with RBISBatch(N, device=local) as b:
    b.set_process_noise(np.ascontiguousarray(q_gyro), np.ascontiguousarray(q_accel), np.full(N, p["q_gyro_bias"]), np.full(N, p["q_accel_bias"]))
    b.set_state(vec0, quat0, cov0)
    cmap = np.zeros(N, dtype=np.int32)
    for w in (-1, 0, 1):
        b.set_column_map(w, cmap, C_)
    def step(c):
        b.run_fused(progs[c], imu=host[c]["imu"], streams=[MeasStream(synth.LEGODO_IDX, host[c]["legodo"], r_v, per_filter_diag=True),
                                                            MeasStream(synth.POSE_IDX, host[c]["pose_z"], R_pose, quat=host[c]["pose_q"])])
    step(0); step(1)
    b.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for c in range(2, K + 2):
        step(c)
    chunks, _ = b.stats(tv, tq, chunk=CHUNK)
    table = allreduce_chunks(chunks, lo // CHUNK, NT // CHUNK, device=dev if world > 1 else None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    variant = b.last_kernel_variant
tmax = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tot = reduce_chunks(table)
summ = summarize(tot)
import hashlib
digest = hashlib.sha256(np.ascontiguousarray(tot).tobytes()).hexdigest()[:16]
if rank == 0:
    print(f"config 4: {NT} filters on {world} GPU(s) ({N} per GPU), {K} launches x {Tc} steps, kernel variant {variant}: "
          f"{float(tmax.item()) * 1e3:.1f} ms (host wall clock, max over ranks, statistics all-reduce included) = "
          f"{NT * K * Tc / float(tmax.item()) / 1e9:.2f} G filter-steps/s; filters={summ['filters']} non_finite={summ['non_finite']} "
          f"mean NEES(9)={summ['mean_nees']:.3f}; sha256(totals)[:16]={digest}")
if world > 1:
    dist.destroy_process_group()
