"""Host-side multi-GPU logic on CPU: sharding and the exact chunk all-reduce over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest

from pronto_b200.ensemble import allreduce_chunks, shard_range, summarize


def test_shard_range_covers_and_aligns():
    for n, world, chunk in [(65_536, 8, 1024), (1_048_576, 8, 1024), (3000, 4, 256), (100, 8, 32), (1, 2, 32)]:
        covered = 0
        for r in range(world):
            lo, hi = shard_range(n, r, world, chunk)
            assert lo == covered and lo % chunk == 0 or lo == n
            covered = hi
        assert covered == n
    assert shard_range(65_536, 3, 8) == (3 * 8192, 4 * 8192)
    with pytest.raises(ValueError):
        shard_range(10, 0, 2, chunk=100)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, table, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total_chunks = table.shape[0]
        per = (total_chunks + world - 1) // world
        lo, hi = rank * per, min(total_chunks, (rank + 1) * per)
        out = allreduce_chunks(table[lo:hi], lo, total_chunks)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_allreduce_chunks_is_exact_over_gloo_world2():
    import torch.multiprocessing as mp

    rng = np.random.default_rng(0)
    table = rng.normal(size=(7, 96)) * 10.0 ** rng.integers(-12, 12, size=(7, 96))
    table[3, 5] = -0.0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, table, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert np.array_equal(outs[r], table)  # bit-exact: one non-zero contributor per row
    # and the fixed-order final sum equals the single-process one
    from pronto_b200 import reduce_chunks

    assert np.array_equal(reduce_chunks(outs[0]), reduce_chunks(table))


def test_single_process_allreduce_is_identity():
    t = np.arange(3 * 96, dtype=np.float64).reshape(3, 96)
    assert np.array_equal(allreduce_chunks(t[1:], 1, 3)[1:], t[1:])
    assert np.all(allreduce_chunks(t[1:], 1, 3)[0] == 0)
    with pytest.raises(ValueError):
        allreduce_chunks(t, 2, 3)


def test_summarize():
    tot = np.zeros(96)
    tot[46], tot[45], tot[42], tot[47] = 10, 2, 72.0, 8
    tot[3], tot[24] = 4.0, 32.0
    s = summarize(tot)
    assert s["filters"] == 10 and s["non_finite"] == 2 and s["mean_nees"] == 9.0 and s["mean_err"][3] == 0.5 and s["rms_err"][3] == 2.0
