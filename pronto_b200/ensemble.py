"""Multi-GPU plumbing for the ensemble: shard by filter index, one process per GPU, no data-path
collective; ONE all-reduce at the end for the error / NEES statistics.

Bit-exactness for any GPU count: statistics are reduced on the device per fixed-size chunk of filters
in a fixed tree order; every rank writes its chunk partials into its own rows of a zero-initialised
[total_chunks][96] buffer; the SUM all-reduce is then exact (each row has exactly one non-zero
contributor, x + 0 = x); every rank finally adds the chunk rows in ascending order
(rbis_stats_reduce_chunks).  torch.distributed is plumbing only (NCCL on GPUs, gloo in CPU tests)."""
import numpy as np

from . import capi


def shard_range(n_filters, rank, world, chunk=1024):
    """Contiguous filter range [lo, hi) of `rank`; boundaries fall on statistics-chunk boundaries."""
    if chunk <= 0 or (chunk & (chunk - 1)):
        raise ValueError("chunk must be a power of two")
    n_chunks = (n_filters + chunk - 1) // chunk
    per = (n_chunks + world - 1) // world
    lo = min(n_filters, rank * per * chunk)
    hi = min(n_filters, (rank + 1) * per * chunk)
    return lo, hi


def allreduce_chunks(local_chunks, first_chunk, total_chunks, group=None, device=None):
    """local_chunks [n_local][96] (numpy) sitting at rows first_chunk.. of the global table.
    Returns the full [total_chunks][96] table, identical on every rank."""
    import torch
    import torch.distributed as dist

    local_chunks = np.ascontiguousarray(local_chunks, dtype=np.float64)
    if local_chunks.ndim != 2 or local_chunks.shape[1] != capi.NUM_STATS:
        raise ValueError("local_chunks must be [n][96]")
    if first_chunk < 0 or first_chunk + local_chunks.shape[0] > total_chunks:
        raise ValueError("chunk rows out of range")
    table = torch.zeros((total_chunks, capi.NUM_STATS), dtype=torch.float64, device=device or "cpu")
    if local_chunks.shape[0]:
        table[first_chunk:first_chunk + local_chunks.shape[0]] = torch.from_numpy(local_chunks).to(table.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    return table.cpu().numpy()


def ensemble_stats(batch, truth_vec, truth_quat, n_filters_total, rank, world, chunk=1024, group=None, device=None):
    """Statistics of the whole (sharded) ensemble: dict with the [96] totals and the chunk table."""
    from .batch import reduce_chunks

    lo, hi = shard_range(n_filters_total, rank, world, chunk)
    if hi - lo != batch.N:
        raise ValueError(f"rank {rank} holds {batch.N} filters, shard_range says {hi - lo}")
    local, _ = batch.stats(truth_vec, truth_quat, chunk=chunk)
    total_chunks = (n_filters_total + chunk - 1) // chunk
    table = allreduce_chunks(local, lo // chunk, total_chunks, group=group, device=device)
    return dict(total=reduce_chunks(table), chunks=table)


def summarize(total):
    """Human-readable view of the [96] totals."""
    n = total[46] - total[45]
    n = max(n, 1.0)
    mean = total[0:21] / n
    rms = np.sqrt(total[21:42] / n)
    return dict(filters=int(total[46]), non_finite=int(total[45]), mean_err=mean, rms_err=rms, mean_nees=total[42] / n,
                nees_in_95pct=total[47] / n, mean_loglik=total[44] / n)
