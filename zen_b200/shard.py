"""Multi-GPU plumbing of the batched path: independent streams are sharded over
ranks (one process per GPU), no collective touches the data path; the only
communication is a barrier and a MAX-reduction of the measured time.
Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def stream_seed(base, rank, streams_per_rank, s):
    """global, collision-free seed of local stream s on `rank` (weak scaling: every rank owns streams_per_rank streams)"""
    return base + rank * streams_per_rank + s


def shard_range(n_total, rank, world):
    """contiguous block of a fixed total (strong scaling); blocks differ by at most one stream"""
    q, r = divmod(n_total, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def max_over_ranks(dist, value, device=None):
    """max of a python float over all ranks (identity when not distributed)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, world, seconds_max):
    """whole-job throughput: units all ranks processed / slowest rank's time"""
    return world * units_per_rank / seconds_max
