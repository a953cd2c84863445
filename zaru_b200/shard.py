"""Multi-GPU sharding of the perception path: independent streams, no data-path collective.

Frames / camera streams are independent end to end (the reference's own parallelism is one `Detector` per
thread: crates/zaru/examples/eval_face_recognition.rs:67-70), so stream `s` simply goes to GPU `s mod G`
(one process per GPU, weights replicated).  The only communication is the timing reduction (max over ranks)
and an optional gather of the fixed-size result records.
"""
from __future__ import annotations


def streams_for_rank(n_streams: int, world: int, rank: int) -> list[int]:
    """Round-robin partition: stream s -> rank s % world."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return list(range(rank, n_streams, world))


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (time-like quantities)."""
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sum_over_ranks(values, dist=None, device="cpu"):
    import torch

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def aggregate_throughput(units_this_rank: float, seconds_this_rank: float, dist=None, device="cpu") -> float:
    """Whole-job throughput: units processed by ALL ranks / the slowest rank's time."""
    (total,) = sum_over_ranks([units_this_rank], dist, device)
    (slowest,) = max_over_ranks([seconds_this_rank], dist, device)
    return total / slowest


def gather_records(records, dist=None):
    """Gather per-rank result records (small, picklable) on every rank, ordered by rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [records]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, records)
    return out
