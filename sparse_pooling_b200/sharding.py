"""Frame sharding across the GPUs of one box.

Frames are independent (each has its own point cloud, M and feature maps; the
reference processes one frame per step, sparse_pool_utils.py:98, rpn_model.py:762-763),
so a batch is split by frame with NO collective on the data path.  torch.distributed
is used only to combine per-rank timing records (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def frames_for_rank(n_frames, rank, world):
    """Contiguous split of frame ids [0, n_frames) -- rank r gets the r-th slice; sizes differ by at most 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def max_over_ranks(value, device="cpu"):
    """Max of a per-rank scalar (a device-side time): what a multi-GPU number must be quoted on."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_records(record):
    """All ranks' small python records on every rank (benchmark tables only)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [record]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, record)
    return out
