"""Multi-GPU plumbing: one process per GPU, environments sharded as contiguous slabs, no data-path
collective. The only collective on this path is a sum all-reduce of the 32-counter statistics vector
(NCCL on GPUs; the same code runs over gloo on CPU tensors, which is how it is tested without GPUs).

The reference has no distributed code at all; the partitioning rule here is the one
BASELINE.json's north_star states: rank r owns global env ids [offset_r, offset_r + count_r) and the
Philox stream of an env is keyed by its GLOBAL id, so results do not depend on the number of GPUs.
"""
import os

import torch
import torch.distributed as dist


def world_from_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = (0, 0, 1))."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_envs(total_envs, world_size, rank):
    """Contiguous slab of rank `rank`: (global offset, count). Slabs differ by at most one env."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def init_process_group(backend=None, device=None):
    """Initialises torch.distributed when launched under torchrun with WORLD_SIZE > 1."""
    rank, local, world = world_from_env()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return rank, local, world


def reduce_stats(stats, group=None):
    """Sum of the per-rank statistics vectors (int64 tensor, any device). Returns a new tensor."""
    out = stats.clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def max_over_ranks(value, device="cpu", group=None):
    """Max over ranks of a Python float (used for device-timed milliseconds)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def barrier(group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)
