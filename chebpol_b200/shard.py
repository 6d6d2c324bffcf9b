"""Point partitioning for multi-GPU evaluation.

Points are independent, so the path shards with no data-path collective: rank r of
W evaluates the contiguous range ``shard_range(P, r, W)`` against its own replica of
the tensor and owns that slice of the result.  The only communication is the
barrier / MAX-reduction that brackets a timed region.  The same split is used
inside one process by ``chebpol_cuda_eval*`` (csrc/api.cu: run_sharded).
"""
from __future__ import annotations

import os


def shard_range(npts: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of ceil(P/W) points per rank (the last ranks may be short or empty)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} not in a world of {world}")
    per = (npts + world - 1) // world
    lo = min(npts, rank * per)
    hi = min(npts, lo + per)
    return lo, hi


def dist_env() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX all-reduce of a scalar (the timed region's duration)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
