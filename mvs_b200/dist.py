"""Multi-GPU plumbing for the hot path: reference views are independent units, so
the path shards over the batch of reference views with NO data-path collective at inference.  The
only communication is rendezvous, a barrier and the max-over-ranks of the device time (bench.py), plus
an optional metric reduce to rank 0 mirroring CasMVSNet/utils.py:183-201.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str = "nccl", device=None):
    """init_process_group(env://) when WORLD_SIZE > 1 (CasMVSNet/train.py:297-302); no-op otherwise."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, init_method="env://", **kwargs)
    return rank, local, world


def shard_ref_views(n_views: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of reference-view indices for `rank` (first n % world ranks get one
    extra) -- the DistributedSampler role of CasMVSNet/train.py:383-393 without padding/duplication."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """Device-time aggregation rule of bench.py: the slowest rank defines the step time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_scalars_to_rank0(scalars: dict) -> dict:
    """Mean of a dict of python floats on rank 0 (CasMVSNet/utils.py:183-201 reduce_scalar_outputs)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
        return dict(scalars)
    names = sorted(scalars)
    t = torch.tensor([float(scalars[k]) for k in names], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.reduce(t, dst=0)
    if dist.get_rank() == 0:
        t /= dist.get_world_size()
    return {k: float(v) for k, v in zip(names, t.cpu())}


def gather_counts(n_local: int) -> List[int]:
    """All ranks' unit counts (whole-job throughput = sum of units / max time)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [n_local]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, int(n_local))
    return out
