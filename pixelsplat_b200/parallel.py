"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The hot path shards by scene with no data-path collective: every rank rasterizes
its own scenes.  The only collectives are
  * the gradient all-reduce of the trainable (encoder) parameters, once per step -- DDP semantics
    (mean over ranks), flattened into a few large buckets so NVLink/NVSwitch sees big messages
    (the reference gets this from Lightning's DDP strategy, /root/reference/src/main.py:94-98);
  * max-over-ranks of a timing, for benchmarks.
"""
from __future__ import annotations

import os
from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Initialises the default process group from the torchrun environment (no-op for world 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of `n_items` scenes: the first n % world ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def scene_seed(base_seed: int, rank: int, index: int) -> int:
    """Disjoint synthetic-scene seeds per rank (the reference seeds `seed + rank`, main.py:106)."""
    return base_seed + 1000 * rank + index


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> int:
    """Averages `.grad` over ranks in flat buckets of ~bucket_bytes; parameters whose grad is None
    (unused this step -- the reference runs DDP with find_unused_parameters=True) contribute zeros.
    Returns the number of collectives issued."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    plist = [p for p in params if p.requires_grad]
    buckets: list[list[torch.nn.Parameter]] = [[]]
    size = 0
    for p in plist:
        nbytes = p.numel() * p.element_size()
        if buckets[-1] and size + nbytes > bucket_bytes:
            buckets.append([])
            size = 0
        buckets[-1].append(p)
        size += nbytes
    n = 0
    for b in buckets:
        if not b:
            continue
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= world
        off = 0
        for p in b:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        n += 1
    return n


def aggregate_throughput(units_per_rank: Sequence[float] | float, seconds_max: float, world: int) -> float:
    """Whole-job throughput: all ranks' units over the slowest rank's time."""
    total = sum(units_per_rank) if not isinstance(units_per_rank, (int, float)) else units_per_rank * world
    return total / seconds_max
