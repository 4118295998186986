"""Multi-GPU batch sharding: one process per GPU, independent seeds/prompts per rank, weights replicated.

The reference has no distributed code; images are independent, so the path shards by batch with
NO per-step collective.  The only communication is the one-time broadcast of the packed weight arena from rank 0
(NCCL over NVLink/NVSwitch) and, optionally, a gather of the finished uint8 images.

torch.distributed is the rendezvous/plumbing (backend "nccl" on GPUs, "gloo" in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

from .weights import Spec


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise the default process group from the torchrun environment (MASTER_ADDR must be 127.0.0.1-reachable)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous slice of the global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def pack_arena(specs: Sequence[Spec], dtype: torch.dtype, device) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One flat buffer holding every parameter (each 16-byte aligned) + name -> view dict."""
    offsets, total = [], 0
    for _, shape, _ in specs:
        n = 1
        for s in shape:
            n *= s
        offsets.append(total)
        total += (n + 7) // 8 * 8
    arena = torch.empty(total, dtype=dtype, device=device)
    views = {}
    for (name, shape, _), off in zip(specs, offsets):
        n = 1
        for s in shape:
            n *= s
        views[name] = arena[off:off + n].view(*shape)
    return arena, views


def replicate_params(specs: Sequence[Spec], init_fn, dtype: torch.dtype, device, src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` materialises the parameters (init_fn() -> dict name -> tensor) into the arena; one broadcast
    replicates them to every rank.  With world size 1 this is just init + pack."""
    arena, views = pack_arena(specs, dtype, device)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == src:
        params = init_fn()
        for name, v in views.items():
            v.copy_(params[name].to(device=v.device, dtype=dtype))
        del params
    if world > 1:
        # chunked so that no single collective exceeds 2^31 elements
        step = 1 << 30
        for off in range(0, arena.numel(), step):
            dist.broadcast(arena[off:off + step], src=src)
    return views


def gather_to_rank0(t: torch.Tensor) -> List[torch.Tensor] | None:
    """Gather equally-shaped result tensors (e.g. uint8 images) on rank 0."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [t]
    world = dist.get_world_size()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return outs if dist.get_rank() == 0 else None


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
