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


_comm_ready = False


def warm_up_communicator(device) -> float:
    """Force the lazy NCCL communicator creation (ring/tree discovery, NVLS setup: seconds at 8 ranks) with a 1-element
    all-reduce so that it is not billed to the weight broadcast.  -> seconds spent (0 when already done / single rank)."""
    global _comm_ready
    import time

    if _comm_ready or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0.0
    t0 = time.time()
    t = torch.zeros(1, device=device)
    dist.all_reduce(t)
    if t.is_cuda:
        torch.cuda.synchronize(device)
    _comm_ready = True
    return time.time() - t0


def replicate_params(specs: Sequence[Spec], init_fn, dtype: torch.dtype, device, src: int = 0,
                     timings: Dict[str, float] | None = None) -> Dict[str, torch.Tensor]:
    """Rank `src` materialises the parameters (init_fn() -> dict name -> tensor) into the arena; ONE pass of broadcasts
    over the packed arena replicates them to every rank (NCCL over NVLink / NVSwitch; chunks of < 2^31 elements, the
    collective's count limit).  With world size 1 this is just init + pack.
    timings (optional dict) receives comm_init_s / init_s / broadcast_s / broadcast_gbs, accumulated over calls."""
    import time

    def sync():
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    t_comm = warm_up_communicator(device)
    arena, views = pack_arena(specs, dtype, device)
    t0 = time.time()
    if rank == src:
        params = init_fn()
        for name, v in views.items():
            v.copy_(params[name].to(device=v.device, dtype=dtype))
        del params
    sync()
    t_init = time.time() - t0
    t_bc = 0.0
    if world > 1:
        dist.barrier()                      # the broadcast is timed from the moment the source data exists
        sync()
        t0 = time.time()
        step = (1 << 31) - 16
        for off in range(0, arena.numel(), step):
            dist.broadcast(arena[off:off + step], src=src)
        sync()
        t_bc = time.time() - t0
    if timings is not None:
        timings["comm_init_s"] = timings.get("comm_init_s", 0.0) + t_comm
        timings["init_s"] = timings.get("init_s", 0.0) + t_init
        timings["broadcast_s"] = timings.get("broadcast_s", 0.0) + t_bc
        timings["broadcast_bytes"] = timings.get("broadcast_bytes", 0) + (arena.numel() * arena.element_size() if world > 1 else 0)
        if timings["broadcast_s"] > 0:
            timings["broadcast_gbs"] = timings["broadcast_bytes"] / timings["broadcast_s"] / 1e9
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
