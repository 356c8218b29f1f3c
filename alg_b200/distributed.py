"""Multi-GPU plumbing of the ALG sampler: independent video samples, one per GPU.

Nothing inside the denoise loop crosses GPUs, so there is no data-path collective: ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the box, gloo in the CPU tests) is used only to broadcast rank 0's weights at init and to reduce
the timings at the end.  The reference has no distributed code at all; this is new.
"""
from __future__ import annotations

import os
from typing import Dict, Sequence

import torch
import torch.distributed as dist


def env_rank() -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def sample_seed(rank: int, base: int = 42) -> int:
    """run.py:94 seeds the one sample with 42; rank r owns sample r."""
    return base + rank


def broadcast_state_dict(sd: Dict[str, torch.Tensor], src: int = 0) -> Dict[str, torch.Tensor]:
    """In-place broadcast of every tensor of ``sd`` from ``src`` (all ranks hold same-shaped tensors; sorted by name so
    every rank issues the collectives in the same order)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for name in sorted(sd):
            dist.broadcast(sd[name], src=src)
    return sd


def arena_state_dict(shapes_dtypes: Dict[str, tuple], device, align: int = 256):
    """One flat byte arena holding every tensor of a model (name -> (shape, dtype)); returns (views, arena).

    The views are ordinary tensors (16-byte-aligned, as the TMA descriptors need) that ``load_state_dict`` borrows; the
    arena is what crosses NVLink: ONE ``dist.broadcast`` instead of one per parameter (the Wan DiT has ~1 100)."""
    offs, total = {}, 0
    for name in sorted(shapes_dtypes):
        shape, dtype = shapes_dtypes[name]
        n = 1
        for d in shape:
            n *= int(d)
        offs[name] = total
        total += (n * torch.empty((), dtype=dtype).element_size() + align - 1) // align * align
    arena = torch.empty(total, dtype=torch.uint8, device=device)
    views = {}
    for name, off in offs.items():
        shape, dtype = shapes_dtypes[name]
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        views[name] = arena[off:off + nbytes].view(dtype).view(*shape)
    return views, arena


def broadcast_arena(arena: torch.Tensor, src: int = 0, chunk_bytes: int = 1 << 31) -> torch.Tensor:
    """Broadcast a flat arena from ``src`` (in <= 2 GiB pieces: NCCL counts elements in 32-bit-friendly chunks)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for o in range(0, arena.numel(), chunk_bytes):
            dist.broadcast(arena[o:o + chunk_bytes], src=src)
    return arena


def max_over_ranks(values: Sequence[float], device) -> list:
    """Device-timed milliseconds -> the slowest rank's (what a multi-GPU throughput number must be computed from)."""
    t = torch.tensor(list(values), device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def aggregate_rate(units_per_rank: float, world: int, ms_max: float) -> float:
    """Whole-job units/s when every rank processed ``units_per_rank`` units in (at most) ``ms_max`` milliseconds."""
    return world * units_per_rank / (ms_max / 1e3)
