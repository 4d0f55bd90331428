"""Multi-GPU plumbing: one process per GPU (torchrun), independent units per rank.

The BGN hot path shards by unit with no exchange step: every
coefficient encryption, every pairing and every decrypt is independent
(poly.go:15, 37, 140-141).  The single cross-unit dependency is the GT product
that realises an L2 sum (Add -> bgn.go:460): each rank reduces its own shard
on its GPU, the fixed-size serialised partial elements (ncoeff x 2B bytes per
rank, a few KB) are all-gathered, and every rank folds them with the same
device reduction.  torch.distributed is plumbing only (NCCL on the GPU box,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(count: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous slice [lo, hi) of `count` units for `rank`; sizes differ by at most one"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(local: np.ndarray, group=None) -> np.ndarray:
    """all-gather one fixed-size uint8 buffer per rank -> [world, len(local)] (rank-major).
    With the NCCL backend the buffer is staged through the rank's GPU."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.ascontiguousarray(local, dtype=np.uint8).reshape(1, -1)
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, -1)


def fold_l2_sum(local_partial: np.ndarray, ncoeff: int, reduce_fn: Callable[[np.ndarray, int, int], np.ndarray],
                group=None) -> np.ndarray:
    """local_partial: this rank's L2 sum (ncoeff serialised GT elements).  Returns the sum over all
    ranks, identical on every rank.  reduce_fn is Engine.l2_sum_reduce (terms, nterms, ncoeff)."""
    allp = gather_partials(local_partial, group)
    world = allp.shape[0]
    if world == 1:
        return np.ascontiguousarray(local_partial, dtype=np.uint8)
    return reduce_fn(allp.reshape(-1), world, ncoeff)


def inner_product(engine, u, d1: int, v, d2: int, count_local: int, group=None) -> np.ndarray:
    """Encrypted inner product sum_i u[i]*v[i] over ALL ranks' local shards (BASELINE.json config 5):
    MultPoly batch -> per-GPU GT product tree -> all-gather of (d1+d2) elements per rank -> fold."""
    prod = engine.multpoly_batch(u, d1, v, d2, count_local)
    part = engine.l2_sum_reduce(prod, count_local, d1 + d2)
    if hasattr(part, "cpu"):
        part = part.cpu().numpy()
    return fold_l2_sum(part, d1 + d2, engine.l2_sum_reduce, group)
