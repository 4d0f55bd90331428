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

from typing import Callable, Tuple  # noqa: F401

import numpy as np


def shard_range(count: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous slice [lo, hi) of `count` units for `rank`; sizes differ by at most one"""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def gather_partials(local, group=None):
    """all-gather one fixed-size uint8 buffer per rank -> [world, len(local)] (rank-major).
    A CUDA tensor stays on its GPU (NCCL all-gather device to device, result a CUDA tensor): the
    partial products of an L2 sum never visit the host.  A host array is gathered as it is with gloo
    and staged through the rank's GPU with NCCL."""
    import torch
    import torch.distributed as dist
    cuda_in = _is_cuda_tensor(local)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if cuda_in:
            return local.reshape(1, -1)
        return np.ascontiguousarray(local, dtype=np.uint8).reshape(1, -1)
    world = dist.get_world_size(group)
    if cuda_in:
        t = local.contiguous().reshape(-1)
    else:
        t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8).copy())
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    if cuda_in:
        return out.reshape(world, -1)
    return out.cpu().numpy().reshape(world, -1)


def fold_l2_sum(local_partial, ncoeff: int, reduce_fn: Callable, group=None):
    """local_partial: this rank's L2 sum (ncoeff serialised GT elements; host array or CUDA tensor).
    Returns the sum over all ranks, identical on every rank, where the input lives.  reduce_fn is
    Engine.l2_sum_reduce (terms, nterms, ncoeff)."""
    allp = gather_partials(local_partial, group)
    world = allp.shape[0]
    if world == 1:
        return local_partial if _is_cuda_tensor(local_partial) else np.ascontiguousarray(local_partial, dtype=np.uint8)
    return reduce_fn(allp.reshape(-1), world, ncoeff)


def inner_product(engine, u, d1: int, v, d2: int, count_local: int, group=None, timings: dict = None):
    """Encrypted inner product sum_i u[i]*v[i] over ALL ranks' local shards (BASELINE.json config 5):
    MultPoly batch -> per-GPU GT product tree -> all-gather of (d1+d2) elements per rank -> fold.
    With CUDA inputs everything, the exchange included, stays on the GPUs.  `timings` (optional dict)
    receives the device ms of the three engine calls and the wall ms of the exchange."""
    import time
    prod = engine.multpoly_batch(u, d1, v, d2, count_local)
    if timings is not None:
        timings["multpoly_ms"] = engine.timing_last_call()
    part = engine.l2_sum_reduce(prod, count_local, d1 + d2)
    if timings is not None:
        timings["tree_ms"] = engine.timing_last_call()
        timings["exchange_bytes_per_rank"] = int(part.numel() if hasattr(part, "numel") else part.size)
    t0 = time.perf_counter()
    allp = gather_partials(part, group)
    if timings is not None:
        if _is_cuda_tensor(allp):
            import torch
            torch.cuda.synchronize(allp.device)
        timings["allgather_wall_ms"] = (time.perf_counter() - t0) * 1e3
    world = allp.shape[0]
    if world == 1:
        if timings is not None:
            timings["fold_ms"] = 0.0
        return part
    total = engine.l2_sum_reduce(allp.reshape(-1), world, d1 + d2)
    if timings is not None:
        timings["fold_ms"] = engine.timing_last_call()
    return total
