"""Calibration statistics across ranks: the two real exchange steps of the path.

Calibration batches are split over the ranks (``shard_batches``); each rank reduces its own batches
on its GPU and ONE collective per statistic combines them:

* activation ranges, momentum 0 — a single all-reduce with ``MIN`` over ``[min_0.., −max_0..]`` for
  all tensors at once (exact and order-independent, minmax.py:63-64);
* activation ranges, momentum > 0 — the EMA depends on the batch order (minmax.py:55-60), so the
  per-batch (min, max) pairs are all-gathered (8 bytes per batch and tensor) and replayed in the
  global batch order on every rank;
* GPTQ Hessians — rank r holds ``H_r = (2/n_r)·Σ_r XᵀX`` after its batches (gptq.py:254-258 applied
  locally); the global ``H = (2/n)·Σ XᵀX`` is ``Σ_r (n_r/n)·H_r``: scale in place, all-reduce SUM
  (NCCL over NVLink; or ``reduce`` to the rank that owns the layer's solve).

The functions take torch tensors on whatever device the process group's backend serves (CUDA for
NCCL, CPU for the gloo tests); no kernel of libb200quant is involved here.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from onnx_quantize_b200.parallel.shard import world


def shard_batches(n_batches: int, rank: int | None = None, n_ranks: int | None = None) -> list[int]:
    """Global batch indices of one rank: round-robin, so every rank sees early and late batches."""
    r, w = world()
    rank = r if rank is None else rank
    n_ranks = w if n_ranks is None else n_ranks
    return list(range(rank, n_batches, n_ranks))


def allreduce_minmax(ranges: torch.Tensor, group=None) -> torch.Tensor:
    """``ranges`` (n_tensors, 2) = per-tensor (min, max) of this rank → global, in place.

    One collective for all tensors: max(x) = −min(−x), so ``[min.., −max..]`` reduces with MIN."""
    if world()[1] == 1:
        return ranges
    packed = torch.cat([ranges[:, 0], -ranges[:, 1]])
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    n = ranges.shape[0]
    ranges[:, 0] = packed[:n]
    ranges[:, 1] = -packed[n:]
    return ranges


def gather_batch_pairs(local_pairs: torch.Tensor, n_batches: int, group=None) -> torch.Tensor:
    """Per-batch (min, max) pairs of this rank's batches (``shard_batches`` order), shape
    (n_local, n_tensors, 2) → all pairs in GLOBAL batch order, shape (n_batches, n_tensors, 2)."""
    rank, n_ranks = world()
    if n_ranks == 1:
        return local_pairs
    per_rank = -(-n_batches // n_ranks)
    padded = torch.zeros((per_rank,) + tuple(local_pairs.shape[1:]), dtype=local_pairs.dtype,
                         device=local_pairs.device)
    padded[: local_pairs.shape[0]] = local_pairs
    out = [torch.empty_like(padded) for _ in range(n_ranks)]
    dist.all_gather(out, padded, group=group)
    ordered = torch.empty((n_batches,) + tuple(local_pairs.shape[1:]), dtype=local_pairs.dtype,
                          device=local_pairs.device)
    for r in range(n_ranks):
        idx = shard_batches(n_batches, r, n_ranks)
        ordered[idx] = out[r][: len(idx)]
    return ordered


def replay_ema(pairs: torch.Tensor, momentum: float) -> torch.Tensor:
    """minmax.py:50-64 over (n_batches, n_tensors, 2) in batch order, float32 → (n_tensors, 2).
    (Host-side restatement used after ``gather_batch_pairs``; a handful of scalars per tensor.)"""
    state = pairs[0].clone()
    m = torch.tensor(momentum, dtype=pairs.dtype, device=pairs.device)
    one_m = torch.tensor(1 - momentum, dtype=pairs.dtype, device=pairs.device)
    for b in range(1, pairs.shape[0]):
        if momentum > 0:
            state = m * state + one_m * pairs[b]
        else:
            state[:, 0] = torch.minimum(state[:, 0], pairs[b, :, 0])
            state[:, 1] = torch.maximum(state[:, 1], pairs[b, :, 1])
    return state


def allreduce_hessian(h: torch.Tensor, n_local: int, group=None, dst: int | None = None
                      ) -> tuple[torch.Tensor, int]:
    """Combine per-rank Hessians ``H_r = (2/n_r)·Σ XᵀX`` into the global one, in place.

    Returns ``(H, n_total)``.  With ``dst`` the sum is only delivered to that rank (the owner of
    the layer's solve); other ranks' ``h`` is then scratch."""
    _, n_ranks = world()
    if n_ranks == 1:
        return h, n_local
    counts = torch.tensor([float(n_local)], dtype=torch.float64, device=h.device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    n_total = int(round(float(counts.item())))
    h.mul_(n_local / n_total if n_total else 0.0)
    if dst is None:
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(h, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return h, n_total
