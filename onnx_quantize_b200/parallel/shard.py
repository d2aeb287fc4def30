"""Sharding independent units (weight matrices, GPTQ solves) over the ranks of one box.

The hot path partitions naturally: every (K,N) weight is an independent unit, so
ranks take disjoint subsets and NO data-path collective is needed — only the results travel (packed
codes + parameters, ~4.5 bit/element) to the rank that builds the ONNX initializers.  One process
per GPU, ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

from typing import Sequence

import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def assign_units(costs: Sequence[float], n_ranks: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of units to ranks.

    ``costs[i]`` is the cost of unit i (K·N for RTN, ~K²·(K/3 + N) for a GPTQ solve).  Returns the
    unit indices of every rank, each list in descending cost order.  Deterministic (ties by index),
    so every rank computes the same plan without communicating.  A Llama-3-8B layer set (224
    matrices of 3 sizes) balances to < 2 % over 8 ranks.
    """
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * n_ranks
    plan: list[list[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda j: (load[j], j))
        plan[r].append(i)
        load[r] += float(costs[i])
    return plan


def imbalance(costs: Sequence[float], plan: list[list[int]]) -> float:
    """max rank load / mean rank load − 1."""
    loads = [sum(float(costs[i]) for i in p) for p in plan]
    mean = sum(loads) / len(loads)
    return max(loads) / mean - 1.0 if mean > 0 else 0.0


def quantize_weights_sharded(named_weights: dict, spec, *, dst: int = 0, group=None,
                             publish: bool = True) -> dict | None:
    """RTN-quantize a model's weights across the ranks of ``group``.

    ``named_weights`` (initializer name → host (K,N) float32 array) must be identical on every
    rank (each rank loads the same model).  Every rank runs ``pipeline.quantize_weights_bulk`` on
    its share; the triples are gathered on ``dst``, which returns ``{name: (codes, scale, zp)}``
    and (``publish``) stores them in ``parallel.prequantized`` where the registered plugins'
    ``quantize_weights`` find them (the reference's rewriter is single-threaded and asks for one
    weight at a time: qrules/_common.py:133).  Other ranks return None.
    """
    from onnx_quantize_b200.parallel import prequantized
    from onnx_quantize_b200.pipeline import quantize_weights_bulk

    rank, n_ranks = world()
    names = sorted(named_weights)
    costs = [named_weights[n].shape[0] * named_weights[n].shape[1] for n in names]
    mine = assign_units(costs, n_ranks)[rank]
    results = quantize_weights_bulk([named_weights[names[i]] for i in mine], spec)
    local = {names[i]: r for i, r in zip(mine, results)}
    if n_ranks == 1:
        merged = local
    else:
        gathered = [None] * n_ranks if rank == dst else None
        dist.gather_object(local, gathered, dst=dst, group=group)
        if rank != dst:
            return None
        merged = {}
        for part in gathered:
            merged.update(part)
    if publish:
        for name, triple in merged.items():
            prequantized.put(name, triple)
    return merged
