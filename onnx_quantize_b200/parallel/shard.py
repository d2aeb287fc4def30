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


def packed_layout(shapes_and_dtypes) -> tuple[list[list[tuple[int, tuple, object]]], int]:
    """Byte offsets of result tensors laid out back to back (256-byte aligned) in one flat buffer.

    ``shapes_and_dtypes``: per unit, a sequence of ``(shape, torch dtype)``.  Returns
    ``([[(offset, shape, dtype), ...] per unit], total bytes)``; every rank computes the same
    layout from the shapes alone, so no metadata has to travel."""
    import numpy as np
    import torch

    out, off = [], 0
    for unit in shapes_and_dtypes:
        entries = []
        for shape, dtype in unit:
            nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
            entries.append((off, tuple(shape), dtype))
            off += (nbytes + 255) // 256 * 256
        out.append(entries)
    return out, off


def pack_results(results, layout, total: int, device):
    """Copy every result tensor of this rank into one flat uint8 buffer following ``layout``."""
    import torch

    flat = torch.empty((max(total, 1),), dtype=torch.uint8, device=device)
    for unit, entries in zip(results, layout):
        for t, (off, shape, dtype) in zip(unit, entries):
            nbytes = t.numel() * t.element_size()
            flat[off:off + nbytes].view(dtype).view(shape).copy_(t.reshape(shape))
    return flat


def unpack_results(flat, layout):
    """Views of ``flat`` (a uint8 tensor on any device) following ``layout``."""
    import numpy as np
    import torch

    out = []
    for entries in layout:
        unit = []
        for off, shape, dtype in entries:
            nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
            unit.append(flat[off:off + nbytes].view(dtype).view(shape))
        out.append(tuple(unit))
    return out


def gather_packed(flat, totals: Sequence[int], dst: int = 0, group=None):
    """Point-to-point gather of one flat uint8 buffer per rank (``totals[r]`` bytes, known to every
    rank) on ``dst`` — tensors over the process group's own transport (NCCL over NVLink for CUDA
    tensors, gloo for the CPU tests), no pickling.  Returns the list of buffers on ``dst`` (its own
    ``flat`` in place), None elsewhere."""
    import torch

    rank, n_ranks = world()
    if n_ranks == 1:
        return [flat]
    if rank != dst:
        if totals[rank] > 0:
            dist.send(flat[:totals[rank]], dst=dst, group=group)
        return None
    bufs, reqs = [], []
    for r in range(n_ranks):
        if r == dst:
            bufs.append(flat)
            continue
        buf = torch.empty((max(totals[r], 1),), dtype=torch.uint8, device=flat.device)
        bufs.append(buf)
        if totals[r] > 0:
            reqs.append(dist.irecv(buf[:totals[r]], src=r, group=group))
    for req in reqs:
        req.wait()
    return bufs


def quantize_weights_sharded(named_weights: dict, spec, *, dst: int = 0, group=None,
                             publish: bool = True) -> dict | None:
    """RTN-quantize a model's weights across the ranks of ``group``.

    ``named_weights`` (initializer name → host (K,N) float32 array) must be identical on every
    rank (each rank loads the same model).  Every rank runs ``pipeline.quantize_weights_bulk`` on
    its share (largest first); the packed results travel as tensors to ``dst`` (``gather_packed``)
    which copies them to the host once.  ``dst`` returns ``{name: (codes, scale, zp)}`` with
    exactly the dtypes and shapes ``_rtn_quantize`` returns for ``spec`` (``layout="kn"``; the
    packed layouts return the kernels' own arrays) and (``publish``, ``"kn"`` only) stores them in
    ``parallel.prequantized`` where the registered RTN plugin finds them (the reference's rewriter
    is single-threaded and asks for one weight at a time: qrules/_common.py:133).  Other ranks
    return None.
    """
    import torch

    from onnx_quantize_b200 import device_api as D
    from onnx_quantize_b200.core._algorithms.rtn import _finalize_triple
    from onnx_quantize_b200.core._qconfig import QuantizationStrategy
    from onnx_quantize_b200.parallel import prequantized
    from onnx_quantize_b200.pipeline import quantize_weights_bulk

    if publish and spec.layout != "kn":
        raise ValueError("only layout='kn' results can be published to the plugins: the rewriter packs "
                         "them itself (_prepare_for_matmul_nbits), a packed blob would be packed twice")
    rank, n_ranks = world()
    names = sorted(named_weights)
    costs = [named_weights[n].shape[0] * named_weights[n].shape[1] for n in names]
    plan = assign_units(costs, n_ranks)
    mine = plan[rank]
    if n_ranks == 1:
        raw = dict(zip((names[i] for i in mine),
                       quantize_weights_bulk([named_weights[names[i]] for i in mine], spec)))
    else:
        dtypes = (torch.uint8, torch.float32, torch.uint8)
        layouts, totals = [], []
        for r in range(n_ranks):
            shapes = [D.output_shapes(*named_weights[names[i]].shape, spec.quant_type, spec.strategy,
                                      spec.group_size, spec.layout) for i in plan[r]]
            lay, tot = packed_layout([list(zip(s, dtypes)) for s in shapes])
            layouts.append(lay)
            totals.append(tot)
        results = quantize_weights_bulk([named_weights[names[i]] for i in mine], spec, keep_on_device=True)
        device = torch.device("cuda", torch.cuda.current_device())
        flat = pack_results(results, layouts[rank], totals[rank], device)
        del results
        bufs = gather_packed(flat, totals, dst=dst, group=group)
        if bufs is None:
            return None
        raw = {}
        for r, buf in enumerate(bufs):
            host = torch.empty((max(totals[r], 1),), dtype=torch.uint8, pin_memory=True)
            host.copy_(buf, non_blocking=True)
            for i, unit in zip(plan[r], unpack_results(host, layouts[r])):
                raw[names[i]] = unit
        torch.cuda.synchronize()
        raw = {n: tuple(t.numpy() for t in unit) for n, unit in raw.items()}
    if spec.layout != "kn":
        return raw
    wa = spec.as_weight_args()
    strategy = QuantizationStrategy(wa.strategy)
    merged = {n: _finalize_triple(*raw[n], spec.quant_type, strategy, wa.scale_dtype, wa.zp_dtype)
              for n in names}
    if publish:
        for name, triple in merged.items():
            prequantized.put(name, prequantized.request_digest(wa, "rtn", named_weights[name]), triple)
    return merged
