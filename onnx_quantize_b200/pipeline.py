"""Bulk weight quantization from HOST buffers: the end-to-end path a model quantization takes.

``quantize_weights_bulk`` streams a list of host (K,N) float32 weights through the GPU with three
CUDA streams — H2D copy of weight i+1, kernels of weight i and D2H copy of the results of weight
i-1 overlap — using double-buffered device slots and pinned staging for the results.  It is what
the multi-GPU pre-pass (``parallel/shard.py``) runs on each rank and what ``bench.py`` times as
the ``e2e`` figure.  Results are identical to calling ``_rtn_quantize`` per weight.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._dtypes import QuantType


@dataclass
class RtnSpec:
    """The arguments of ``_rtn_quantize`` (reference rtn.py:54-65) plus the output layout."""

    quant_type: QuantType
    strategy: str = "group"
    group_size: int = 128
    is_symmetric: bool = False
    reduce_range: bool = False
    clip_ratio: float = 1.0
    mse: bool = False
    layout: str = "kn"

    @classmethod
    def from_weight_args(cls, wa, layout: str = "kn") -> "RtnSpec":
        return cls(wa.dtype, wa.strategy.value, wa.group_size if wa.group_size else -1,
                   wa.symmetric, wa.reduce_range, wa.clip_ratio, wa.mse, layout)


class _Slot:
    def __init__(self, device):
        self.device = device
        self.w = None
        self.ready = torch.cuda.Event()      # H2D done
        self.consumed = torch.cuda.Event()   # kernels done reading w
        self.consumed.record()

    def weight(self, k: int, n: int) -> torch.Tensor:
        need = k * n
        if self.w is None or self.w.numel() < need:
            self.w = torch.empty(need, dtype=torch.float32, device=self.device)
        return self.w[:need].view(k, n)


def quantize_weights_bulk(weights, spec: RtnSpec, *, keep_on_device: bool = False):
    """Quantize every host weight in ``weights`` (iterable of (K,N) float32 arrays / CPU tensors).

    Returns a list of ``(codes, scale, zp)`` — NumPy arrays (pinned-memory backed) by default,
    CUDA tensors with ``keep_on_device=True``.  Copies from pinned inputs run at PCIe speed;
    pageable inputs work but are staged by the driver.
    """
    device = dev.require_cuda()
    h2d, compute, d2h = torch.cuda.Stream(), torch.cuda.current_stream(), torch.cuda.Stream()
    slots = [_Slot(device), _Slot(device)]
    results, pending = [], []
    for i, w in enumerate(weights):
        src = w if isinstance(w, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
        if src.dtype != torch.float32 or src.dim() != 2:
            raise ValueError("weights must be 2-D float32")
        slot = slots[i & 1]
        k, n = src.shape
        with torch.cuda.stream(h2d):
            h2d.wait_event(slot.consumed)
            wd = slot.weight(k, n)
            wd.copy_(src, non_blocking=True)
            slot.ready.record(h2d)
        compute.wait_event(slot.ready)
        out = D.rtn_quantize(wd, spec.quant_type, spec.strategy, spec.group_size, spec.is_symmetric,
                             spec.reduce_range, spec.clip_ratio, spec.mse, layout=spec.layout)
        slot.consumed.record(compute)
        if keep_on_device:
            results.append(out)
            continue
        done = torch.cuda.Event()
        done.record(compute)
        with torch.cuda.stream(d2h):
            d2h.wait_event(done)
            host = []
            for t in out:
                t.record_stream(d2h)
                hbuf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                hbuf.copy_(t, non_blocking=True)
                host.append(hbuf)
        pending.append(host)
    if keep_on_device:
        return results
    d2h.synchronize()
    for host in pending:
        results.append(tuple(h.numpy() for h in host))
    return results


def result_bytes(results) -> int:
    return int(sum(sum(a.nbytes if isinstance(a, np.ndarray) else a.numel() * a.element_size()
                       for a in r) for r in results))
