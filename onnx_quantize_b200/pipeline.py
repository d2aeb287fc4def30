"""Bulk weight quantization from HOST buffers: the end-to-end path a model quantization takes.

``quantize_weights_bulk`` streams a list of host (K,N) float32 weights through the GPU with three
CUDA streams — H2D copy of weight i+1, kernels of weight i and D2H copy of the results of weight
i-1 overlap — using double-buffered device slots and pinned staging for the results.  It is what
the multi-GPU pre-pass (``parallel/shard.py``) runs on each rank and what ``bench.py`` times as
the ``e2e`` figure.  The arrays returned are the kernels' own: code BYTES (K,N) uint8, flat float32
scales and flat zero-point bytes (the values are those of ``_rtn_quantize``; its dtypes and shapes
are applied by ``core._algorithms.rtn._finalize_triple``, which ``parallel.shard`` does before
publishing results to the plugins).
"""
from __future__ import annotations

import threading
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
    scale_dtype: object = np.dtype(np.float32)
    zp_dtype: object = None            # None: the code dtype (QWeightArgs' default)

    @classmethod
    def from_weight_args(cls, wa, layout: str = "kn") -> "RtnSpec":
        return cls(wa.dtype, wa.strategy.value, wa.group_size if wa.group_size else -1,
                   wa.symmetric, wa.reduce_range, wa.clip_ratio, wa.mse, layout,
                   wa.scale_dtype, wa.zp_dtype)

    def as_weight_args(self):
        """The same request in ``QWeightArgs`` attribute names (what the plugin sees)."""
        from types import SimpleNamespace

        qt = QuantType.from_string(self.quant_type) if isinstance(self.quant_type, str) else self.quant_type
        return SimpleNamespace(dtype=qt, strategy=getattr(self.strategy, "value", self.strategy),
                               group_size=self.group_size, symmetric=self.is_symmetric,
                               reduce_range=self.reduce_range, clip_ratio=self.clip_ratio, mse=self.mse,
                               scale_dtype=np.dtype(self.scale_dtype),
                               zp_dtype=np.dtype(self.zp_dtype) if self.zp_dtype is not None else qt.np_dtype)


class _PinnedCarver:
    """Pinned host memory for the results of one bulk call, carved out of fixed-size chunks.

    One ``torch.empty(pin_memory=True)`` per result tensor (672 of ~20 different sizes for a
    Llama-3-8B-shaped set) made the end-to-end time swing between 530 and 820 ms per step on the
    same box: whenever the host allocator's cache had no block of the right size it fell back to
    ``cudaHostAlloc``, which pins pages and synchronises (tools/prof_e2e.py).  Chunks of one size
    are always found in the cache once the previous call's results have been dropped; the arrays
    handed out keep their chunk alive."""

    CHUNK = 256 << 20

    def __init__(self):
        self.cur = None
        self.off = 0

    def take(self, shape, dtype: torch.dtype) -> torch.Tensor:
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
        padded = (nbytes + 255) // 256 * 256
        if padded > self.CHUNK:
            return torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        if self.cur is None or self.off + padded > self.CHUNK:
            self.cur = torch.empty((self.CHUNK,), dtype=torch.uint8, pin_memory=True)
            self.off = 0
        view = self.cur[self.off:self.off + nbytes].view(dtype).view(tuple(shape))
        self.off += padded
        return view


def _carve(buf: torch.Tensor, shapes, dtypes):
    """Contiguous views of ``shapes`` / ``dtypes`` laid out back to back (256-byte aligned) in ``buf``."""
    views, off = [], 0
    for shape, dtype in zip(shapes, dtypes):
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
        views.append(buf[off:off + nbytes].view(dtype).view(tuple(shape)))
        off += (nbytes + 255) // 256 * 256
    return views


class _Slot:
    def __init__(self, device):
        self.device = device
        self.w = None
        self.ready = torch.cuda.Event()      # H2D done
        self.consumed = torch.cuda.Event()   # kernels done reading w
        self.consumed.record()
        self.res = None                      # device staging of the results of the weight in this slot
        self.drained = torch.cuda.Event()    # D2H copies of the previous results of this slot done
        self.drained.record()

    def weight(self, k: int, n: int) -> torch.Tensor:
        need = k * n
        if self.w is None or self.w.numel() < need:
            self.w = torch.empty(need, dtype=torch.float32, device=self.device)
        return self.w[:need].view(k, n)

    def results(self, shapes) -> list:
        """Device tensors for (codes, scale, zp) of this slot's weight, reused from call to call —
        no allocator traffic (and no allocator-induced synchronisation) inside the pipeline."""
        dtypes = (torch.uint8, torch.float32, torch.uint8)
        need = sum((int(np.prod(s, dtype=np.int64)) * (4 if d == torch.float32 else 1) + 255) // 256 * 256
                   for s, d in zip(shapes, dtypes))
        if self.res is None or self.res.numel() < need:
            self.res = torch.empty(need, dtype=torch.uint8, device=self.device)
        return _carve(self.res, shapes, dtypes)


# The two staging slots of a device live as long as the process: their buffers are allocated on one
# stream and used on another, so handing them back to torch's stream-ordered allocator between
# calls could let a later allocation overwrite a weight that kernels of the previous call (which
# does not synchronise when the results stay on the device) are still reading.
_SLOTS: dict[int, list] = {}
_SLOT_LOCKS: dict[int, threading.Lock] = {}
_SLOTS_GUARD = threading.Lock()


def _device_slots(device) -> tuple[list, threading.Lock]:
    with _SLOTS_GUARD:
        idx = device.index
        if idx not in _SLOTS:
            _SLOTS[idx] = [_Slot(device), _Slot(device)]
            _SLOT_LOCKS[idx] = threading.Lock()
        return _SLOTS[idx], _SLOT_LOCKS[idx]


def release_slots() -> None:
    """Free the persistent staging buffers (after a ``torch.cuda.synchronize()``)."""
    with _SLOTS_GUARD:
        _SLOTS.clear()


def quantize_weights_bulk(weights, spec: RtnSpec, *, keep_on_device: bool = False):
    """Quantize every host weight in ``weights`` (iterable of (K,N) float32 arrays / CPU tensors).

    Returns a list of ``(codes, scale, zp)`` — NumPy arrays (pinned-memory backed) by default,
    CUDA tensors with ``keep_on_device=True``.  Copies from pinned inputs run at PCIe speed;
    pageable inputs work but are staged by the driver.
    """
    device = dev.require_cuda()
    slots, lock = _device_slots(device)
    with lock:
        return _bulk_locked(weights, spec, keep_on_device, device, slots)


def _bulk_locked(weights, spec: RtnSpec, keep_on_device: bool, device, slots):
    h2d, compute, d2h = torch.cuda.Stream(), torch.cuda.current_stream(), torch.cuda.Stream()
    results, pending = [], []
    pinned = _PinnedCarver()
    for i, w in enumerate(weights):
        src = w if isinstance(w, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
        if src.dtype != torch.float32 or src.dim() != 2:
            raise ValueError("weights must be 2-D float32")
        slot = slots[i & 1]
        k, n = src.shape
        with torch.cuda.stream(h2d):
            h2d.wait_event(slot.consumed)
            wd = slot.weight(k, n)
            dev.upload_into(src, wd)          # pinned: plain DMA; pageable model weights: staged, multi-threaded
            slot.ready.record(h2d)
        compute.wait_event(slot.ready)
        if keep_on_device:
            out = D.rtn_quantize(wd, spec.quant_type, spec.strategy, spec.group_size, spec.is_symmetric,
                                 spec.reduce_range, spec.clip_ratio, spec.mse, layout=spec.layout)
            slot.consumed.record(compute)
            results.append(out)
            continue
        compute.wait_event(slot.drained)
        staged = slot.results(D.output_shapes(k, n, spec.quant_type, spec.strategy, spec.group_size, spec.layout))
        D.rtn_quantize(wd, spec.quant_type, spec.strategy, spec.group_size, spec.is_symmetric, spec.reduce_range,
                       spec.clip_ratio, spec.mse, layout=spec.layout, out=tuple(staged))
        slot.consumed.record(compute)
        with torch.cuda.stream(d2h):
            d2h.wait_event(slot.consumed)
            host = []
            for t in staged:
                hbuf = pinned.take(t.shape, t.dtype)
                hbuf.copy_(t, non_blocking=True)
                host.append(hbuf)
            slot.drained.record(d2h)
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

