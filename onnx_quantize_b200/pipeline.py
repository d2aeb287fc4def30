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


# ------------------------------------------------------------------------------------------------
# One weight per call, pageable NumPy in, NumPy out — the plugin seam itself (qrules/_common.py:133
# hands `w.const_value.numpy()` to `quantize_weights`, one weight at a time, and waits for the result)
# ------------------------------------------------------------------------------------------------
_STREAM_MIN_BYTES = 16 << 20
_DOWN_CHUNK_BYTES = 16 << 20


class _HostPipe:
    """Per-device streams, events and pinned down-staging for ``rtn_quantize_streamed``."""

    def __init__(self, device):
        self.h2d, self.d2h = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        self.down = [torch.empty((_DOWN_CHUNK_BYTES,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        self.down_np = [b.numpy() for b in self.down]
        self.down_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.lock = threading.Lock()


_PIPES: dict[int, _HostPipe] = {}


def _host_pipe(device) -> _HostPipe:
    with _SLOTS_GUARD:
        if device.index not in _PIPES:
            _PIPES[device.index] = _HostPipe(device)
        return _PIPES[device.index]


def streamed_chunk_rows(k: int, n: int, gs: int) -> int:
    """Rows of W per pipeline chunk of ``rtn_quantize_streamed`` (a multiple of the group size), or 0
    when the weight is not worth / not able to be chunked: chunks of 1/8 of the weight, at least
    8 MB and at most one 32 MB staging chunk of f32 (codes: a quarter of that)."""
    nbytes = 4 * k * n
    if nbytes < _STREAM_MIN_BYTES or gs <= 0 or gs >= k or k % gs:
        return 0
    target = min(max(nbytes // 8, 8 << 20), dev._STAGE_CHUNK_ELEMS * 4)
    rows = (target // (4 * n)) // gs * gs
    if rows < gs or rows >= k:
        return 0
    return int(rows)


def rtn_quantize_streamed(array: np.ndarray, quant_type, group_size: int, is_symmetric: bool,
                          reduce_range: bool, clip_ratio: float):
    """GROUP-strategy RTN (no MSE search) of one pageable host weight with the three legs of the
    call overlapped INSIDE the call: while row chunk c goes up (host threads fill a pinned staging
    chunk, DMA), chunk c-1 is quantized and chunk c-2's codes come down (DMA, host threads copy
    them into the result array).  Groups never span chunks (chunk rows are a multiple of the group
    size), so every chunk is an independent ``b200q_rtn_quantize`` on a (rows, N) slice; its
    parameters are scattered into the reference's (N, G) order on the device.  The MSE search cannot
    be chunked this way: its early stop is global over all rows (utils.py:232-237).

    Returns raw ``(codes (K,N) uint8, scale (N*G,) f32, zp (N*G,) uint8)`` NumPy arrays, or None
    when the weight does not qualify (small, ragged groups, not C-contiguous float32).
    """
    a = array
    if not (isinstance(a, np.ndarray) and a.ndim == 2 and a.dtype == np.float32 and a.flags.c_contiguous):
        return None
    k, n = a.shape
    gs = int(group_size)
    rows = streamed_chunk_rows(k, n, gs)
    if rows == 0:
        return None
    device = dev.require_cuda()
    pipe = _host_pipe(device)
    g_total, g_chunk = k // gs, rows // gs
    chunks = [(r0, min(r0 + rows, k)) for r0 in range(0, k, rows)]
    src = a.reshape(-1)
    out_codes = np.empty((k, n), dtype=np.uint8)
    out_flat = out_codes.reshape(-1)
    compute = torch.cuda.current_stream(device)
    with pipe.lock, dev._stage_lock:
        up_bufs, up_views, up_evs = dev._staging(device)
        wd = torch.empty((k, n), dtype=torch.float32, device=device)
        codes_d = torch.empty((k, n), dtype=torch.uint8, device=device)
        scale_d = torch.empty((n, g_total), dtype=torch.float32, device=device)
        zp_d = torch.empty((n, g_total), dtype=torch.uint8, device=device)
        sc_c = torch.empty((n * g_chunk,), dtype=torch.float32, device=device)
        zp_c = torch.empty((n * g_chunk,), dtype=torch.uint8, device=device)
        start = torch.cuda.Event()
        start.record(compute)
        pipe.h2d.wait_event(start)          # the fresh buffers may be recycled blocks of the compute stream
        pipe.d2h.wait_event(start)
        landed = [torch.cuda.Event() for _ in chunks]
        done = [torch.cuda.Event() for _ in chunks]
        for b in range(2):
            up_evs[b].synchronize()
            pipe.down_ev[b].synchronize()
        for c in range(len(chunks) + 2):
            futures = []
            if c < len(chunks):
                r0, r1 = chunks[c]
                up_evs[c & 1].synchronize()                       # the DMA that last read this staging chunk
                futures += dev.parallel_copy(up_views[c & 1][:(r1 - r0) * n], src[r0 * n:r1 * n])
            if c >= 2:
                r0, r1 = chunks[c - 2]
                pipe.down_ev[c & 1].synchronize()                 # codes of chunk c-2 are in the pinned chunk
                futures += dev.parallel_copy(out_flat[r0 * n:r1 * n], pipe.down_np[c & 1][:(r1 - r0) * n],
                                             piece=4 << 20)
            for f in futures:
                f.result()
            if c >= len(chunks):
                continue
            r0, r1 = chunks[c]
            m = r1 - r0
            with torch.cuda.stream(pipe.h2d):
                wd[r0:r1].view(-1).copy_(up_bufs[c & 1][:m * n], non_blocking=True)
                up_evs[c & 1].record(pipe.h2d)
                landed[c].record(pipe.h2d)
            compute.wait_event(landed[c])
            gc = m // gs
            sc_v, zp_v = (sc_c, zp_c) if gc == g_chunk else (sc_c[:n * gc], zp_c[:n * gc])
            D.rtn_quantize(wd[r0:r1], quant_type, "group", gs, is_symmetric, reduce_range, clip_ratio, False,
                           out=(codes_d[r0:r1], sc_v, zp_v))
            g0 = r0 // gs
            scale_d[:, g0:g0 + gc].copy_(sc_v.view(n, gc))
            zp_d[:, g0:g0 + gc].copy_(zp_v.view(n, gc))
            done[c].record(compute)
            with torch.cuda.stream(pipe.d2h):
                pipe.d2h.wait_event(done[c])
                pipe.down[c & 1][:m * n].copy_(codes_d[r0:r1].view(-1), non_blocking=True)
                pipe.down_ev[c & 1].record(pipe.d2h)
        scale_np = scale_d.reshape(-1).cpu().numpy()
        zp_np = zp_d.reshape(-1).cpu().numpy()
    return out_codes, scale_np, zp_np
