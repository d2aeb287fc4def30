"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out.

These are thin wrappers that allocate outputs / workspace with torch and call the C ABI of
libb200quant.so (include/b200q.h) on torch's current stream.  They never synchronise and never
touch host memory; the NumPy-facing mirrors of the reference API
(``core/_algorithms/*.py``, ``core/_calibration/minmax.py``, ``qrules/_common.py``) are built on
top of them.
"""
from __future__ import annotations

import ctypes

import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200.core._dtypes import QuantType

_STRATEGY_NAMES = ("tensor", "channel", "group")
N_MSE_CANDIDATES = 20


def _qt(quant_type) -> int:
    name = QuantType.coerce(quant_type).short_name
    if name not in _lib.QTYPE:
        raise NotImplementedError(f"{quant_type} is not supported on the weight path")
    return _lib.QTYPE[name]


def _strategy(strategy) -> int:
    name = getattr(strategy, "value", strategy)
    if name not in _STRATEGY_NAMES:
        raise ValueError(f"unknown strategy {strategy!r}")
    return _lib.STRATEGY[name]


def _mse_mode(mse) -> int:
    """False → 0, True → 1 (two-tier search), "exact" → 2 (every candidate exactly)."""
    if isinstance(mse, str):
        return {"off": 0, "on": 1, "exact": 2}[mse]
    return int(bool(mse))


def _check_weight(w: torch.Tensor) -> tuple[int, int]:
    if not (isinstance(w, torch.Tensor) and w.is_cuda and w.dtype == torch.float32
            and w.dim() == 2 and w.is_contiguous()):
        raise ValueError("weight must be a contiguous 2-D float32 CUDA tensor of shape (K, N)")
    return int(w.shape[0]), int(w.shape[1])


def resolve_group(k: int, strategy: int, group_size) -> tuple[int, int]:
    """→ (rows of W per parameter row, groups per column) following utils.py:15-24."""
    if strategy != _lib.STRATEGY["group"]:
        return k, 1
    gs = k if (group_size in (-1, None) or group_size > k) else int(group_size)
    if gs <= 0 or k % gs:
        raise ValueError(f"group_size {group_size} does not divide in_channels {k}")
    return gs, k // gs


def num_rows(k: int, n: int, strategy: int, group_size) -> int:
    if strategy == _lib.STRATEGY["tensor"]:
        return 1
    return n * resolve_group(k, strategy, group_size)[1]


def output_shapes(k: int, n: int, quant_type, strategy, group_size=-1, layout="kn"):
    """((codes shape, uint8), (scale shape, float32), (zp shape, uint8)) of ``rtn_quantize``."""
    qt, st = _qt(quant_type), _strategy(strategy)
    gs, g = resolve_group(k, st, group_size)
    rows = num_rows(k, n, st, group_size)
    bits = 4 if qt in (0, 1) else 8
    if layout == "kn":
        return (k, n), (rows,), (rows,)
    if layout == "packed_flat":
        return ((k * n + 1) // 2,), (rows,), (rows,)
    return (n, g, gs * bits // 8), (n, g), (n, (g + 1) // 2 if (bits == 4 and g > 1) else g)


def rtn_quantize(w: torch.Tensor, quant_type, strategy, group_size=-1, is_symmetric=False,
                 reduce_range=False, clip_ratio=1.0, mse=False, layout="kn", return_info=False, out=None):
    """RTN (+MSE) quantization of one (K,N) float32 weight on the GPU.

    Returns ``(codes, scale, zp)`` as CUDA tensors:
      layout "kn"           codes uint8 (K,N) (ml_dtypes byte representation), scale f32 (rows,),
                            zp uint8 (rows,) — rows = 1 | N | N*K/gs in the reference's order;
      layout "packed_flat"  codes uint8 (ceil(K*N/2),) — the INT4/UINT4 initializer bytes;
      layout "matmul_nbits" codes uint8 (N, G, gs*bits/8), scale f32 (N, G), zp uint8
                            (N, ceil(G/2)) [4-bit, G>1] or (N, G).
    ``out`` = pre-allocated ``(codes, scale, zp)`` CUDA tensors of ``output_shapes(...)``.
    """
    lib = _lib.load()
    k, n = _check_weight(w)
    qt, st, lay = _qt(quant_type), _strategy(strategy), _lib.LAYOUT[layout]
    gs, g = resolve_group(k, st, group_size)
    rows = num_rows(k, n, st, group_size)
    bits = 4 if qt in (0, 1) else 8
    device = w.device
    if out is not None:
        codes, scale, zp = out
        want = output_shapes(k, n, quant_type, strategy, group_size, layout)
        if (tuple(codes.shape), tuple(scale.shape), tuple(zp.shape)) != tuple(tuple(s) for s in want) or not (
                codes.is_contiguous() and scale.is_contiguous() and zp.is_contiguous()
                and codes.dtype == torch.uint8 and scale.dtype == torch.float32 and zp.dtype == torch.uint8):
            raise ValueError("out tensors do not match output_shapes()")
    elif layout == "kn":
        codes = torch.empty((k, n), dtype=torch.uint8, device=device)
        zp = torch.empty((rows,), dtype=torch.uint8, device=device)
        scale = torch.empty((rows,), dtype=torch.float32, device=device)
    elif layout == "packed_flat":
        codes = torch.empty(((k * n + 1) // 2,), dtype=torch.uint8, device=device)
        zp = torch.empty((rows,), dtype=torch.uint8, device=device)
        scale = torch.empty((rows,), dtype=torch.float32, device=device)
    else:
        codes = torch.empty((n, g, gs * bits // 8), dtype=torch.uint8, device=device)
        zp_cols = (g + 1) // 2 if (bits == 4 and g > 1) else g
        zp = torch.empty((n, zp_cols), dtype=torch.uint8, device=device)
        scale = torch.empty((n, g), dtype=torch.float32, device=device)
    info = torch.zeros((2,), dtype=torch.int32, device=device) if (mse and return_info) else None
    nbytes = lib.b200q_rtn_workspace_bytes(k, n, st, int(group_size or -1), _mse_mode(mse))
    ws = dev.workspace(nbytes)
    rc = lib.b200q_rtn_quantize(w.data_ptr(), k, n, qt, st, int(group_size or -1),
                                int(bool(is_symmetric)), int(bool(reduce_range)),
                                float(clip_ratio), _mse_mode(mse), lay, codes.data_ptr(),
                                scale.data_ptr(), zp.data_ptr(), dev.ptr(info), ws.data_ptr(),
                                ws.numel(), dev.stream_ptr())
    _lib.check(rc, "b200q_rtn_quantize")
    return (codes, scale, zp, info) if return_info else (codes, scale, zp)


def _alloc_outputs(k, n, qt, st, group_size, layout, device):
    gs, g = resolve_group(k, st, group_size)
    rows = num_rows(k, n, st, group_size)
    bits = 4 if qt in (0, 1) else 8
    if layout == "kn":
        codes = torch.empty((k, n), dtype=torch.uint8, device=device)
    elif layout == "packed_flat":
        codes = torch.empty(((k * n + 1) // 2,), dtype=torch.uint8, device=device)
    else:
        codes = torch.empty((n, g, gs * bits // 8), dtype=torch.uint8, device=device)
    if layout == "matmul_nbits":
        zp_cols = (g + 1) // 2 if (bits == 4 and g > 1) else g
        return codes, torch.empty((n, g), dtype=torch.float32, device=device), \
            torch.empty((n, zp_cols), dtype=torch.uint8, device=device)
    return codes, torch.empty((rows,), dtype=torch.float32, device=device), \
        torch.empty((rows,), dtype=torch.uint8, device=device)


class RtnBatchPlan:
    """Pre-built job list for ``rtn_quantize_batch``: output tensors allocated once, reusable."""

    def __init__(self, weights, quant_type, strategy, group_size=-1, is_symmetric=False,
                 reduce_range=False, clip_ratio=1.0, mse=False, layout="kn"):
        lib = _lib.load()
        self.args = (_qt(quant_type), _strategy(strategy), int(group_size or -1),
                     int(bool(is_symmetric)), int(bool(reduce_range)), float(clip_ratio),
                     _mse_mode(mse), _lib.LAYOUT[layout])
        qt, st = self.args[0], self.args[1]
        self.weights = list(weights)
        self.outputs = []
        self.jobs = (_lib.RtnJob * len(self.weights))()
        for i, w in enumerate(self.weights):
            k, n = _check_weight(w)
            out = _alloc_outputs(k, n, qt, st, group_size, layout, w.device)
            self.outputs.append(out)
            self.jobs[i] = _lib.RtnJob(w.data_ptr(), k, n, out[0].data_ptr(), out[1].data_ptr(),
                                       out[2].data_ptr(), None)
        self.ws_bytes = lib.b200q_rtn_batch_workspace_bytes(self.jobs, len(self.weights), st,
                                                            self.args[2], self.args[6])
        if len(self.weights) and self.ws_bytes == 0:
            raise ValueError("invalid shape / group size in the job list")

    def run(self):
        lib = _lib.load()
        ws = dev.workspace(self.ws_bytes)
        qt, st, gs, sym, rr, clip, mse, lay = self.args
        rc = lib.b200q_rtn_quantize_batch(self.jobs, len(self.weights), qt, st, gs, sym, rr, clip,
                                          mse, lay, ws.data_ptr(), ws.numel(), dev.stream_ptr())
        _lib.check(rc, "b200q_rtn_quantize_batch")
        return self.outputs


def rtn_quantize_batch(weights, quant_type, strategy, group_size=-1, is_symmetric=False,
                       reduce_range=False, clip_ratio=1.0, mse=False, layout="kn"):
    """``rtn_quantize`` for a list of weights with one configuration, issued from one C call."""
    return RtnBatchPlan(weights, quant_type, strategy, group_size, is_symmetric, reduce_range,
                        clip_ratio, mse, layout).run()


def mse_error_table(w: torch.Tensor, quant_type, strategy, group_size=-1, is_symmetric=False,
                    reduce_range=False) -> torch.Tensor:
    """f32 (20, rows): the error sum of every shrink candidate for every parameter row."""
    lib = _lib.load()
    k, n = _check_weight(w)
    qt, st = _qt(quant_type), _strategy(strategy)
    rows = num_rows(k, n, st, group_size)
    out = torch.empty((N_MSE_CANDIDATES, rows), dtype=torch.float32, device=w.device)
    ws = dev.workspace(lib.b200q_rtn_workspace_bytes(k, n, st, int(group_size or -1), 1))
    rc = lib.b200q_mse_error_table(w.data_ptr(), k, n, qt, st, int(group_size or -1),
                                   int(bool(is_symmetric)), int(bool(reduce_range)), out.data_ptr(),
                                   ws.data_ptr(), ws.numel(), dev.stream_ptr())
    _lib.check(rc, "b200q_mse_error_table")
    return out


def row_ranges(w: torch.Tensor, quant_type, strategy, group_size=-1, is_symmetric=False,
               reduce_range=False, clip_ratio=1.0, mse=False):
    """(min, max) per parameter row, zero included: A2, or the MSE-optimal range (A6)."""
    lib = _lib.load()
    k, n = _check_weight(w)
    qt, st = _qt(quant_type), _strategy(strategy)
    rows = num_rows(k, n, st, group_size)
    lo = torch.empty((rows,), dtype=torch.float32, device=w.device)
    hi = torch.empty((rows,), dtype=torch.float32, device=w.device)
    ws = dev.workspace(lib.b200q_rtn_workspace_bytes(k, n, st, int(group_size or -1), int(bool(mse))))
    rc = lib.b200q_row_ranges(w.data_ptr(), k, n, qt, st, int(group_size or -1),
                              int(bool(is_symmetric)), int(bool(reduce_range)), float(clip_ratio),
                              int(bool(mse)), lo.data_ptr(), hi.data_ptr(), ws.data_ptr(), ws.numel(),
                              dev.stream_ptr())
    _lib.check(rc, "b200q_row_ranges")
    return lo, hi


def quantize_with_qparams(w: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor, quant_type,
                          strategy, group_size=-1, is_symmetric=False, reduce_range=False):
    lib = _lib.load()
    k, n = _check_weight(w)
    qt, st = _qt(quant_type), _strategy(strategy)
    rows = num_rows(k, n, st, group_size)
    if scale.numel() != rows or zp.numel() != rows:
        raise ValueError(f"scale/zp must hold {rows} entries")
    codes = torch.empty((k, n), dtype=torch.uint8, device=w.device)
    rc = lib.b200q_quantize_with_qparams(w.data_ptr(), k, n, qt, st, int(group_size or -1),
                                         int(bool(is_symmetric)), int(bool(reduce_range)),
                                         scale.data_ptr(), zp.data_ptr(), codes.data_ptr(),
                                         dev.stream_ptr())
    _lib.check(rc, "b200q_quantize_with_qparams")
    return codes


def qparams(rmin: torch.Tensor, rmax: torch.Tensor, quant_type, is_symmetric=False,
            reduce_range=False):
    """A3 on device: f32 ranges (zero already included) → (scale f32, zp uint8 bytes)."""
    lib = _lib.load()
    n = rmin.numel()
    scale = torch.empty((n,), dtype=torch.float32, device=rmin.device)
    zp = torch.empty((n,), dtype=torch.uint8, device=rmin.device)
    rc = lib.b200q_qparams(rmin.data_ptr(), rmax.data_ptr(), n, _qt(quant_type),
                           int(bool(is_symmetric)), int(bool(reduce_range)), scale.data_ptr(),
                           zp.data_ptr(), dev.stream_ptr())
    _lib.check(rc, "b200q_qparams")
    return scale, zp


def dequantize(codes: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor, quant_type, strategy,
               group_size=-1) -> torch.Tensor:
    lib = _lib.load()
    k, n = int(codes.shape[0]), int(codes.shape[1])
    out = torch.empty((k, n), dtype=torch.float32, device=codes.device)
    fn = lib.b200q_dequantize_float_zp if zp.dtype == torch.float32 else lib.b200q_dequantize
    rc = fn(codes.data_ptr(), k, n, _qt(quant_type), _strategy(strategy), int(group_size or -1),
            scale.data_ptr(), zp.data_ptr(), out.data_ptr(), dev.stream_ptr())
    _lib.check(rc, "b200q_dequantize")
    return out


def quantize_bias(bias: torch.Tensor, input_scale: float, weight_scale: torch.Tensor):
    lib = _lib.load()
    n, ns = bias.numel(), weight_scale.numel()
    q = torch.empty((n,), dtype=torch.int32, device=bias.device)
    s = torch.empty((ns,), dtype=torch.float32, device=bias.device)
    rc = lib.b200q_quantize_bias(bias.data_ptr(), n, weight_scale.data_ptr(), ns,
                                 ctypes.c_float(float(input_scale)), q.data_ptr(), s.data_ptr(),
                                 dev.stream_ptr())
    _lib.check(rc, "b200q_quantize_bias")
    return q, s


def debug_pow_approx(x: torch.Tensor) -> torch.Tensor:
    """|x| ** 2.4 as the first tier of the MSE search evaluates it (MUFU lg2 / ex2)."""
    lib = _lib.load()
    out = torch.empty_like(x)
    _lib.check(lib.b200q_debug_pow_approx(x.data_ptr(), x.numel(), out.data_ptr(), dev.stream_ptr()),
               "b200q_debug_pow_approx")
    return out


def pack4_flat(codes: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    n = codes.numel()
    out = torch.empty(((n + 1) // 2,), dtype=torch.uint8, device=codes.device)
    _lib.check(lib.b200q_pack4_flat(codes.data_ptr(), n, out.data_ptr(), dev.stream_ptr()),
               "b200q_pack4_flat")
    return out


def unpack4_flat(packed: torch.Tensor, n_elements: int) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty((n_elements,), dtype=torch.uint8, device=packed.device)
    _lib.check(lib.b200q_unpack4_flat(packed.data_ptr(), n_elements, out.data_ptr(),
                                      dev.stream_ptr()), "b200q_unpack4_flat")
    return out


def pack_matmul_nbits(codes: torch.Tensor, zp_rows: torch.Tensor, group_size: int, bits: int):
    lib = _lib.load()
    k, n = int(codes.shape[0]), int(codes.shape[1])
    if group_size <= 0 or k % group_size:
        raise ValueError("group_size must divide in_channels")
    g = k // group_size
    b = torch.empty((n, g, group_size * bits // 8), dtype=torch.uint8, device=codes.device)
    zp_cols = (g + 1) // 2 if (bits == 4 and g > 1) else g
    zp = torch.empty((n, zp_cols), dtype=torch.uint8, device=codes.device)
    rc = lib.b200q_pack_matmul_nbits(codes.data_ptr(), k, n, group_size, bits, zp_rows.data_ptr(),
                                     b.data_ptr(), zp.data_ptr(), dev.stream_ptr())
    _lib.check(rc, "b200q_pack_matmul_nbits")
    return b, zp


def minmax_reduce(x: torch.Tensor, out_pair: torch.Tensor | None = None) -> torch.Tensor:
    """Global (min, max) of a float32 CUDA tensor → f32[2] on device (no synchronisation)."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("activation must be a contiguous float32 CUDA tensor")
    n = x.numel()
    if out_pair is None:
        out_pair = torch.empty((2,), dtype=torch.float32, device=x.device)
    ws = dev.workspace(lib.b200q_minmax_workspace_bytes(n))
    rc = lib.b200q_minmax_reduce(x.data_ptr(), n, out_pair.data_ptr(), ws.data_ptr(), ws.numel(),
                                 dev.stream_ptr())
    _lib.check(rc, "b200q_minmax_reduce")
    return out_pair


def minmax_merge(state: torch.Tensor, valid: torch.Tensor, pairs: torch.Tensor, momentum: float):
    lib = _lib.load()
    rc = lib.b200q_minmax_merge(state.data_ptr(), valid.data_ptr(), pairs.data_ptr(),
                                pairs.numel() // 2, float(momentum), dev.stream_ptr())
    _lib.check(rc, "b200q_minmax_merge")


def minmax_partials_stride() -> int:
    """float2 entries per batch slot of ``minmax_partials``."""
    return int(_lib.load().b200q_minmax_partials_stride())


def minmax_partials(x: torch.Tensor, slot: torch.Tensor, count: torch.Tensor) -> None:
    """Per-CTA partial (min, max) pairs of one batch into ``slot`` (f32 (stride, 2)); the number of
    valid pairs into the device int32 ``count``.  One launch, no fold (see ``minmax_fold_merge``).
    Inside ``_device.inputs_resident()`` the launch overlaps the tail of the previous kernel."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise ValueError("activation must be a contiguous float32 CUDA tensor")
    _lib.check(lib.b200q_minmax_partials(x.data_ptr(), x.numel(), slot.data_ptr(), count.data_ptr(),
                                         dev.stream_ptr()), "b200q_minmax_partials")


def minmax_fold_merge(state: torch.Tensor, valid: torch.Tensor, slots: torch.Tensor, counts: torch.Tensor,
                      n_batches: int, momentum: float, out_pairs: torch.Tensor | None = None,
                      out_range: torch.Tensor | None = None) -> None:
    """Fold ``n_batches`` slots and apply the running min/max (or EMA) update in batch order;
    ``out_range`` (f32[2]) receives the range with zero included (minmax.py:84-87)."""
    lib = _lib.load()
    _lib.check(lib.b200q_minmax_fold_merge(state.data_ptr(), valid.data_ptr(), slots.data_ptr(),
                                           counts.data_ptr(), int(n_batches), float(momentum),
                                           dev.ptr(out_pairs), dev.ptr(out_range), dev.stream_ptr()),
               "b200q_minmax_fold_merge")
