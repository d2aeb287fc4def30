"""NumPy-facing mirrors of the reference's numeric primitives, computed on the GPU.

Same names, argument meaning and result shapes/dtypes as the reference's
``core/_algorithms/utils.py``; every arithmetic step runs in libb200quant.so (via
``onnx_quantize_b200.device_api``).  Only the pure layout helpers ``_preprocess_array`` /
``_post_process_array`` (views and reshapes, no arithmetic; reference utils.py:6-39) stay in
NumPy.  Arrays are staged host→device→host here; the device-resident API is ``device_api``.
"""
from __future__ import annotations

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.core._qconfig import QuantizationStrategy, coerce_strategy


# ------------------------------------------------------------------------------------------------
# layout helpers (no arithmetic)
# ------------------------------------------------------------------------------------------------
def _preprocess_array(array, strategy, group_size=-1):
    """(K,N) → rows sharing one (scale, zp): tensor → as is, channel → W.T, group → (N*G, gs)."""
    strategy = coerce_strategy(strategy)
    if strategy == QuantizationStrategy.TENSOR:
        return array
    if strategy == QuantizationStrategy.CHANNEL:
        return array.T
    k = array.shape[0]
    gs = k if (group_size == -1 or group_size > k) else group_size
    return array.T.reshape((-1, gs))


def _post_process_array(preprocessed_array, original_array, strategy, group_size=-1):
    strategy = coerce_strategy(strategy)
    if strategy == QuantizationStrategy.TENSOR:
        return preprocessed_array
    if strategy == QuantizationStrategy.CHANNEL:
        return preprocessed_array.T
    return preprocessed_array.reshape(original_array.T.shape).T


# ------------------------------------------------------------------------------------------------
# host <-> device glue
# ------------------------------------------------------------------------------------------------
def _codes_to_numpy(codes_t: torch.Tensor, quant_type: QuantType) -> np.ndarray:
    """uint8 device bytes → host array in the reference's dtype (zero-copy reinterpretation)."""
    return dev.to_numpy(codes_t).view(quant_type.np_dtype)


def _zp_to_numpy(zp_t: torch.Tensor, quant_type: QuantType, zp_dtype) -> np.ndarray:
    z = _codes_to_numpy(zp_t, quant_type)
    zp_dtype = np.dtype(zp_dtype) if zp_dtype is not None else quant_type.np_dtype
    return z if z.dtype == zp_dtype else z.astype(zp_dtype)


def _zp_to_device_bytes(zero_point, quant_type: QuantType, count: int) -> torch.Tensor:
    """Zero points in any integer dtype → the byte representation the kernels read."""
    z = np.asarray(zero_point)
    z = np.broadcast_to(z.astype(np.int64).reshape(-1) if z.size > 1 else z.astype(np.int64).reshape(1),
                        (count,))
    mask = 0xF if quant_type.bitwidth == 4 else 0xFF
    return torch.from_numpy((z & mask).astype(np.uint8)).to(dev.require_cuda())


class _RowView:
    """How a 'rows' array of the reference (one (scale, zp) per row) maps onto a (K,N) weight.

    The summation order of the MSE error follows the memory layout NumPy would see:
      * C-contiguous rows (the group reshape copy, or any plain 2-D array)  → pairwise per row:
        flatten to a (R*C, 1) weight with group size C;
      * an F-ordered view (``W.T``)                                           → sequential in k:
        the (K,N) weight is the transpose, channel strategy.
    """

    def __init__(self, array: np.ndarray, strategy: QuantizationStrategy):
        a = np.asarray(array)
        self.shape = a.shape
        self.tensor = strategy == QuantizationStrategy.TENSOR
        if self.tensor:
            self.weight = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 1)
            self.strategy, self.gs = "tensor", -1
        elif a.ndim == 2 and not a.flags.c_contiguous and a.T.flags.c_contiguous:
            self.weight = a.T.astype(np.float32, copy=False)
            self.strategy, self.gs = "channel", -1
        else:
            a2 = np.ascontiguousarray(a.reshape(a.shape[0], -1), dtype=np.float32)
            self.weight = a2.reshape(-1, 1)
            self.strategy, self.gs = "group", a2.shape[1]
        self.rows = 1 if self.tensor else a.shape[0]

    def device_weight(self) -> torch.Tensor:
        return dev.to_device_f32(self.weight)

    def codes_to_rows(self, codes: np.ndarray) -> np.ndarray:
        if self.strategy == "channel":
            return codes.T
        return codes.reshape(self.shape)

    def per_row(self, values: np.ndarray) -> np.ndarray:
        return values.reshape(()) if self.tensor else values.reshape(self.rows, 1)


# ------------------------------------------------------------------------------------------------
# A2 / A6: ranges
# ------------------------------------------------------------------------------------------------
def _compute_min_max(array, strategy, group_size=-1, clip_ratio=1.0):
    """(min*clip, max*clip) per row with zero included (reference utils.py:42-69)."""
    strategy = coerce_strategy(strategy)
    view = _RowView(array, strategy)
    lo, hi = D.row_ranges(view.device_weight(), QuantType.QInt8, view.strategy, view.gs,
                          clip_ratio=clip_ratio, mse=False)
    return view.per_row(lo.cpu().numpy()), view.per_row(hi.cpu().numpy())


def _compute_min_max_mse(array, quant_type, strategy, group_size, is_symmetric, reduce_range,
                         scale_dtype, zp_dtype, maxshrink=0.20, patience=5, grid=100.0, norm=2.4):
    """Best shrunk (min, max) per row (reference utils.py:140-239).

    The kernels implement the reference's *defaults* (maxshrink 0.20, patience 5, grid 100,
    norm 2.4 — the only values any caller of the reference passes); other values raise.
    """
    if (maxshrink, patience, grid, norm) != (0.20, 5, 100.0, 2.4):
        raise NotImplementedError(
            "the device MSE search implements maxshrink=0.20, patience=5, grid=100, norm=2.4")
    quant_type = QuantType.coerce(quant_type)
    view = _RowView(array, strategy)
    lo, hi = D.row_ranges(view.device_weight(), quant_type, view.strategy, view.gs, is_symmetric,
                          reduce_range, 1.0, mse=True)
    return view.per_row(lo.cpu().numpy()), view.per_row(hi.cpu().numpy())


# ------------------------------------------------------------------------------------------------
# A3: scale / zero point from ranges
# ------------------------------------------------------------------------------------------------
def _compute_qparams(rmin, rmax, quant_type, is_symmetric, reduce_range, scale_dtype, zp_dtype):
    """(scale, zero_point) with the shape of ``rmin`` (reference utils.py:242-299)."""
    quant_type = QuantType.coerce(quant_type)
    rmin, rmax = np.asarray(rmin), np.asarray(rmax)
    shape = np.broadcast(rmin, rmax).shape
    lo = dev.to_device_f32(np.ascontiguousarray(np.broadcast_to(rmin, shape)).reshape(-1))
    hi = dev.to_device_f32(np.ascontiguousarray(np.broadcast_to(rmax, shape)).reshape(-1))
    scale, zp = D.qparams(lo, hi, quant_type, is_symmetric, reduce_range)
    scale = scale.cpu().numpy().reshape(shape).astype(scale_dtype, copy=False)
    return scale, _zp_to_numpy(zp, quant_type, zp_dtype).reshape(shape)


def _compute_qparams_from_array(array, quant_type, strategy, group_size, is_symmetric,
                                reduce_range, clip_ratio, mse, scale_dtype, zp_dtype):
    """Ranges (A2, or A6 when ``mse``) then A3 — reference utils.py:302-348."""
    quant_type = QuantType.coerce(quant_type)
    view = _RowView(array, strategy)
    _, scale, zp = D.rtn_quantize(view.device_weight(), quant_type, view.strategy, view.gs,
                                  is_symmetric, reduce_range, clip_ratio, mse)
    scale = view.per_row(scale.cpu().numpy()).astype(scale_dtype, copy=False)
    return scale, view.per_row(_zp_to_numpy(zp, quant_type, zp_dtype))


# ------------------------------------------------------------------------------------------------
# A4 / A5: quantize, dequantize
# ------------------------------------------------------------------------------------------------
def _quantize_array_from_qparams(array, scale, zero_point, quant_type, is_symmetric, reduce_range):
    """clip(int32(rint(x / scale)) + zp) in ``quant_type`` (reference utils.py:72-79).

    ``scale`` / ``zero_point`` are scalars or one value per row of ``array``.
    """
    quant_type = QuantType.coerce(quant_type)
    a = np.asarray(array)
    per_row = np.size(scale) > 1
    if a.ndim == 1 and per_row:   # one value per element: a column of single-element rows
        a = a.reshape(-1, 1)
    view = _RowView(a, QuantizationStrategy.CHANNEL if per_row else QuantizationStrategy.TENSOR)
    if quant_type in (QuantType.QInt32, QuantType.QUInt32):
        raise NotImplementedError("32-bit codes are produced by _quantize_bias only")
    s = dev.to_device_f32(np.ascontiguousarray(np.asarray(scale, dtype=np.float32)).reshape(-1))
    z = _zp_to_device_bytes(zero_point, quant_type, view.rows)
    if s.numel() != view.rows:
        s = s.expand(view.rows).contiguous()
    codes = D.quantize_with_qparams(view.device_weight(), s, z, quant_type, view.strategy, view.gs,
                                    is_symmetric, reduce_range)
    return view.codes_to_rows(_codes_to_numpy(codes, quant_type)).reshape(np.asarray(array).shape)


def _dequantize_array(q_array, scale, zero_point, *, preprocess=False, strategy=None, group_size=-1):
    """(f32(q) − f32(zp)) · scale (reference utils.py:102-137)."""
    q = np.asarray(q_array)
    quant_type = _quant_type_of(q.dtype)
    codes = torch.from_numpy(np.ascontiguousarray(q).view(np.uint8).copy())
    if preprocess:
        assert strategy is not None, "strategy must be provided if preprocess is True"
        k, n = q.shape
        st = strategy.value
        rows = D.num_rows(k, n, D._strategy(st), group_size)
        gs = group_size
    else:   # scalar parameters, or one per row of q
        per_row = np.size(scale) > 1
        if q.ndim == 1 and per_row:
            codes = codes.reshape(-1, 1)
        if per_row:   # rows of q are parameter rows → channel strategy on the transpose
            codes = codes.reshape(codes.shape[0], -1).t().contiguous()
            st, rows, gs = "channel", codes.shape[1], -1
        else:
            codes = codes.reshape(-1, 1)
            st, rows, gs = "tensor", 1, -1
    codes = codes.to(dev.require_cuda())
    s = dev.to_device_f32(np.ascontiguousarray(np.asarray(scale, dtype=np.float32)).reshape(-1))
    if s.numel() != rows:
        s = s.expand(rows).contiguous()
    if np.asarray(zero_point).dtype.kind == "f":      # HQQ: float zero points (hqq.py:77)
        z = dev.to_device_f32(np.ascontiguousarray(np.asarray(zero_point, dtype=np.float32)).reshape(-1))
        if z.numel() != rows:
            z = z.expand(rows).contiguous()
    else:
        z = _zp_to_device_bytes(zero_point, quant_type, rows)
    out = D.dequantize(codes.contiguous(), s, z, quant_type, st, gs)
    if preprocess:
        return dev.to_numpy(out)
    if st == "channel":
        out = out.t()
    return dev.to_numpy(out).reshape(q.shape)


def _fake_quantize_array(array, scale, zero_point, quant_type, is_symmetric, reduce_range):
    """Quantize then dequantize (reference utils.py:82-99)."""
    q = _quantize_array_from_qparams(array, scale, zero_point, quant_type, is_symmetric, reduce_range)
    return _dequantize_array(q, np.asarray(scale), np.asarray(zero_point))


def _quant_type_of(dtype) -> QuantType:
    for qt in (QuantType.QInt4, QuantType.QUInt4, QuantType.QInt8, QuantType.QUInt8):
        if np.dtype(dtype) == qt.np_dtype:
            return qt
    raise TypeError(f"not a quantized weight dtype: {dtype}")
