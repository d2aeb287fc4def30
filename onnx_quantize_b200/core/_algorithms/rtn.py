"""Round-to-nearest weight quantization on the GPU: the ``RTNConfig`` plugin and its array-level
entry points, mirroring the reference's ``core/_algorithms/rtn.py`` (``RTNConfig`` :28-51,
``_rtn_quantize`` :54-109, ``_quantize_bias`` :112-138)."""
from __future__ import annotations

__all__ = ["RTNConfig", "_rtn_quantize"]

from typing import TYPE_CHECKING, Literal

import numpy as np

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._algorithms.utils import _codes_to_numpy, _zp_to_numpy
from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.core._qconfig import (
    AlgorithmConfig,
    QuantizationStrategy,
    coerce_strategy,
    register_algorithm_config,
)

if TYPE_CHECKING:  # pragma: no cover
    import onnx_ir as ir

    from onnx_quantize_b200.core._qconfig import QConfig


@register_algorithm_config
class RTNConfig(AlgorithmConfig):
    """Round-to-nearest: the default algorithm; all settings come from ``QWeightArgs``."""

    algorithm_type: Literal["rtn"] = "rtn"

    def quantize_weights(self, w: "ir.Value", qconfig: "QConfig", out: "ir.Value | None" = None
                         ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        from onnx_quantize_b200.parallel import prequantized

        wa = qconfig.weights
        cached = prequantized.lookup(w, wa, self.algorithm_type)
        if cached is not None:   # filled by the multi-GPU pre-pass (SURVEY.md §8b "Threading")
            return cached
        return _rtn_quantize(w.const_value.numpy(), wa.dtype, strategy=wa.strategy,
                             group_size=wa.group_size, is_symmetric=wa.symmetric,
                             reduce_range=wa.reduce_range, clip_ratio=wa.clip_ratio, mse=wa.mse,
                             scale_dtype=wa.scale_dtype, zp_dtype=wa.zp_dtype)


def _shape_like_reference(scale: np.ndarray, zp: np.ndarray, strategy: QuantizationStrategy):
    """tensor → 0-d, channel → (N,), group → (N*G, 1) (reference rtn.py:101-104)."""
    if strategy == QuantizationStrategy.TENSOR:
        return scale.reshape(()), zp.reshape(())
    if strategy == QuantizationStrategy.CHANNEL:   # np.squeeze of (N,1): a single channel becomes 0-d
        return np.squeeze(scale.reshape(-1, 1)), np.squeeze(zp.reshape(-1, 1))
    return scale.reshape(-1, 1), zp.reshape(-1, 1)


def _rtn_quantize(array: np.ndarray, quant_type: QuantType, strategy: QuantizationStrategy,
                  group_size: int, is_symmetric: bool, reduce_range: bool, clip_ratio: float,
                  mse: bool, scale_dtype: np.dtype, zp_dtype: np.dtype
                  ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Quantize a (K,N) float32 weight → ``(codes (K,N), scale, zero_point)``.

    Host arrays in, host arrays out (one H2D copy of the weight, one D2H copy of codes + params);
    dtypes and shapes are the reference's: codes in ``quant_type.np_dtype`` (ml_dtypes int4/uint4
    are one byte per element), scale float32, zero point ``zp_dtype``.
    """
    strategy = coerce_strategy(strategy)
    quant_type = QuantType.coerce(quant_type)
    w = dev.to_device_f32(array)
    if w.dim() != 2:
        raise ValueError("weights must be 2-D (in_channels, out_channels)")
    codes, scale, zp = D.rtn_quantize(w, quant_type, strategy, group_size, is_symmetric,
                                      reduce_range, clip_ratio, mse)
    return _finalize_triple(codes, scale, zp, quant_type, strategy, scale_dtype, zp_dtype)


def _finalize_triple(codes, scale, zp, quant_type: QuantType, strategy: QuantizationStrategy,
                     scale_dtype, zp_dtype) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Raw kernel outputs — (K,N) code bytes, flat float32 scales, flat zero-point bytes, as CUDA
    tensors or host arrays — to what the reference returns (rtn.py:96-109): codes in
    ``quant_type.np_dtype`` (a reinterpretation of the bytes), scale in ``scale_dtype``, zero point
    in ``zp_dtype``, parameters shaped 0-d / (N,) / (N*G, 1)."""
    import torch

    if isinstance(codes, torch.Tensor):
        codes_np = _codes_to_numpy(codes, quant_type)
        scale_np = scale.cpu().numpy()
        zp_np = _zp_to_numpy(zp, quant_type, zp_dtype)
    else:
        codes_np = np.asarray(codes).view(quant_type.np_dtype)
        scale_np = np.asarray(scale)
        zp_np = np.asarray(zp).view(quant_type.np_dtype)
        want = np.dtype(zp_dtype) if zp_dtype is not None else quant_type.np_dtype
        if zp_np.dtype != want:
            zp_np = zp_np.astype(want)
    scale_np = scale_np.reshape(-1).astype(scale_dtype, copy=False)
    scale_np, zp_np = _shape_like_reference(scale_np, zp_np.reshape(-1), strategy)
    return codes_np, scale_np, zp_np


def _quantize_bias(bias, input_scale, weight_scale):
    """int32 bias with scale ``weight_scale * input_scale`` and zero point 0."""
    assert bias.ndim == 1
    assert bias.dtype == np.float32
    assert np.size(input_scale) == 1
    assert weight_scale.dtype == np.float32
    assert weight_scale.size == 1 or bias.size == weight_scale.size
    b = dev.to_device_f32(bias)
    ws = dev.to_device_f32(np.ascontiguousarray(weight_scale).reshape(-1))
    q, s = D.quantize_bias(b, float(np.asarray(input_scale).reshape(())), ws)
    return q.cpu().numpy(), s.cpu().numpy().reshape(np.shape(weight_scale)), 0
