"""GPTQ weight quantization on the GPU: the ``GPTQConfig`` plugin and its array-level entry
points, mirroring the reference's ``core/_algorithms/gptq.py`` (``GPTQConfig`` :34-73, ``_gptq``
:76-243, ``_accumulate_hessian`` :246-260, ``_gptq_quantize`` :263-324)."""
from __future__ import annotations

__all__ = ["GPTQConfig", "_gptq_quantize", "_gptq", "_accumulate_hessian"]

import logging
from typing import TYPE_CHECKING, ClassVar, Literal

import numpy as np

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200.core._algorithms.rtn import _shape_like_reference
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

logger = logging.getLogger(__name__)


@register_algorithm_config
class GPTQConfig(AlgorithmConfig):
    """GPTQ parameters: ``block_size`` (128), ``percdamp`` (0.01), ``actorder`` (False).

    ``mode`` is an extension: "reference" reproduces the reference's update rule exactly as
    written (no error propagation, SURVEY.md finding 3); "propagate" is GPTQ as published.
    """

    requires_calibration: ClassVar[bool] = True

    algorithm_type: Literal["gptq"] = "gptq"
    block_size: int = 128
    percdamp: float = 0.01
    actorder: bool = False
    mode: Literal["reference", "propagate"] = "reference"
    precision: Literal["tf32", "tf32x3", "bf16x3", "fp32"] = "bf16x3"

    def quantize_weights(self, w: "ir.Value", qconfig: "QConfig", out: "ir.Value | None" = None
                         ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        assert out is not None, "Output value is required for GPTQ quantization."
        node = out.producer()
        assert "input" in node.meta, "GPTQ requires calibration data in node meta."
        wa = qconfig.weights
        return _gptq_quantize(w.const_value.numpy(), node.meta["input"], quant_type=wa.dtype,
                              strategy=wa.strategy, is_symmetric=wa.symmetric,
                              reduce_range=wa.reduce_range, clip_ratio=wa.clip_ratio,
                              block_size=self.block_size, percdamp=self.percdamp,
                              group_size=wa.group_size, actorder=self.actorder, mse=wa.mse,
                              scale_dtype=wa.scale_dtype, zp_dtype=wa.zp_dtype, mode=self.mode,
                              precision=self.precision)


_FALLBACK_WARNING = (
    "Failed to invert hessian due to numerical instability. Consider "
    "increasing percdamp, increasing the number "
    "of calibration samples, or shuffling the calibration dataset. "
    "Falling back to round-to-nearest for this module."
)   # the reference's message (gptq.py:144-149)


def _accumulate_hessian(inp, H, num_samples, precision="bf16x3"):
    """``H ← H·n/(n+b) + (2/(n+b))·XᵀX`` with ``b = inp.shape[0]`` samples (gptq.py:246-260).

    ``H`` may be a NumPy array (staged to the device and back, returned as NumPy like the
    reference) or a float32 CUDA tensor (updated in place, returned as is).
    """
    import torch

    from onnx_quantize_b200.hessian import hessian_accumulate

    added = int(inp.shape[0])
    total = num_samples + added
    x = dev.to_device_f32(inp)
    on_device = isinstance(H, torch.Tensor)
    h = H if on_device else dev.to_device_f32(np.array(H, dtype=np.float32, copy=True))
    hessian_accumulate(x, h, alpha=2.0 / total, beta=num_samples / total, precision=precision)
    return (h if on_device else h.cpu().numpy()), total


def _gptq(W, H, quant_type, strategy, group_size, is_symmetric, reduce_range, clip_ratio,
          block_size, percdamp, actorder, mse, scale_dtype, zp_dtype, mode="reference",
          precision="bf16x3"):
    """GPTQ of one (K,N) weight given its (K,K) Hessian → ``(codes, scale, zero_point)``.

    Shapes and dtypes are the reference's (gptq.py:76-243).  ``W`` / ``H`` may be NumPy arrays or
    float32 CUDA tensors; neither is modified.
    """
    from onnx_quantize_b200 import gptq_device as G

    strategy = coerce_strategy(strategy)
    quant_type = QuantType.coerce(quant_type)
    w = dev.to_device_f32(W)
    h = dev.to_device_f32(H)
    if w.dim() != 2 or h.shape != (w.shape[0], w.shape[0]):
        raise ValueError("W must be (K,N) and H (K,K)")
    f = G.hinv_cholesky_upper(h, percdamp=percdamp, actorder=actorder, precision=precision)
    codes, scale, zp = G.gptq_quantize(w, f, quant_type, strategy, group_size, is_symmetric,
                                       reduce_range, clip_ratio, mse, block_size, mode, precision)
    if not f.ok:
        logger.warning(_FALLBACK_WARNING)
    codes_np = _codes_to_numpy(codes, quant_type)
    scale_np = scale.cpu().numpy().astype(np.float32, copy=False)
    zp_np = _zp_to_numpy(zp, quant_type, codes_np.dtype)   # gptq.py:240 zp.astype(Q_int.dtype)
    scale_np, zp_np = _shape_like_reference(scale_np, zp_np, strategy)
    return codes_np, scale_np, zp_np


def _gptq_quantize(weights, inputs, quant_type=QuantType.QInt8,
                   strategy=QuantizationStrategy.CHANNEL, group_size=32, is_symmetric=False,
                   reduce_range=False, clip_ratio=1.0, block_size=128, percdamp=0.01,
                   actorder=False, mse=False, scale_dtype=np.float32, zp_dtype=np.int8,
                   mode="reference", precision="bf16x3"):
    """Hessian from ``inputs`` (num_samples, ..., in_features), then :func:`_gptq`
    (gptq.py:263-324).  The Hessian never leaves the device."""
    import torch

    k = int(np.shape(weights)[0])
    h = torch.zeros((k, k), dtype=torch.float32, device=dev.require_cuda())
    h, _ = _accumulate_hessian(inputs, h, 0, precision=precision)
    return _gptq(weights, h, quant_type=quant_type, strategy=strategy, group_size=group_size,
                 is_symmetric=is_symmetric, reduce_range=reduce_range, clip_ratio=clip_ratio,
                 block_size=block_size, percdamp=percdamp, actorder=actorder, mse=mse,
                 scale_dtype=scale_dtype, zp_dtype=zp_dtype, mode=mode, precision=precision)
