"""GPTQ weight quantization on the GPU: the ``GPTQConfig`` plugin and its array-level entry
points, mirroring the reference's ``core/_algorithms/gptq.py`` (``GPTQConfig`` :34-73, ``_gptq``
:76-243, ``_accumulate_hessian`` :246-260, ``_gptq_quantize`` :263-324)."""
from __future__ import annotations

__all__ = ["GPTQConfig", "_gptq_quantize"]

import logging
from typing import TYPE_CHECKING, ClassVar, Literal

import numpy as np

from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.core._qconfig import (
    AlgorithmConfig,
    QuantizationStrategy,
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

    def quantize_weights(self, w: "ir.Value", qconfig: "QConfig", out: "ir.Value | None" = None
                         ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        assert out is not None, "Output value is required for GPTQ quantization."
        node = out.producer()
        assert "input" in node.meta, "GPTQ requires calibration data in node meta."
        from onnx_quantize_b200.parallel import prequantized

        cached = prequantized.lookup(w)
        if cached is not None:
            return cached
        wa = qconfig.weights
        return _gptq_quantize(w.const_value.numpy(), node.meta["input"], quant_type=wa.dtype,
                              strategy=wa.strategy, is_symmetric=wa.symmetric,
                              reduce_range=wa.reduce_range, clip_ratio=wa.clip_ratio,
                              block_size=self.block_size, percdamp=self.percdamp,
                              group_size=wa.group_size, actorder=self.actorder, mse=wa.mse,
                              scale_dtype=wa.scale_dtype, zp_dtype=wa.zp_dtype, mode=self.mode)


def _accumulate_hessian(inp, H, num_samples):
    raise NotImplementedError("GPTQ device path is being built")


def _gptq(W, H, quant_type, strategy, group_size, is_symmetric, reduce_range, clip_ratio,
          block_size, percdamp, actorder, mse, scale_dtype, zp_dtype, mode="reference"):
    raise NotImplementedError("GPTQ device path is being built")


def _gptq_quantize(weights, inputs, quant_type=QuantType.QInt8,
                   strategy=QuantizationStrategy.CHANNEL, group_size=32, is_symmetric=False,
                   reduce_range=False, clip_ratio=1.0, block_size=128, percdamp=0.01,
                   actorder=False, mse=False, scale_dtype=np.float32, zp_dtype=np.int8,
                   mode="reference"):
    raise NotImplementedError("GPTQ device path is being built")
