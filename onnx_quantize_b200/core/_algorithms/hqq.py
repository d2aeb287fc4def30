"""HQQ (half-quadratic quantization) on the GPU: the ``HqqConfig`` plugin and ``_hqq_quantize``,
mirroring the reference's ``core/_algorithms/hqq.py`` (``HqqConfig`` :27-96, ``_shrink_op``
:103-104, ``_optimize_zero_point`` :107-146, ``_hqq_quantize`` :149-217).  The whole optimisation
runs in one C call (``b200q_hqq_quantize``, csrc/hqq.cu)."""
from __future__ import annotations

__all__ = ["HqqConfig", "_hqq_quantize"]

from typing import TYPE_CHECKING, Literal

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._algorithms.utils import _codes_to_numpy
from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.core._qconfig import (
    AlgorithmConfig,
    QuantizationStrategy,
    register_algorithm_config,
)

if TYPE_CHECKING:  # pragma: no cover
    import onnx_ir as ir

    from onnx_quantize_b200.core._qconfig import QConfig, QWeightArgs


@register_algorithm_config
class HqqConfig(AlgorithmConfig):
    """HQQ settings: ``lp_norm`` of the shrinkage operator, ``beta`` and its growth ``kappa``,
    number of iterations, early stop on the first non-improving iteration."""

    algorithm_type: Literal["hqq"] = "hqq"
    lp_norm: float = 0.7
    beta: float = 1e1
    kappa: float = 1.01
    iters: int = 20
    early_stop: bool = True

    @staticmethod
    def _check_hqq_constraints(dtype: QuantType, symmetric: bool, strategy: QuantizationStrategy,
                               group_size: int) -> None:
        if dtype != QuantType.QUInt4:
            raise ValueError(f"HQQ only supports uint4 weight type. Found: {dtype}")
        if symmetric:
            raise ValueError("HQQ only supports asymmetric quantization.")
        if strategy != QuantizationStrategy.GROUP:   # MatMulNBits, its only consumer, wants groups
            raise ValueError(f"HQQ only supports 'group' quantization strategy. Found: {strategy}")
        if group_size != -1 and (group_size < 16 or (group_size & (group_size - 1)) != 0):
            raise ValueError("HQQ requires group_size to be greater than 16 and a power of 2. "
                             f"Found: {group_size}")

    def validate_weight_args(self, weight_args: "QWeightArgs") -> None:
        self._check_hqq_constraints(weight_args.dtype, weight_args.symmetric, weight_args.strategy,
                                    weight_args.group_size)
        weight_args.zp_dtype = weight_args.scale_dtype      # HQQ zero points are floats

    def quantize_weights(self, w: "ir.Value", qconfig: "QConfig", out: "ir.Value | None" = None
                         ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        wa = qconfig.weights
        return _hqq_quantize(w.const_value.numpy(), quant_type=wa.dtype, group_size=wa.group_size,
                             reduce_range=wa.reduce_range, clip_ratio=wa.clip_ratio, mse=wa.mse,
                             scale_dtype=wa.scale_dtype, zp_dtype=wa.zp_dtype, lp_norm=self.lp_norm,
                             beta=self.beta, kappa=self.kappa, iters=self.iters, early_stop=self.early_stop)


def hqq_quantize_device(w: torch.Tensor, group_size: int, reduce_range: bool = False, clip_ratio: float = 1.0,
                        mse: bool = False, lp_norm: float = 0.7, beta: float = 1e1, kappa: float = 1.01,
                        iters: int = 20, early_stop: bool = True, return_info: bool = False):
    """CUDA tensors in and out: ``(codes uint8 (K,N), scale f32 (rows,), zp f32 (rows,))`` and, with
    ``return_info``, ``(best_iter int32[1], errors f64[iters])``."""
    lib = _lib.load()
    k, n = D._check_weight(w)
    gs, g = D.resolve_group(k, 2, group_size)
    rows = n * g
    codes = torch.empty((k, n), dtype=torch.uint8, device=w.device)
    scale = torch.empty((rows,), dtype=torch.float32, device=w.device)
    zp = torch.empty((rows,), dtype=torch.float32, device=w.device)
    best = torch.empty((1,), dtype=torch.int32, device=w.device)
    errors = torch.empty((max(int(iters), 1),), dtype=torch.float64, device=w.device)
    mse_mode = D._mse_mode(mse)
    ws = dev.workspace(lib.b200q_hqq_workspace_bytes(k, n, int(group_size), mse_mode, int(iters)))
    rc = lib.b200q_hqq_quantize(w.data_ptr(), k, n, 1, int(group_size), int(bool(reduce_range)), float(clip_ratio),
                                mse_mode, float(lp_norm), float(beta), float(kappa), int(iters),
                                int(bool(early_stop)), codes.data_ptr(), scale.data_ptr(), zp.data_ptr(),
                                best.data_ptr(), errors.data_ptr(), ws.data_ptr(), ws.numel(), dev.stream_ptr())
    _lib.check(rc, "b200q_hqq_quantize")
    if return_info:
        return codes, scale, zp, (best, errors[:int(iters)])
    return codes, scale, zp


def _hqq_quantize(w_f: np.ndarray, quant_type: QuantType, group_size: int, reduce_range: bool = False,
                  clip_ratio: float = 1.0, mse: bool = False, scale_dtype: np.dtype = np.float32,
                  zp_dtype: np.dtype = np.float32, lp_norm: float = 0.7, beta: float = 1e1,
                  kappa: float = 1.01, iters: int = 20, early_stop: bool = True
                  ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(K,N) float32 weight → ``(codes (K,N) uint4, scale (N*G,1), zero_point (N*G,1) float)``."""
    assert zp_dtype == scale_dtype            # hqq.py:178
    quant_type = QuantType.coerce(quant_type)
    if quant_type != QuantType.QUInt4:
        raise ValueError(f"HQQ only supports uint4 weight type. Found: {quant_type}")
    w = dev.to_device_f32(w_f)
    if w.dim() != 2:
        raise ValueError("weights must be 2-D (in_channels, out_channels)")
    codes, scale, zp = hqq_quantize_device(w, group_size, reduce_range, clip_ratio, mse, lp_norm, beta, kappa,
                                           iters, early_stop)
    scale_np = scale.cpu().numpy().astype(scale_dtype, copy=False).reshape(-1, 1)
    zp_np = zp.cpu().numpy().astype(zp_dtype, copy=False).reshape(-1, 1)
    return _codes_to_numpy(codes, quant_type), scale_np, zp_np
