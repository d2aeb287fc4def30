"""The call site of the weight-quantization plugin and the MatMulNBits repack ("layout B").

Mirrors the array-level part of the reference's ``qrules/_common.py``: ``_resolve_group_size``
(:13-29), ``is_matmul_nbits_compatible`` (:32-62), ``_prepare_for_matmul_nbits`` (:65-123) and
``quantize_weights`` (:126-144).  The graph rewriting that calls them is the reference's own and
stays untouched; ``quantize_weights`` needs onnx_ir (to wrap arrays as initializers) and raises
ImportError without it.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.core._qconfig import QConfig, QuantizationStrategy

logger = logging.getLogger(__name__)


def _resolve_group_size(w, group_size: int) -> int:
    """group sizes that exceed or do not divide in_channels collapse to in_channels."""
    k = w.const_value.numpy().shape[0]
    if group_size:
        note = f"Adjusting group size from {group_size} to {k} for weight '{w.name}'"
        if group_size > k:
            logger.debug(note + f" as it exceeds the number of input channels {k}.")
            group_size = k
        if k % group_size != 0:
            logger.debug(note + f" as it does not divide the number of input channels {k}.")
            group_size = k
    return group_size


def is_matmul_nbits_compatible(qconfig: QConfig, name: str = "") -> bool:
    """weights-only + uint4/uint8 + group strategy + power-of-two group size >= 16."""
    why = f"Found uncompatibility for MatMulNBits in {name}: "
    if qconfig.input_activations is not None or qconfig.output_activations is not None:
        logger.debug(why + "It only supports weight-only quantization.")
        return False
    if qconfig.weights.dtype not in (QuantType.QUInt4, QuantType.QUInt8):
        logger.debug(why + f"It only supports uint4 and uint8 weight types. Found: "
                           f"{qconfig.weights.dtype}")
        return False
    if qconfig.weights.strategy != QuantizationStrategy.GROUP:
        logger.debug(why + "It only supports 'group' quantization strategy. Found: "
                     + str(qconfig.weights.strategy))
        return False
    gs = qconfig.weights.group_size
    if gs != -1 and (gs < 16 or gs & (gs - 1)):
        logger.debug(why + "group_size should be a power of 2 greater than or equal to 16.")
        return False
    return True


def _prepare_for_matmul_nbits(w_q: np.ndarray, w_scale: np.ndarray, w_zero_point: np.ndarray,
                              qconfig: QConfig) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(K,N) codes + per-row params → MatMulNBits operands ``(B, scales, zero_points)``.

    B is (N, K/gs, gs*bits/8) uint8 with ``B[n,g,j] = q[g*gs+2j, n] | q[g*gs+2j+1, n] << 4`` for
    4-bit; scales (N, K/gs) float32; zero points nibble-packed per output channel (low nibble =
    even block, odd block count padded with 0x8) unless there is a single block or the zero point
    is float (HQQ), in which case they are only reshaped.
    """
    k, n = w_q.shape
    gs = qconfig.weights.group_size
    bits = qconfig.weights.dtype.bitwidth
    assert k % gs == 0
    g = k // gs
    device = dev.require_cuda()
    q = np.asarray(w_q)
    if q.dtype in (QuantType.QInt4.np_dtype, QuantType.QUInt4.np_dtype):
        q = q.view(np.uint8)
    codes = torch.from_numpy(np.ascontiguousarray(q).astype(np.uint8, copy=False).copy()).to(device)
    float_zp = qconfig.weights.zp_dtype == w_scale.dtype
    z = np.asarray(w_zero_point)
    if float_zp:
        zp_bytes = torch.zeros((n * g,), dtype=torch.uint8, device=device)
    else:
        if z.dtype in (QuantType.QInt4.np_dtype, QuantType.QUInt4.np_dtype):
            z = z.view(np.uint8)
        zp_bytes = torch.from_numpy(
            np.ascontiguousarray(z).astype(np.uint8, copy=False).reshape(-1).copy()).to(device)
    b, zp = D.pack_matmul_nbits(codes, zp_bytes, gs, bits)
    scales = w_scale.reshape(-1, g)
    if float_zp:   # HQQ keeps float zero points un-packed (not on this package's hot path)
        return dev.to_numpy(b), scales, np.reshape(z, (n, -1)).astype(qconfig.weights.zp_dtype)
    return dev.to_numpy(b), scales, zp.cpu().numpy()


def quantize_weights(op, w, qconfig: QConfig, out=None, is_matmul_nbits_compatible: bool = False):
    """Run the configured algorithm plugin on ``w`` and emit the three initializers."""
    try:
        import onnx_ir as ir
    except ImportError as e:  # pragma: no cover - needs the ONNX stack
        raise ImportError("quantize_weights() builds ONNX initializers and needs onnx_ir") from e

    w_q, w_scale, w_zp = qconfig.weights.algorithm.quantize_weights(w, qconfig, out=out)
    if is_matmul_nbits_compatible:
        w_q, w_scale, w_zp = _prepare_for_matmul_nbits(w_q, w_scale, w_zp, qconfig)
    return (op.initializer(ir.tensor(w_q), name=w.name),
            op.initializer(ir.tensor(w_scale), name=f"{w.name}/scale"),
            op.initializer(ir.tensor(w_zp), name=f"{w.name}/zero_point"))
