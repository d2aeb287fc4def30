"""GPTQ weight quantization on the GPU: the ``GPTQConfig`` plugin and its array-level entry
points, mirroring the reference's ``core/_algorithms/gptq.py`` (``GPTQConfig`` :34-73, ``_gptq``
:76-243, ``_accumulate_hessian`` :246-260, ``_gptq_quantize`` :263-324)."""
from __future__ import annotations

__all__ = ["GPTQConfig", "_gptq_quantize", "_gptq", "_accumulate_hessian", "calibration_cache"]

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
          precision="bf16x3", factor=None):
    """GPTQ of one (K,N) weight given its (K,K) Hessian → ``(codes, scale, zero_point)``.

    Shapes and dtypes are the reference's (gptq.py:76-243).  ``W`` / ``H`` may be NumPy arrays or
    float32 CUDA tensors; neither is modified.  ``factor`` (a ``gptq_device.HinvFactor`` of this
    ``H`` for the same percdamp / actorder) skips the factorization.
    """
    from onnx_quantize_b200 import gptq_device as G

    strategy = coerce_strategy(strategy)
    quant_type = QuantType.coerce(quant_type)
    w = dev.to_device_f32(W)
    h = dev.to_device_f32(H)
    if w.dim() != 2 or h.shape != (w.shape[0], w.shape[0]):
        raise ValueError("W must be (K,N) and H (K,K)")
    f = factor if factor is not None else G.hinv_cholesky_upper(h, percdamp=percdamp, actorder=actorder,
                                                                precision=precision)
    codes, scale, zp = G.gptq_quantize(w, f, quant_type, strategy, group_size, is_symmetric,
                                       reduce_range, clip_ratio, mse, block_size, mode, precision)
    if f.marginal:   # a pivot too small for the tensor-core precision to decide: redo the factor in fp32
        f = G.resolve_marginal(f, h, percdamp, actorder)
        codes, scale, zp = G.gptq_quantize(w, f, quant_type, strategy, group_size, is_symmetric,
                                           reduce_range, clip_ratio, mse, block_size, mode, precision)
    if not f.ok:
        logger.warning(_FALLBACK_WARNING)
    codes_np = _codes_to_numpy(codes, quant_type)
    scale_np = scale.cpu().numpy().astype(np.float32, copy=False)
    zp_np = _zp_to_numpy(zp, quant_type, codes_np.dtype)   # gptq.py:240 zp.astype(Q_int.dtype)
    scale_np, zp_np = _shape_like_reference(scale_np, zp_np, strategy)
    return codes_np, scale_np, zp_np


class _CalibrationCache:
    """Device Hessians (and their inverse factors) of the calibration arrays seen last.

    The reference hands the SAME array object to every node that reads one activation —
    ``_set_qparams_gptq`` stores ``collected_outputs[node.inputs[0].name]`` in ``node.meta["input"]``
    (core/_calibration/calibrate.py:296-307), so q/k/v (and gate/up) share their input — and then
    recomputes ``XᵀX`` and the three LAPACK factorizations once per node (gptq.py:263-324).  Here an
    entry is keyed by the array's identity (a weak reference keeps the id from being recycled) plus
    a sampled fingerprint (in-place edits), so the 4.3 GB upload, the Hessian and the factor happen
    once per distinct input; the factor additionally depends on (percdamp, actorder, precision).
    Two entries are kept (a K = 14336 entry is 1.6 GB of HBM)."""

    max_entries = 2

    def __init__(self):
        from collections import OrderedDict

        self.entries: "OrderedDict[tuple, dict]" = OrderedDict()
        self.hits = self.misses = self.factor_hits = self.factor_misses = 0

    def _key(self, inputs, precision: str):
        from onnx_quantize_b200.parallel.prequantized import weight_fingerprint

        return (id(inputs), precision, weight_fingerprint(inputs))

    def hessian(self, inputs, k: int, precision: str) -> dict:
        import weakref

        if not isinstance(inputs, np.ndarray):          # tensors / lists: no identity to rely on
            self.misses += 1
            return {"h": _hessian_from_host(inputs, k, precision), "factors": {}}
        key = self._key(inputs, precision)
        entry = self.entries.get(key)
        if entry is not None and entry["ref"]() is inputs:
            self.entries.move_to_end(key)
            self.hits += 1
            return entry
        self.misses += 1
        while len(self.entries) >= self.max_entries:
            self.entries.popitem(last=False)
        entry = {"h": _hessian_from_host(inputs, k, precision), "factors": {},
                 "ref": weakref.ref(inputs, lambda _r, key=key: self.entries.pop(key, None))}
        self.entries[key] = entry
        return entry

    def factor(self, entry: dict, percdamp: float, actorder: bool, precision: str):
        from onnx_quantize_b200 import gptq_device as G

        fkey = (float(percdamp), bool(actorder), precision)
        f = entry["factors"].get(fkey)
        if f is None:
            self.factor_misses += 1
            f = G.hinv_cholesky_upper(entry["h"], percdamp=percdamp, actorder=actorder, precision=precision)
            f = entry["factors"][fkey] = G.resolve_marginal(f, entry["h"], percdamp, actorder)
        else:
            self.factor_hits += 1
        return f

    def clear(self) -> None:
        self.entries.clear()


calibration_cache = _CalibrationCache()

_HESSIAN_CHUNK_BYTES = 256 << 20


def _hessian_from_host(inputs, k: int, precision: str):
    """``(2/n)·XᵀX`` on the device for host activations ``(n samples, ..., K)`` (gptq.py:246-260 with
    an empty running Hessian).  The array is streamed: 256 MB chunks of whole token rows go up through
    the pinned staging path (``_device``) into two alternating device buffers and are folded as they
    land — the activations (4.3 GB per Llama-3-8B projection input, 15 GB for down_proj) are never
    resident as a whole."""
    import torch

    from onnx_quantize_b200.hessian import hessian_accumulate

    device = dev.require_cuda()
    h = torch.empty((k, k), dtype=torch.float32, device=device)
    if isinstance(inputs, torch.Tensor) and inputs.is_cuda:
        hessian_accumulate(inputs.to(torch.float32), h, alpha=2.0 / int(inputs.shape[0]), beta=0.0, precision=precision)
        return h
    a = np.asarray(inputs) if not isinstance(inputs, torch.Tensor) else inputs.numpy()
    n = int(a.shape[0])
    if a.dtype != np.float32:
        a = a.astype(np.float32)
    x = np.ascontiguousarray(a).reshape(-1, k)
    rows = max(1024, (_HESSIAN_CHUNK_BYTES // (4 * k)) // 1024 * 1024)
    if x.shape[0] <= rows:
        hessian_accumulate(dev.to_device_f32(x), h, alpha=2.0 / n, beta=0.0, precision=precision)
        return h
    bufs = [torch.empty((rows, k), dtype=torch.float32, device=device) for _ in range(2)]
    for i, r0 in enumerate(range(0, x.shape[0], rows)):
        m = min(rows, x.shape[0] - r0)
        dst = bufs[i & 1][:m]                       # same stream: the kernel that last read it is done
        dev.upload_into(torch.from_numpy(x[r0:r0 + m]), dst)
        hessian_accumulate(dst, h, alpha=2.0 / n, beta=0.0 if i == 0 else 1.0, precision=precision)
    return h


def _gptq_quantize(weights, inputs, quant_type=QuantType.QInt8,
                   strategy=QuantizationStrategy.CHANNEL, group_size=32, is_symmetric=False,
                   reduce_range=False, clip_ratio=1.0, block_size=128, percdamp=0.01,
                   actorder=False, mse=False, scale_dtype=np.float32, zp_dtype=np.int8,
                   mode="reference", precision="bf16x3"):
    """Hessian from ``inputs`` (num_samples, ..., in_features), then :func:`_gptq`
    (gptq.py:263-324).  The Hessian never leaves the device; Hessian and inverse factor are shared
    between calls that receive the same ``inputs`` array (``calibration_cache``)."""
    k = int(np.shape(weights)[0])
    if int(np.shape(inputs)[-1]) != k:
        raise ValueError(f"inputs have {np.shape(inputs)[-1]} features, the weight has {k} input channels")
    entry = calibration_cache.hessian(inputs, k, precision)
    f = calibration_cache.factor(entry, percdamp, actorder, precision)
    return _gptq(weights, entry["h"], quant_type=quant_type, strategy=strategy, group_size=group_size,
                 is_symmetric=is_symmetric, reduce_range=reduce_range, clip_ratio=clip_ratio,
                 block_size=block_size, percdamp=percdamp, actorder=actorder, mse=mse,
                 scale_dtype=scale_dtype, zp_dtype=zp_dtype, mode=mode, precision=precision, factor=f)
