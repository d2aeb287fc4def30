"""AWQ scale / clip search on the GPU — the numeric core of the reference's AWQ pre-pass
(``pre_passes/awq.py``: ``_compute_activation_scale`` :47-50, ``_compute_weight_scale`` :52-70, the
scale grid of ``_apply_awq`` :121-184 and the clip grid of ``_apply_awq_clip`` :207-254).

The reference scores each of its 20 (+10) candidates with two products over ALL calibration tokens.
Here the activations are contracted once, into the Gram matrix ``G = XᵀX`` (the tensor-core Hessian
kernel of the GPTQ path) and ``Σ|x|`` per channel — both streamed batch by batch, nothing is kept —
and a candidate costs one ``(K,K)·(K,N)`` product:  ``‖XW − XŴ‖² = Σ_n d_nᵀ G d_n``, ``D = W − Ŵ``.
The ONNX graph side of the pass (inserting the ``Mul`` node, replacing the initializer) stays the
reference's; this module returns the arrays it needs (``best_scale``, the updated weight, the clip
ratio).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.hessian import hessian_accumulate


class AwqStatistics:
    """Streaming calibration statistics of one layer input: ``G = Σ XᵀX``, ``Σ|x|``, token count."""

    def __init__(self, k: int, precision: str = "bf16x3", device=None):
        self.device = device or dev.require_cuda()
        self.gram = torch.zeros((k, k), dtype=torch.float32, device=self.device)
        self.abs_sum = torch.zeros((k,), dtype=torch.float32, device=self.device)
        self.tokens = 0
        self.precision = precision

    def add(self, inp) -> None:
        x = dev.to_device_f32(inp)
        k = int(x.shape[-1])
        x2 = x.reshape(-1, k)
        hessian_accumulate(x2, self.gram, alpha=1.0, beta=1.0, precision=self.precision)
        lib = _lib.load()
        _lib.check(lib.b200q_awq_abs_sum(x2.data_ptr(), int(x2.shape[0]), k, self.abs_sum.data_ptr(),
                                         dev.stream_ptr()), "b200q_awq_abs_sum")
        self.tokens += int(x2.shape[0])

    @property
    def activation_scale(self) -> torch.Tensor:
        return self.abs_sum / float(self.tokens)

    def rescaled(self, scale: torch.Tensor) -> "AwqStatistics":
        """Statistics of the inputs divided channel-wise by ``scale`` (awq.py:175: the node's
        calibration input after the scale has been folded into the weights)."""
        out = AwqStatistics.__new__(AwqStatistics)
        out.device, out.tokens, out.precision = self.device, self.tokens, self.precision
        out.gram = self.gram / (scale[:, None] * scale[None, :])
        out.abs_sum = self.abs_sum / scale
        return out


@dataclass
class AwqResult:
    best_scale: np.ndarray            # (K,) float32: multiply W rows by it, divide the inputs by it
    losses: np.ndarray                # (n_grid,) float64
    best_clip_ratio: float | None = None
    clip_losses: np.ndarray | None = None


def _loss(w, row_scale, stats, qt, st, gsz, sym, rr, clip, precision, out, ws):
    lib = _lib.load()
    k, n = int(w.shape[0]), int(w.shape[1])
    rc = lib.b200q_awq_loss(w.data_ptr(), k, n, dev.ptr(row_scale), stats.gram.data_ptr(),
                            float(stats.tokens), qt, st, gsz, int(bool(sym)), int(bool(rr)), float(clip),
                            _lib.PRECISION[precision], out.data_ptr(), ws.data_ptr(), ws.numel(),
                            dev.stream_ptr())
    _lib.check(rc, "b200q_awq_loss")


def weight_scale(w: torch.Tensor, strategy, group_size=-1) -> torch.Tensor:
    """``_compute_weight_scale`` (awq.py:52-70) of a (K,N) weight → (K,) float32 on the device."""
    lib = _lib.load()
    k, n = D._check_weight(w)
    st = D._strategy(strategy)
    gsz = int(group_size) if group_size else -1
    out = torch.empty((k,), dtype=torch.float32, device=w.device)
    ws = dev.workspace(lib.b200q_awq_workspace_bytes(k, n, st, gsz))
    rc = lib.b200q_awq_weight_scale(w.data_ptr(), k, n, st, gsz, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                    dev.stream_ptr())
    _lib.check(rc, "b200q_awq_weight_scale")
    return out


def awq_search(weights, stats: AwqStatistics, quant_type, strategy, group_size=-1, is_symmetric=False,
               reduce_range=False, clip_search=False, n_grid: int = 20, precision: str = "bf16x3") -> AwqResult:
    """The scale grid (and optionally the clip grid) of the reference's AWQ pass for one weight."""
    lib = _lib.load()
    w = dev.to_device_f32(weights)
    k, n = D._check_weight(w)
    qt, st = D._qt(quant_type), D._strategy(strategy)
    gsz = int(group_size) if group_size else -1
    ws = dev.workspace(lib.b200q_awq_workspace_bytes(k, n, st, gsz))
    act = stats.activation_scale
    wsc = weight_scale(w, strategy, gsz)
    losses = torch.zeros((n_grid,), dtype=torch.float64, device=w.device)
    scales = []
    for i in range(n_grid):
        ratio = i * 1 / n_grid
        s = torch.clamp(torch.pow(act, ratio) / torch.pow(wsc, 1 - ratio), min=1e-4)      # awq.py:147-149
        s = s / torch.sqrt(s.max() * s.min())                                               # awq.py:150
        scales.append(s)
        _loss(w, s, stats, qt, st, gsz, is_symmetric, reduce_range, 1.0, precision, losses[i:i + 1], ws)
    host = losses.cpu().numpy()
    best = int(np.argmin(host))          # first minimum, like the reference's strict `<`
    res = AwqResult(scales[best].cpu().numpy(), host)
    if clip_search:
        s = scales[best]
        w2 = (w * s[:, None]).contiguous()
        st2 = stats.rescaled(s)
        closs = torch.zeros((10,), dtype=torch.float64, device=w.device)
        for i in range(10):
            _loss(w2, None, st2, qt, st, gsz, is_symmetric, reduce_range, 1 - i / 100, precision,
                  closs[i:i + 1], ws)
        ch = closs.cpu().numpy()
        res.best_clip_ratio = 1 - int(np.argmin(ch)) / 100
        res.clip_losses = ch
    return res
