"""SmoothQuant per-channel scale migration on the GPU — the numerics of the reference's
``pre_passes/smooth_quant.py`` (``_compute_activation_scale`` :62-69, ``_compute_weight_scale``
:71-74, the scale formula and its fusion into the weights :110-116).  The graph side (the ``Mul``
node, the initializer swap) stays the reference's; this module returns the arrays it needs.
"""
from __future__ import annotations

import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200 import device_api as D


class SmoothQuantStatistics:
    """Streaming ``max |x|`` per input channel over all calibration batches."""

    def __init__(self, k: int, device=None):
        self.device = device or dev.require_cuda()
        self.abs_max = torch.zeros((k,), dtype=torch.float32, device=self.device)

    def add(self, inp) -> None:
        x = dev.to_device_f32(inp)
        k = int(x.shape[-1])
        x2 = x.reshape(-1, k)
        _lib.check(_lib.load().b200q_col_abs_max(x2.data_ptr(), int(x2.shape[0]), k, self.abs_max.data_ptr(),
                                                 dev.stream_ptr()), "b200q_col_abs_max")

    @property
    def activation_scale(self) -> torch.Tensor:
        return torch.clamp(self.abs_max, min=1e-5)          # smooth_quant.py:66-67


def smooth_quant(weights, stats: SmoothQuantStatistics, alpha: float = 0.5):
    """→ ``(scale (K,), updated_weights (K,N))`` as NumPy arrays: multiply the weight rows by
    ``scale`` and the layer input by ``1/scale`` (smooth_quant.py:104-116)."""
    lib = _lib.load()
    w = dev.to_device_f32(weights)
    k, n = D._check_weight(w)
    wmax = torch.empty((k,), dtype=torch.float32, device=w.device)
    _lib.check(lib.b200q_row_abs_max(w.data_ptr(), k, n, wmax.data_ptr(), dev.stream_ptr()), "b200q_row_abs_max")
    scale = torch.pow(stats.activation_scale, alpha) / torch.pow(wmax + 1e-9, 1 - alpha)
    out = torch.empty_like(w)
    _lib.check(lib.b200q_scale_rows(w.data_ptr(), k, n, scale.data_ptr(), out.data_ptr(), dev.stream_ptr()),
               "b200q_scale_rows")
    return scale.cpu().numpy(), out.cpu().numpy()
