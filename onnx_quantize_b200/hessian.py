"""GPTQ Hessian accumulation on the tensor cores (device-resident)."""
from __future__ import annotations

import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib


def hessian_accumulate(x: torch.Tensor, h: torch.Tensor, alpha: float, beta: float,
                       precision: str = "bf16x3") -> torch.Tensor:
    """H <- beta*H + alpha * XᵀX in place; ``x`` is (..., K) float32 CUDA, ``h`` (K,K) float32."""
    lib = _lib.load()
    if not (x.is_cuda and x.dtype == torch.float32 and h.is_cuda and h.dtype == torch.float32):
        raise ValueError("x and h must be float32 CUDA tensors")
    k = int(x.shape[-1])
    x2 = x.reshape(-1, k)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if h.shape != (k, k) or not h.is_contiguous():
        raise ValueError("h must be a contiguous (K, K) tensor")
    t = int(x2.shape[0])
    ws = dev.workspace(lib.b200q_hessian_workspace_bytes(t, k, _lib.PRECISION[precision]))
    rc = lib.b200q_hessian_accumulate(x2.data_ptr(), t, k, float(alpha), float(beta), h.data_ptr(),
                                      _lib.PRECISION[precision], ws.data_ptr(), ws.numel(),
                                      dev.stream_ptr())
    _lib.check(rc, "b200q_hessian_accumulate")
    return h


class HessianAccumulator:
    """Streaming form of the reference's ``_accumulate_hessian`` (gptq.py:246-260): batches are
    folded as they arrive, activations are never kept.  ``num_samples`` counts *samples* (the
    leading dimension of each batch), exactly as the reference does."""

    def __init__(self, k: int, precision: str = "bf16x3", device=None):
        self.device = device or dev.require_cuda()
        self.h = torch.zeros((k, k), dtype=torch.float32, device=self.device)
        self.num_samples = 0
        self.precision = precision

    def add(self, inp) -> None:
        x = dev.to_device_f32(inp)
        added = int(x.shape[0])
        total = self.num_samples + added
        hessian_accumulate(x, self.h, alpha=2.0 / total, beta=self.num_samples / total,
                           precision=self.precision)
        self.num_samples = total
