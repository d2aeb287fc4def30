"""Token-major dense layers on the tensor cores through the BF16x3 split (csrc/dense_bf16.cu):
``Y = act(alpha * X @ W + bias)`` with X (T,K), W (K,N), Y (T,N) float32 CUDA tensors.  A layer's
output is directly the next layer's input; the weight planes are prepared once per layer."""
from __future__ import annotations

import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib


def supported(k: int, n: int) -> bool:
    """Shapes the tensor-core route takes (the rest goes through ``gptq_device.gemm_tn``)."""
    return n % 32 == 0 and k % 4 == 0


class Planes:
    """The two bf16 planes of a float32 matrix, contraction dimension contiguous."""

    def __init__(self, rows: int, k: int, device):
        self.rows, self.k = int(rows), int(k)
        self.buf = torch.empty((_lib.load().b200q_dense_planes_bytes(self.rows, self.k),), dtype=torch.uint8,
                               device=device)

    @classmethod
    def of_rows(cls, x: torch.Tensor, into: "Planes | None" = None) -> "Planes":
        """x (rows, K) row-major: an activation batch, or a symmetric matrix as the row operand."""
        rows, k = int(x.shape[0]), int(x.shape[1])
        p = into if into is not None and (into.rows, into.k) == (rows, k) else cls(rows, k, x.device)
        _lib.check(_lib.load().b200q_dense_split_rows(x.data_ptr(), rows, k, p.buf.data_ptr(), p.buf.numel(),
                                                      dev.stream_ptr()), "b200q_dense_split_rows")
        return p

    @classmethod
    def of_weight(cls, w: torch.Tensor, into: "Planes | None" = None) -> "Planes":
        """w (K, N) row-major, stored transposed as (N, K)."""
        k, n = int(w.shape[0]), int(w.shape[1])
        p = into if into is not None and (into.rows, into.k) == (n, k) else cls(n, k, w.device)
        _lib.check(_lib.load().b200q_dense_split_transposed(w.data_ptr(), k, n, p.buf.data_ptr(), p.buf.numel(),
                                                            dev.stream_ptr()), "b200q_dense_split_transposed")
        return p


def forward_planes(a: Planes, b: Planes, alpha: float = 1.0, bias: torch.Tensor | None = None, relu: bool = False,
                   out: torch.Tensor | None = None) -> torch.Tensor:
    """(a.rows, b.rows) = act(alpha * A @ B^T + bias) from prepared planes (a.k == b.k)."""
    if a.k != b.k:
        raise ValueError("contraction lengths differ")
    y = out if out is not None else torch.empty((a.rows, b.rows), dtype=torch.float32, device=a.buf.device)
    _lib.check(_lib.load().b200q_dense_forward_planes(a.buf.data_ptr(), a.rows, b.buf.data_ptr(), b.rows, a.k,
                                                      float(alpha), dev.ptr(bias), int(bool(relu)), y.data_ptr(),
                                                      int(y.stride(0)), dev.stream_ptr()),
               "b200q_dense_forward_planes")
    return y


def dense_forward(x: torch.Tensor, w_planes: Planes, bias: torch.Tensor | None = None, relu: bool = False,
                  x_planes: Planes | None = None) -> torch.Tensor:
    """x (T, K) float32 CUDA → act(x @ W + bias) (T, N)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()):
        raise ValueError("x must be a contiguous 2-D float32 CUDA tensor")
    return forward_planes(Planes.of_rows(x, x_planes), w_planes, 1.0, bias, relu)
