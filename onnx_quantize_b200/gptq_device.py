"""GPTQ on the device: inverse-Hessian factor and block loop (torch CUDA tensors in and out).

Thin wrappers over ``b200q_hinv_cholesky_upper`` / ``b200q_gptq_quantize`` / ``b200q_gemm_tn``
(include/b200q.h); nothing here synchronises except ``HinvFactor.ok`` which reads the status flag.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200 import device_api as D


@dataclass
class HinvFactor:
    """Result of :func:`hinv_cholesky_upper` — everything the block loop needs from H."""

    u: torch.Tensor        # (K,K) f32 upper, UᵀU = (H[perm][:,perm] + damp·I)⁻¹ (identity on failure)
    perm: torch.Tensor     # (K,) int32 row order of the loop
    dead: torch.Tensor     # (K,) uint8, 1 where diag(H) == 0
    status: torch.Tensor   # (1,) int32 on device, bit flags NOT_POSITIVE_DEFINITE | MARGINAL_PIVOT

    @property
    def ok(self) -> bool:
        """True when the factorization succeeded (synchronises)."""
        return (int(self.status.item()) & _lib.NOT_POSITIVE_DEFINITE) == 0

    @property
    def marginal(self) -> bool:
        """True when a tensor-core precision met a pivot too small for it to decide positive
        definiteness the way float32 LAPACK would (synchronises) — see ``resolve_marginal``."""
        return (int(self.status.item()) & _lib.MARGINAL_PIVOT) != 0


def hinv_cholesky_upper(h: torch.Tensor, percdamp: float = 0.01, actorder: bool = False,
                        precision: str = "tf32x3") -> HinvFactor:
    """gptq.py:119-150 on the device: dead mask, act-order, damping, upper inverse factor."""
    lib = _lib.load()
    if not (h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and h.shape[0] == h.shape[1]
            and h.is_contiguous()):
        raise ValueError("H must be a contiguous square float32 CUDA tensor")
    k = int(h.shape[0])
    u = torch.empty((k, k), dtype=torch.float32, device=h.device)
    perm = torch.empty((k,), dtype=torch.int32, device=h.device)
    dead = torch.empty((k,), dtype=torch.uint8, device=h.device)
    status = torch.zeros((1,), dtype=torch.int32, device=h.device)
    ws = dev.workspace(lib.b200q_hinv_workspace_bytes(k))
    rc = lib.b200q_hinv_cholesky_upper(h.data_ptr(), k, float(percdamp), int(bool(actorder)),
                                       u.data_ptr(), perm.data_ptr(), dead.data_ptr(),
                                       status.data_ptr(), _lib.PRECISION[precision], ws.data_ptr(),
                                       ws.numel(), dev.stream_ptr())
    _lib.check(rc, "b200q_hinv_cholesky_upper")
    return HinvFactor(u, perm, dead, status)


def resolve_marginal(f: HinvFactor, h: torch.Tensor, percdamp: float = 0.01, actorder: bool = False) -> HinvFactor:
    """``f`` itself unless its status carries MARGINAL_PIVOT, in which case the factor is redone in
    fp32 arithmetic (CUDA cores), whose accept / LinAlgError decisions follow LAPACK's float32
    Cholesky (gptq.py:139-150).  Synchronises; call it where the host looks at the result anyway."""
    return hinv_cholesky_upper(h, percdamp, actorder, "fp32") if f.marginal else f


def gptq_quantize(w: torch.Tensor, f: HinvFactor, quant_type, strategy, group_size=-1,
                  is_symmetric=False, reduce_range=False, clip_ratio=1.0, mse=False,
                  block_size=128, mode="reference", precision="tf32x3", return_deq=False):
    """The GPTQ block loop + epilogue (gptq.py:153-243) → ``(codes (K,N) u8, scale, zp[, deq])``.

    ``scale`` / ``zp`` have ``rows`` entries in the order ``rtn_quantize`` uses (1 | N | N*K/gs).
    """
    lib = _lib.load()
    k, n = D._check_weight(w)
    qt, st = D._qt(quant_type), D._strategy(strategy)
    gsz = int(group_size) if group_size else -1
    rows = D.num_rows(k, n, st, gsz if st == _lib.STRATEGY["group"] else -1)
    codes = torch.empty((k, n), dtype=torch.uint8, device=w.device)
    scale = torch.empty((rows,), dtype=torch.float32, device=w.device)
    zp = torch.empty((rows,), dtype=torch.uint8, device=w.device)
    deq = torch.empty((k, n), dtype=torch.float32, device=w.device) if return_deq else None
    mse_i = int(bool(mse))
    nbytes = lib.b200q_gptq_workspace_bytes(k, n, st, gsz, mse_i, int(block_size))
    if nbytes == 0:
        raise ValueError("invalid GPTQ shape / block size")
    ws = dev.workspace(nbytes)
    rc = lib.b200q_gptq_quantize(w.data_ptr(), k, n, f.u.data_ptr(), f.perm.data_ptr(),
                                 f.dead.data_ptr(), qt, st, gsz, int(bool(is_symmetric)),
                                 int(bool(reduce_range)), float(clip_ratio), mse_i, int(block_size),
                                 _lib.GPTQ_MODE[mode], _lib.PRECISION[precision], codes.data_ptr(),
                                 scale.data_ptr(), zp.data_ptr(), dev.ptr(deq), ws.data_ptr(),
                                 ws.numel(), dev.stream_ptr())
    _lib.check(rc, "b200q_gptq_quantize")
    return (codes, scale, zp, deq) if return_deq else (codes, scale, zp)


def gemm_tn(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor | None = None, alpha: float = 1.0,
            accumulate: bool = False, precision: str = "tf32x3") -> torch.Tensor:
    """D ← [D +] alpha·AᵀB for row-major 2-D float32 CUDA tensors (views with a row stride work)."""
    lib = _lib.load()
    t, m = a.shape
    t2, n = b.shape
    if t != t2 or a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError("A (T,M) and B (T,N) must share T and have unit column stride")
    if d is None:
        d = torch.zeros((m, n), dtype=torch.float32, device=a.device)
    rc = lib.b200q_gemm_tn(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), d.data_ptr(),
                           d.stride(0), t, m, n, float(alpha), int(bool(accumulate)),
                           _lib.PRECISION[precision], dev.stream_ptr())
    _lib.check(rc, "b200q_gemm_tn")
    return d
