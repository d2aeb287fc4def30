"""CPU oracle for the quantization hot path (TEST INFRASTRUCTURE ONLY — never shipped, never
imported by the product package ``onnx_quantize_b200``).

This is a NumPy restatement of the numeric core of AyoubMDL/onnx_quantize v0.3.0.  It is written
against the *behaviour* of the reference and is pinned two ways:

  * ``tests/test_oracle_golden.py`` checks it against ``tests/golden/*.npz`` — outputs of the
    unmodified reference executed in the build container through ``oracle/ref_shim.py`` by
    ``oracle/gen_golden.py`` (both committed) — and against the known-answer tables of the
    reference's own test-suite (scale/zp table, nibble packing bytes, MatMulNBits zero-point
    nibble order, MinMax/EMA values);
  * when ``/root/reference`` is present the same tests also compare it live against the
    reference on fresh random inputs.

Parity status: PINNED (golden vectors + live reference) for every function in this file except
``gptq(..., mode="propagate")`` which has no counterpart in the reference (the reference's GPTQ
performs no error propagation as written — SURVEY.md finding 3); that mode is pinned against
the reference source with the two-token transposition fix applied at test time
(``oracle/gen_golden.py::patched_reference_gptq``).

Why NumPy and not C: the reference *is* NumPy, and three of its numerics are properties of NumPy
itself (NEP-50 weak-scalar promotion, the pairwise / strided summation order of ``np.sum``, the
SIMD ``np.power``).  Restating on the same library keeps those identical by construction; the
explicit scalar orders are spelled out in ``oracle/c_sumorder.c`` (a C restatement of the
summation orders only) which the tests use to pin the device kernels' summation order
independently of NumPy.

Conventions: quantization types are the strings "int4", "uint4", "int8", "uint8", "int32";
strategies are "tensor", "channel", "group".  Weights are (K=in, N=out) row-major float32, the
ONNX MatMul layout.
"""
from __future__ import annotations

import math

import ml_dtypes
import numpy as np

# ----------------------------------------------------------------------------------------------
# q-ranges — reference: core/_dtypes.py:8-31 (tables) and :61-70 (selection order)
# ----------------------------------------------------------------------------------------------
_FULL = {"uint4": (0, 15), "int4": (-8, 7), "uint8": (0, 255), "int8": (-128, 127),
         "uint32": (0, 2**32 - 1), "int32": (-(2**31), 2**31 - 1)}
_SYM = {"int4": (-7, 7), "int8": (-127, 127), "int32": (-(2**31 - 1), 2**31 - 1)}
_REDUCED = {"uint4": (0, 7), "int4": (-4, 3), "uint8": (0, 127), "int8": (-64, 64),
            "uint32": (0, 2**31 - 1), "int32": (-(2**30), 2**30)}
_NP = {"uint4": np.dtype(ml_dtypes.uint4), "int4": np.dtype(ml_dtypes.int4),
       "uint8": np.dtype(np.uint8), "int8": np.dtype(np.int8),
       "uint32": np.dtype(np.uint32), "int32": np.dtype(np.int32)}


def qrange(qtype: str, symmetric: bool, reduce_range: bool = False):
    """reduce_range wins, then the symmetric table (signed types only), else the full range."""
    if reduce_range:
        return _REDUCED[qtype]
    if symmetric and qtype in _SYM:
        return _SYM[qtype]
    return _FULL[qtype]


def np_dtype(qtype: str) -> np.dtype:
    return _NP[qtype]


# ----------------------------------------------------------------------------------------------
# A1 layout — reference: core/_algorithms/utils.py:6-39
# ----------------------------------------------------------------------------------------------
def to_rows(w: np.ndarray, strategy: str, group_size: int = -1) -> np.ndarray:
    """View/copy of ``w`` whose rows share one (scale, zp).

    tensor → w itself; channel → w.T (an F-ordered *view*: this matters for the summation order
    of the MSE error); group → C-contiguous (N*K/gs, gs) copy, row index = n*(K/gs) + g.
    """
    if strategy == "tensor":
        return w
    if strategy == "channel":
        return w.T
    k = w.shape[0]
    gs = k if (group_size == -1 or group_size > k) else group_size
    return w.T.reshape((-1, gs))


def from_rows(rows: np.ndarray, like: np.ndarray, strategy: str) -> np.ndarray:
    if strategy == "tensor":
        return rows
    if strategy == "channel":
        return rows.T
    return rows.reshape(like.T.shape).T


# ----------------------------------------------------------------------------------------------
# A2 min/max — reference: utils.py:42-69
# ----------------------------------------------------------------------------------------------
def row_min_max(rows: np.ndarray, strategy: str, clip_ratio: float = 1.0):
    if strategy == "tensor":
        lo, hi = np.min(rows), np.max(rows)
    else:
        lo = np.min(rows, axis=1, keepdims=True)
        hi = np.max(rows, axis=1, keepdims=True)
    # python-float clip ratio is a weak scalar → the multiply happens in the array dtype (f32)
    lo, hi = lo * clip_ratio, hi * clip_ratio
    return np.array(np.minimum(lo, 0)), np.array(np.maximum(hi, 0))


# ----------------------------------------------------------------------------------------------
# A3 scale / zero-point — reference: utils.py:242-299
# ----------------------------------------------------------------------------------------------
def qparams(rmin, rmax, qtype: str, symmetric: bool, reduce_range: bool,
            scale_dtype=np.float32, zp_dtype=None):
    zp_dtype = np_dtype(qtype) if zp_dtype is None else zp_dtype
    rmin, rmax = np.asarray(rmin), np.asarray(rmax)
    if symmetric:
        amax = np.maximum(np.abs(rmin), np.abs(rmax))
        lo, hi = qrange(qtype, True, reduce_range)
        mid = np.round((hi + lo) / 2.0)          # np.float64 scalar: a *strong* type under NEP-50
        levels = min(hi - mid, mid - lo)          # → the division below is carried out in float64
        scale = amax / levels
        scale = np.where(scale < np.finfo(amax.dtype).tiny, 1, scale)
        zp = np.ones(amax.shape) * mid
        return scale.astype(scale_dtype), np.asarray(zp, dtype=zp_dtype)
    lo, hi = qrange(qtype, False, reduce_range)
    scale = (rmax - rmin) / (hi - lo)            # python int divisor → stays in the array dtype
    scale = np.where(scale < np.finfo(rmax.dtype).tiny, 1, scale)
    zp = np.round(np.clip(lo - (rmin / scale), lo, hi))
    return scale.astype(scale_dtype), np.asarray(zp, dtype=zp_dtype)


# ----------------------------------------------------------------------------------------------
# A4 / A5 quantize / dequantize — reference: utils.py:72-79, :102-137
# ----------------------------------------------------------------------------------------------
def quantize_rows(rows, scale, zp, qtype: str, symmetric: bool, reduce_range: bool):
    lo, hi = qrange(qtype, symmetric, reduce_range)
    shifted = np.round(rows / scale).astype(np.int32) + zp
    return np.clip(shifted, lo, hi).astype(np_dtype(qtype))


def dequantize(q, scale, zp):
    return (q.astype(np.float32) - zp.astype(np.float32)) * scale


def dequantize_weight(q, scale, zp, strategy: str, group_size: int = -1):
    """(K,N) codes → (K,N) float32, broadcasting scale/zp according to the strategy."""
    rows = to_rows(q, strategy, group_size)
    if strategy == "channel":
        scale, zp = np.expand_dims(scale, 1), np.expand_dims(zp, 1)
    return from_rows(dequantize(rows, scale, zp), q, strategy)


# ----------------------------------------------------------------------------------------------
# A6 MSE shrink-grid search — reference: utils.py:140-239
# ----------------------------------------------------------------------------------------------
def mse_min_max(rows, qtype: str, strategy: str, symmetric: bool, reduce_range: bool,
                zp_dtype=None, maxshrink: float = 0.20, patience: int = 5, grid: float = 100.0,
                norm: float = 2.4, return_trace: bool = False):
    """Best (rmin, rmax) per row over the shrink grid p = 1 - i/grid, i = 0 .. maxshrink*grid-1.

    Error = sum |fake_quant(x) - x| ** norm with ``norm`` a weak python float (→ float32(2.4)).
    Per-row strict ``<`` keeps the first minimum.  The early-stop counter is *global*: it is
    incremented only when no row at all improved at step i, it is never reset, and the loop ends
    when it reaches ``patience``.  The starting range uses clip_ratio = 1.0.
    """
    whole = strategy == "tensor"
    lo0, hi0 = row_min_max(rows, strategy, 1.0)
    best_err = np.full_like(lo0, np.finfo(lo0.dtype).max)
    best_lo, best_hi = lo0.copy(), hi0.copy()
    stalls = 0
    trace = []
    for i in range(int(maxshrink * grid)):
        p = 1 - i / grid
        lo, hi = p * lo0, p * hi0
        s, z = qparams(lo, hi, qtype, symmetric, reduce_range, np.float32, zp_dtype)
        resid = dequantize(quantize_rows(rows, s, z, qtype, symmetric, reduce_range), s, z)
        resid -= rows
        resid = np.power(np.abs(resid), norm)
        err = np.sum(resid) if whole else np.sum(resid, axis=1, keepdims=True)
        better = err < best_err
        if return_trace:
            trace.append(np.array(err, copy=True))
        if np.any(better):
            best_err[better] = err[better]
            best_lo[better] = lo[better]
            best_hi[better] = hi[better]
        else:
            stalls += 1
        if stalls >= patience:
            break
    if return_trace:
        return best_lo, best_hi, trace
    return best_lo, best_hi


# ----------------------------------------------------------------------------------------------
# A7 / A8 RTN — reference: utils.py:302-348, rtn.py:54-109
# ----------------------------------------------------------------------------------------------
def qparams_from_rows(rows, qtype, strategy, symmetric, reduce_range, clip_ratio, mse,
                      zp_dtype=None):
    lo, hi = row_min_max(rows, strategy, clip_ratio)
    if mse:  # replaces the clipped range: clip_ratio has no effect when mse=True
        lo, hi = mse_min_max(rows, qtype, strategy, symmetric, reduce_range, zp_dtype)
    return qparams(lo, hi, qtype, symmetric, reduce_range, np.float32, zp_dtype)


def rtn_quantize(w, qtype: str, strategy: str, group_size: int = -1, symmetric: bool = False,
                 reduce_range: bool = False, clip_ratio: float = 1.0, mse: bool = False,
                 zp_dtype=None):
    """→ (codes (K,N) in the 1-byte-per-element numpy dtype, scale f32, zp).

    scale/zp shapes: tensor → (), channel → (N,), group → (N*K/gs, 1) with row = n*(K/gs)+g.
    """
    rows = to_rows(w, strategy, group_size)
    s, z = qparams_from_rows(rows, qtype, strategy, symmetric, reduce_range, clip_ratio, mse,
                             zp_dtype)
    q = quantize_rows(rows, s, z, qtype, symmetric, reduce_range)
    if strategy in ("tensor", "channel"):
        s, z = np.squeeze(s), np.squeeze(z)
    return from_rows(q, w, strategy), s, z


# ----------------------------------------------------------------------------------------------
# A9 int32 bias — reference: rtn.py:112-138
# ----------------------------------------------------------------------------------------------
def quantize_bias(bias, input_scale, weight_scale):
    assert bias.ndim == 1 and bias.dtype == np.float32
    assert np.size(input_scale) == 1 and weight_scale.dtype == np.float32
    assert weight_scale.size == 1 or weight_scale.size == bias.size
    s = weight_scale * input_scale
    return quantize_rows(bias, s, 0, "int32", False, False), s, 0


# ----------------------------------------------------------------------------------------------
# P1 flat nibble packing (layout A) — reference: core/_pack.py:8-22, :25-38, :52-66
# ----------------------------------------------------------------------------------------------
def pack4_flat(codes, qtype: str) -> np.ndarray:
    """Row-major flatten; byte j = (e[2j] & 0xF) | (e[2j+1] & 0xF) << 4; odd tail padded with 0."""
    flat = np.asarray(codes).astype(np.int16).ravel()
    if qtype == "int4":
        flat = np.where(flat < 0, flat + 16, flat)
    flat = (flat & 0xF).astype(np.uint8)
    if flat.size % 2:
        flat = np.concatenate([flat, np.zeros(1, np.uint8)])
    return flat[0::2] | (flat[1::2] << 4)


def unpack4_flat(packed, shape, qtype: str) -> np.ndarray:
    assert packed.dtype == np.uint8
    n = int(np.prod(shape))
    both = np.empty(packed.size * 2, np.uint8)
    both[0::2] = packed & 0xF
    both[1::2] = packed >> 4
    both = both[:n].reshape(shape)
    if qtype == "int4":
        s = both.astype(np.int8)
        return np.where(s > 7, s - 16, s).astype(np.int8)
    return both


# ----------------------------------------------------------------------------------------------
# P2 MatMulNBits layout (layout B) — reference: qrules/_common.py:65-123
# ----------------------------------------------------------------------------------------------
def matmul_nbits_layout(q, scale, zp, group_size: int, bits: int, zp_is_float: bool = False):
    """codes (K,N) u8-valued → B (N, K/gs, gs*bits/8) u8; scale (N, K/gs) f32; zp (N, ceil(G/2)).

    4-bit: B[n,g,j] = q[g*gs+2j, n] | q[g*gs+2j+1, n] << 4.  Zero points are packed low nibble
    first along g per output channel, an odd count padded with 0x8; they are NOT packed when
    there is a single block per channel (or when zp is float — HQQ, not on this path).
    """
    k, n = q.shape
    assert k % group_size == 0
    g = k // group_size
    blob = group_size * bits // 8
    cols = np.asarray(q).astype(np.uint8).T.reshape(-1, group_size)
    if bits == 4:
        cols = cols[:, 0::2] | (cols[:, 1::2] << 4)
    b = cols.reshape(-1, g, blob)
    scale = scale.reshape(-1, g)
    out_zp = zp
    if bits == 4 and g > 1 and not zp_is_float:
        z = np.asarray(zp).astype(np.uint8).reshape(n, g)
        if g % 2:
            z = np.concatenate([z, np.full((n, 1), 0x8, np.uint8)], axis=1)
        out_zp = (z[:, 0::2] & 0xF) | ((z[:, 1::2] & 0xF) << 4)
    out_zp = np.reshape(out_zp, (n, -1))
    if not zp_is_float:
        out_zp = out_zp.astype(np.uint8)
    return b, scale, out_zp


# ----------------------------------------------------------------------------------------------
# C1 MinMax calibrator — reference: core/_calibration/minmax.py:40-87
# ----------------------------------------------------------------------------------------------
class MinMax:
    """Running (or EMA-smoothed) global min/max per tensor name."""

    def __init__(self, momentum: float = 0.0):
        assert 0 <= momentum < 1, "Momentum must be in the range [0, 1)."
        self.momentum = momentum
        self.stats: dict[str, list] = {}

    def collect(self, name: str, array) -> None:
        lo, hi = np.min(array), np.max(array)
        if name not in self.stats:
            self.stats[name] = [lo, hi]
            return
        cur = self.stats[name]
        m = self.momentum
        if m > 0:  # python-float weights are weak scalars → the EMA runs in the array dtype
            cur[0] = m * cur[0] + (1 - m) * lo
            cur[1] = m * cur[1] + (1 - m) * hi
        else:
            cur[0] = np.minimum(cur[0], lo)
            cur[1] = np.maximum(cur[1], hi)

    def compute_range(self, name: str):
        if name not in self.stats:
            raise KeyError(f"No calibration data collected for '{name}'")
        lo, hi = self.stats[name]
        return (np.array(np.minimum(lo, 0), dtype=np.float32),
                np.array(np.maximum(hi, 0), dtype=np.float32))


# ----------------------------------------------------------------------------------------------
# G1 Hessian — reference: core/_algorithms/gptq.py:246-260
# ----------------------------------------------------------------------------------------------
def accumulate_hessian(inp, h, num_samples: int):
    """H ← H*n/(n+b) + (sqrt(2/(n+b))·X)ᵀ(sqrt(2/(n+b))·X), b = inp.shape[0] (SAMPLES, not tokens)."""
    added = inp.shape[0]
    x = np.reshape(inp, (-1, inp.shape[-1]))
    h *= num_samples / (num_samples + added)
    num_samples += added
    x = math.sqrt(2 / num_samples) * x.astype(np.float32)
    h += np.matmul(x.T, x)
    return h, num_samples


# ----------------------------------------------------------------------------------------------
# G2 Hessian inverse factor — reference: gptq.py:119-150
# ----------------------------------------------------------------------------------------------
def hinv_cholesky_upper(h, percdamp: float):
    """U upper-triangular with UᵀU = (H + damp·I)⁻¹; (U, ok).  ok=False → identity fallback."""
    k = h.shape[0]
    h = h.copy()
    try:
        damp = percdamp * np.mean(np.diag(h))
        idx = np.arange(k)
        h[idx, idx] += damp
        l = np.linalg.cholesky(h)
        linv = np.linalg.inv(l)
        return np.linalg.cholesky(linv.T @ linv).T, True
    except np.linalg.LinAlgError:
        return np.eye(k, dtype=h.dtype), False


# ----------------------------------------------------------------------------------------------
# G2-G4 GPTQ — reference: gptq.py:76-243
# ----------------------------------------------------------------------------------------------
def gptq(w, h, qtype="int8", strategy="channel", group_size=32, symmetric=False,
         reduce_range=False, clip_ratio=1.0, block_size=128, percdamp=0.01, actorder=False,
         mse=False, zp_dtype=None, mode="reference", return_aux=False):
    """Blockwise GPTQ over the K rows of ``w`` (K,N).

    mode="reference": exactly what the reference computes.  Its in-block rank-1 update reads
        column i of the *upper-triangular* factor below the diagonal (``U1[i:, i]``) and its
        block update reads ``U[i2:, i1:i2]`` — both are the zero triangle, so no quantization
        error ever reaches another row; the codes equal slice-wise RTN of the (dead-masked,
        optionally act-ordered) weight.
    mode="propagate": the transposed (non-zero) triangle is used — ``U1[i, i:]`` and
        ``U[i1:i2, i2:]ᵀ`` — i.e. GPTQ as published (arXiv:2210.17323, lazy batch updates).

    In both modes the returned scale/zp are re-derived from the *dequantized* result with the
    user's strategy/group size (the clip ratio is therefore applied twice).
    """
    assert mode in ("reference", "propagate")
    zp_dtype = np_dtype(qtype) if zp_dtype is None else zp_dtype
    w = w.copy()
    h = h.copy()
    k = w.shape[0]

    # per-output-channel parameters over all K (only used when there are no groups)
    s, z = qparams_from_rows(w.T, qtype, "channel" if strategy == "group" else strategy,
                             symmetric, reduce_range, clip_ratio, mse, zp_dtype)
    s, z = np.squeeze(s), np.squeeze(z)

    dead = np.diag(h) == 0
    h[dead, dead] = 1
    w[dead, :] = 0

    perm = None
    if actorder:
        perm = np.argsort(np.diag(h))[::-1]
        w = w[perm, :]
        h = h[perm, :][:, perm]

    u, ok = hinv_cholesky_upper(h, percdamp)

    deq = np.zeros_like(w)
    codes = np.zeros_like(w)
    grouped = bool(group_size) and group_size != -1
    for b0 in range(0, k, block_size):
        b1 = min(b0 + block_size, k)
        blk = w[b0:b1, :].copy()
        ub = u[b0:b1, b0:b1]
        errs = np.zeros_like(blk)
        for r in range(b1 - b0):
            row = blk[r, :]
            d = ub[r, r]
            if grouped and (b0 + r) % group_size == 0:
                # NB: slices the *global* w (updated only by finished blocks), F-ordered view
                s, z = qparams_from_rows(w[b0 + r:b0 + r + group_size, :].T, qtype, "channel",
                                         symmetric, reduce_range, clip_ratio, mse, zp_dtype)
                s, z = np.squeeze(s), np.squeeze(z)
            qi = quantize_rows(row, s, z, qtype, symmetric, reduce_range).flatten()
            dq = dequantize(qi, s, z)
            deq[b0 + r, :] = dq
            codes[b0 + r, :] = qi
            e = (row - dq) / d
            coeff = ub[r:, r] if mode == "reference" else ub[r, r:]
            blk[r:, :] -= np.matmul(coeff[:, None], e[None, :])
            errs[r, :] = e
        tail = u[b1:, b0:b1] if mode == "reference" else u[b0:b1, b1:].T
        w[b1:, :] -= np.matmul(tail, errs)

    if actorder:
        inv = np.argsort(perm)
        deq, codes = deq[inv, :], codes[inv, :]

    codes = codes.astype(np_dtype(qtype))
    rows = to_rows(deq, strategy, group_size)
    s, z = qparams_from_rows(rows, qtype, strategy, symmetric, reduce_range, clip_ratio, mse,
                             zp_dtype)
    if strategy in ("tensor", "channel"):
        s, z = np.squeeze(s), np.squeeze(z)
    out = (codes, s.astype(np.float32), z.astype(codes.dtype))
    if return_aux:
        return out + ({"deq": deq, "u": u, "chol_ok": ok, "perm": perm, "dead": dead},)
    return out


def gptq_quantize(weights, inputs, qtype="int8", strategy="channel", group_size=32,
                  symmetric=False, reduce_range=False, clip_ratio=1.0, block_size=128,
                  percdamp=0.01, actorder=False, mse=False, zp_dtype=None, mode="reference"):
    """reference: gptq.py:263-324 — one Hessian accumulation over all inputs, then ``gptq``."""
    zp_dtype = np.dtype(np.int8) if zp_dtype is None else zp_dtype
    h = np.zeros((weights.shape[0], weights.shape[0]), dtype=np.float32)
    h, _ = accumulate_hessian(inputs, h, 0)
    return gptq(weights, h, qtype, strategy, group_size, symmetric, reduce_range, clip_ratio,
                block_size, percdamp, actorder, mse, zp_dtype, mode)


def layer_output_rel_mse(x, w, w_hat) -> float:
    """‖X(Ŵ−W)‖² / ‖XW‖² in float64 — the GPTQ quality gate of BASELINE.json."""
    x = np.reshape(x, (-1, x.shape[-1])).astype(np.float64)
    num = np.linalg.norm(x @ (w_hat.astype(np.float64) - w.astype(np.float64))) ** 2
    return float(num / np.linalg.norm(x @ w.astype(np.float64)) ** 2)


# ----------------------------------------------------------------------------------------------
# AWQ scale / clip search — reference: pre_passes/awq.py:47-70 (scales), :114-184 (scale grid),
# :207-254 (clip grid).  Pinned live against the reference's AwqPass driven through
# oracle/ref_shim.py::run_reference_awq (tests/test_oracle_golden.py).
# ----------------------------------------------------------------------------------------------
def awq_activation_scale(x):
    """mean |x| per input channel over all tokens (awq.py:47-50)."""
    return np.mean(np.reshape(np.abs(x), (-1, x.shape[-1])), axis=0)


def awq_weight_scale(w, strategy: str, group_size: int):
    """mean over output channels of |w| / max|w| of its parameter row (awq.py:52-70); w is (K,N)."""
    wt = w.T
    shape = wt.shape
    rows = np.reshape(wt, (-1, group_size)) if strategy == "group" else wt
    if strategy == "tensor":
        s = np.abs(rows) / np.max(np.abs(rows))
    else:
        s = np.abs(rows) / np.max(np.abs(rows), axis=1, keepdims=True)
    return np.mean(np.reshape(s, shape), axis=0)


def awq_fake_quant(w, qtype, strategy, group_size, symmetric, reduce_range, clip_ratio=1.0):
    q, s, z = rtn_quantize(w, qtype, strategy, group_size, symmetric, reduce_range, clip_ratio, False)
    return dequantize_weight(q, s, z, strategy, group_size)


def awq_scale_search(w, x, qtype, strategy, group_size, symmetric=False, reduce_range=False, n_grid=20):
    """→ (best per-input-channel scale (K,), losses (n_grid,)) — awq.py:121-184."""
    act = awq_activation_scale(x)
    ws = awq_weight_scale(w, strategy, group_size)
    ref_out = np.matmul(x, w)
    best, best_scale, losses = np.inf, None, []
    for i in range(n_grid):
        ratio = i * 1 / n_grid
        scale = np.clip(np.power(act, ratio) / np.power(ws, (1 - ratio)), 1e-4, None)
        scale = scale / np.sqrt(np.max(scale) * np.min(scale))
        col = scale.reshape(-1, 1)
        wq = awq_fake_quant(w * col, qtype, strategy, group_size, symmetric, reduce_range) / col
        diff = (ref_out - np.matmul(x, wq)).ravel()
        loss = float(diff @ diff) / diff.size
        losses.append(loss)
        if loss < best:
            best, best_scale = loss, scale
    return best_scale, np.array(losses)


def awq_clip_search(w, x, qtype, strategy, group_size, symmetric=False, reduce_range=False):
    """→ (best clip ratio, losses (10,)) — awq.py:207-254."""
    ref_out = np.matmul(x, w)
    best, best_ratio, losses = np.inf, 1, []
    for i in range(10):
        ratio = 1 - i / 100
        wq = awq_fake_quant(w, qtype, strategy, group_size, symmetric, reduce_range, ratio)
        diff = (ref_out - np.matmul(x, wq)).ravel()
        loss = float(diff @ diff) / diff.size
        losses.append(loss)
        if loss < best:
            best, best_ratio = loss, ratio
    return best_ratio, np.array(losses)


# ----------------------------------------------------------------------------------------------
# SmoothQuant — reference: pre_passes/smooth_quant.py:62-74, :104-116.  Pinned live through
# oracle/ref_shim.py::run_reference_smooth_quant.
# ----------------------------------------------------------------------------------------------
def smooth_quant(w, x, alpha: float = 0.5):
    """→ (scale (K,), updated weights (K,N)); the layer input is divided by ``scale``."""
    act = np.maximum(np.max(np.abs(x.reshape(-1, x.shape[-1])), axis=0), 1e-5)
    wmax = np.max(np.abs(w), axis=1)
    scale = np.power(act, alpha) / np.power(wmax + 1e-9, (1 - alpha))
    return scale, np.multiply(scale.reshape(-1, 1), w)


# ----------------------------------------------------------------------------------------------
# HQQ — reference: core/_algorithms/hqq.py (`_shrink_op` :103-104, `_optimize_zero_point`
# :107-146, `_hqq_quantize` :149-217).  Pinned live through tests/test_oracle_golden.py and by
# tests/golden/hqq.npz.  uint4, asymmetric, GROUP only.
# ----------------------------------------------------------------------------------------------
def hqq_shrink(x, beta: float, lp_norm: float):
    ax = np.abs(x)
    return np.sign(x) * np.maximum(0, ax - (1.0 / beta) * np.power(ax + 1e-8, lp_norm - 1))


def hqq_optimize_zero_point(rows, scale, zp, qmin, qmax, lp_norm=0.7, beta=1e1, kappa=1.01, iters=20,
                            early_stop=True, return_trace=False):
    """→ best zero point (rows,1) [, (best iteration or -1, list of global mean errors)]."""
    best_err, best_zp, best_it = np.inf, zp.copy(), -1
    inv = 1.0 / scale                                  # HQQ iterates with the inverted scale
    errors = []
    for it in range(iters):
        w_q = np.clip(np.round(rows * inv + zp), qmin, qmax)
        w_r = (w_q - zp) / inv
        w_e = hqq_shrink(rows - w_r, beta, lp_norm)
        beta *= kappa
        err = float(np.mean(np.abs(rows - w_r)))
        errors.append(err)
        if err < best_err:
            best_err, best_zp, best_it = err, zp.copy(), it
        elif early_stop:
            break
        zp = np.mean(w_q - (rows - w_e) * inv, axis=1, keepdims=True)
    if return_trace:
        return best_zp, (best_it, errors)
    return best_zp


def hqq_quantize(w, group_size: int, reduce_range: bool = False, clip_ratio: float = 1.0, mse: bool = False,
                 lp_norm=0.7, beta=1e1, kappa=1.01, iters=20, early_stop=True, return_trace=False):
    """→ (codes (K,N) uint4, scale (N*G,1) f32, zero point (N*G,1) f32)."""
    rows = to_rows(w, "group", group_size)
    scale, zp = qparams_from_rows(rows, "uint4", "group", False, reduce_range, clip_ratio, mse,
                                  zp_dtype=np.float32)
    qmin, qmax = qrange("uint4", False, reduce_range)
    zp, trace = hqq_optimize_zero_point(rows, scale, zp, qmin, qmax, lp_norm, beta, kappa, iters, early_stop,
                                        return_trace=True)
    q = np.clip(np.round(rows / scale + zp), qmin, qmax).astype(np_dtype("uint4"))
    out = (from_rows(q, w, "group"), scale, zp)
    return out + (trace,) if return_trace else out
