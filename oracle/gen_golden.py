"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where /root/reference is mounted:

    python oracle/gen_golden.py

Every array written here is an output of the reference's own functions (imported through
``oracle/ref_shim.py``) on the seeded inputs stored next to it.  The only non-reference output is
``gptq_propagate``: the reference's ``_gptq`` source with the two-token transposition fix of
SURVEY.md §8c applied *at run time* (``inspect.getsource`` → ``str.replace`` → ``exec``); no
reference code is stored in this repository.
"""
from __future__ import annotations

import inspect
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
QT = {"int4": "QInt4", "uint4": "QUInt4", "int8": "QInt8", "uint8": "QUInt8"}


def make_weights():
    """Small (K,N) float32 weights covering the edge cases of SURVEY.md §8c."""
    rng = np.random.default_rng(20261018)
    w = {}
    w["randn"] = (rng.standard_normal((256, 24)) * 0.02).astype(np.float32)
    out = (rng.standard_normal((256, 24)) * 0.02).astype(np.float32)
    out[rng.integers(0, 256, 40), rng.integers(0, 24, 40)] *= 25.0       # outliers
    w["outlier"] = out
    edge = (rng.standard_normal((256, 24)) * 0.5).astype(np.float32)
    edge[:, 0] = 0.0                      # all-zero output channel
    edge[0:128, 1] = 0.0                  # all-zero group
    edge[:, 2] = np.abs(edge[:, 2])       # all-positive channel
    edge[:, 3] = -np.abs(edge[:, 3])      # all-negative channel
    edge[:, 4] = 1e-41                    # denormal → scale below tiny → 1
    edge[:, 5] = 3.0                      # constant channel
    edge[7, :] = 0.0                      # zero row (dead input channel candidate)
    w["edge"] = edge
    w["ragged"] = rng.standard_normal((96, 33)).astype(np.float32)   # odd N, K not /128
    w["tiny"] = rng.standard_normal((8, 5)).astype(np.float32)
    return w


def rtn_cases():
    cases = []
    for wname in ("randn", "outlier", "edge"):
        for qt, (strategy, gs), sym, rr in itertools.product(
                QT, (("tensor", -1), ("channel", -1), ("group", 128), ("group", 32)),
                (False, True), (False, True)):
            for clip, mse in ((1.0, False), (0.9, False), (0.9, True)):
                if wname != "randn" and (rr or clip == 0.9 and not mse):
                    continue  # keep the fixture small: full grid only on one weight
                cases.append((wname, qt, strategy, gs, sym, rr, clip, mse))
    for qt, (strategy, gs) in itertools.product(
            QT, (("tensor", -1), ("channel", -1), ("group", 16), ("group", 7), ("group", 400))):
        for mse in (False, True):
            cases.append(("ragged", qt, strategy, gs, False, False, 1.0, mse))
    for strategy in ("tensor", "channel"):
        cases.append(("tiny", "int8", strategy, -1, True, False, 1.0, True))
    return cases


def case_key(c):
    wname, qt, strategy, gs, sym, rr, clip, mse = c
    return f"{wname}|{qt}|{strategy}|{gs}|{int(sym)}|{int(rr)}|{clip}|{int(mse)}"


def gen_rtn(r, weights):
    blob = {f"w::{k}": v for k, v in weights.items()}
    keys = []
    for c in rtn_cases():
        wname, qt, strategy, gs, sym, rr, clip, mse = c
        w = weights[wname]
        qtype = getattr(r.QuantType, QT[qt])
        if strategy == "group" and (w.shape[0] % min(gs, w.shape[0]) != 0):
            continue  # the reference reshape would fail; callers resolve gs first
        q, s, z = r.rtn._rtn_quantize(w, qtype, r.QuantizationStrategy(strategy), gs, sym, rr,
                                      clip, mse, np.dtype(np.float32), qtype.np_dtype)
        key = case_key(c)
        keys.append(key)
        blob[f"q::{key}"] = q.astype(np.int8 if qt.startswith("int") else np.uint8)
        blob[f"s::{key}"] = s
        blob[f"z::{key}"] = z.astype(np.int8 if qt.startswith("int") else np.uint8)
        if qt in ("int4", "uint4"):
            blob[f"packA::{key}"] = r.pack.pack(q, qtype)
        if qt in ("uint4", "uint8") and strategy == "group" and gs in (128, 32, 16):
            qc = r.QConfig(weights=r.QWeightArgs(dtype=qt, strategy="group", group_size=gs))
            b, bs, bz = r.common._prepare_for_matmul_nbits(q, s, z, qc)
            blob[f"B::{key}"], blob[f"Bs::{key}"], blob[f"Bz::{key}"] = b, bs, bz
    blob["keys"] = np.array(json.dumps(keys))
    np.savez_compressed(os.path.join(OUT, "rtn.npz"), **blob)
    print("rtn cases:", len(keys))


def gen_mse_trace(r, weights):
    """Per-candidate error sums of the reference MSE search (host np.power!) for 3 layouts."""
    blob = {}
    u = r.utils
    orig_sum = np.sum
    for strategy, gs in (("group", 128), ("channel", -1), ("tensor", -1)):
        w = weights["randn"]
        rows = u._preprocess_array(w, r.QuantizationStrategy(strategy), gs)
        errs = []

        def spy(a, *args, **kw):
            out = orig_sum(a, *args, **kw)
            errs.append(np.array(out, copy=True))
            return out

        np.sum = spy
        try:
            lo, hi = u._compute_min_max_mse(rows, r.QuantType.QUInt4,
                                            r.QuantizationStrategy(strategy), gs, False, False,
                                            np.dtype(np.float32), r.QuantType.QUInt4.np_dtype)
        finally:
            np.sum = orig_sum
        blob[f"err::{strategy}"] = np.stack([e.reshape(-1) for e in errs])
        blob[f"lo::{strategy}"], blob[f"hi::{strategy}"] = lo, hi
    blob["w"] = weights["randn"]
    np.savez_compressed(os.path.join(OUT, "mse_trace.npz"), **blob)


def gen_minmax(r):
    rng = np.random.default_rng(7)
    batches = (rng.standard_normal((6, 4, 32, 64)) * 3).astype(np.float32)
    blob = {"batches": batches}
    for m in (0.0, 0.5, 0.9):
        cal = r.calib_factory.get_calibrator(r.calib_base.CalibrationMethod.MINMAX, momentum=m)
        for b in batches:
            cal.collect("x", b)
        lo, hi = cal.compute_range("x")
        blob[f"lo::{m}"], blob[f"hi::{m}"] = lo, hi
        for qt, sym in (("int8", True), ("int8", False), ("uint8", False), ("uint8", True)):
            qtype = getattr(r.QuantType, QT[qt])
            s, z = r.utils._compute_qparams(lo, hi, qtype, sym, False, np.dtype(np.float32),
                                            qtype.np_dtype)
            blob[f"s::{m}|{qt}|{int(sym)}"] = s
            blob[f"z::{m}|{qt}|{int(sym)}"] = z
    np.savez_compressed(os.path.join(OUT, "minmax.npz"), **blob)


def gen_bias(r):
    rng = np.random.default_rng(11)
    bias = (rng.standard_normal(64) * 2).astype(np.float32)
    ws = rng.random(64).astype(np.float32) * 0.01 + 1e-4
    blob = {"bias": bias, "ws": ws}
    q, s, _ = r.rtn._quantize_bias(bias, 0.037, ws)
    blob["q_vec"], blob["s_vec"] = q, s
    q, s, _ = r.rtn._quantize_bias(bias, 0.037, ws[:1])
    blob["q_one"], blob["s_one"] = q, s
    np.savez_compressed(os.path.join(OUT, "bias.npz"), **blob)


def patched_reference_gptq(r):
    """The reference ``_gptq`` with the propagation fix of SURVEY.md §8c (built at run time)."""
    src = inspect.getsource(r.gptq._gptq)
    a, b = "Hinv1[i:, i]", "Hinv[i2:, i1:i2]"
    assert src.count(a) == 1 and src.count(b) == 1
    src = src.replace(a, "Hinv1[i, i:]").replace(b, "Hinv[i1:i2, i2:].T")
    ns = dict(vars(r.gptq))
    exec(compile(src, "<patched _gptq>", "exec"), ns)
    return ns["_gptq"]


def gen_gptq(r):
    rng = np.random.default_rng(5)
    k, n = 256, 40
    w = (rng.standard_normal((k, n)) * 0.05).astype(np.float32)
    mix = np.eye(k, dtype=np.float32) + 0.3 * rng.standard_normal((k, k)).astype(np.float32) / 16
    x = (rng.standard_normal((12, 24, k)).astype(np.float32)) @ mix   # correlated activations
    x[..., 9] = 0.0                                                    # a dead input channel
    blob = {"w": w, "x": x}
    h = np.zeros((k, k), np.float32)
    h, ns = r.gptq._accumulate_hessian(x, h, 0)
    blob["H"] = h
    patched = patched_reference_gptq(r)
    keys = []
    grid = [("int4", "group", 128, True, False, 128), ("uint4", "group", 64, False, False, 128),
            ("int8", "channel", -1, False, False, 64), ("uint8", "tensor", 32, False, True, 128),
            ("int4", "group", 32, True, True, 96), ("uint4", "group", 128, False, False, 128)]
    for qt, strategy, gs, sym, actorder, bs in grid:
        qtype = getattr(r.QuantType, QT[qt])
        kw = dict(quant_type=qtype, strategy=r.QuantizationStrategy(strategy), group_size=gs,
                  is_symmetric=sym, reduce_range=False, clip_ratio=1.0, block_size=bs,
                  percdamp=0.01, actorder=actorder, mse=False, scale_dtype=np.dtype(np.float32),
                  zp_dtype=qtype.np_dtype)
        key = f"{qt}|{strategy}|{gs}|{int(sym)}|{int(actorder)}|{bs}"
        keys.append(key)
        cast = np.int8 if qt.startswith("int") else np.uint8
        q, s, z = r.gptq._gptq(w, h, **kw)
        blob[f"ref_q::{key}"], blob[f"ref_s::{key}"], blob[f"ref_z::{key}"] = \
            q.astype(cast), s, z.astype(cast)
        q, s, z = patched(w, h, **kw)
        blob[f"prop_q::{key}"], blob[f"prop_s::{key}"], blob[f"prop_z::{key}"] = \
            q.astype(cast), s, z.astype(cast)
    blob["keys"] = np.array(json.dumps(keys))
    np.savez_compressed(os.path.join(OUT, "gptq.npz"), **blob)
    print("gptq cases:", len(keys))


HQQ_CASES = [  # (K, N, group_size, reduce_range, clip_ratio, mse, early_stop, iters, lp_norm, beta, kappa)
    (256, 48, 64, False, 1.0, False, True, 20, 0.7, 10.0, 1.01),
    (128, 40, 16, True, 0.9, False, False, 20, 0.7, 10.0, 1.01),
    (512, 24, 128, False, 1.0, True, True, 20, 0.7, 10.0, 1.01),
    (96, 8, -1, False, 1.0, False, False, 7, 0.7, 10.0, 1.01),
    (512, 16, 256, False, 1.0, False, True, 20, 0.5, 5.0, 1.05),
    (1024, 8, -1, False, 0.8, False, False, 12, 0.7, 10.0, 1.01),
]


def gen_hqq(r):
    """`_hqq_quantize` (hqq.py:149-217) on small weights; the np.power inside is host-dependent."""
    rng = np.random.default_rng(77)
    blob = {"cases": np.array(json.dumps(HQQ_CASES))}
    for i, (k, n, gs, rr, clip, mse, es, iters, lp, beta, kappa) in enumerate(HQQ_CASES):
        w = (rng.standard_normal((k, n)) * rng.uniform(0.01, 0.1)).astype(np.float32)
        q, s, z = r.hqq._hqq_quantize(w, r.QuantType.QUInt4, gs, rr, clip, mse, np.float32, np.float32,
                                      lp, beta, kappa, iters, es)
        blob[f"w::{i}"], blob[f"q::{i}"], blob[f"s::{i}"], blob[f"z::{i}"] = w, q.astype(np.uint8), s, z
    np.savez_compressed(os.path.join(OUT, "hqq.npz"), **blob)
    print("hqq cases:", len(HQQ_CASES))


def main():
    os.makedirs(OUT, exist_ok=True)
    r = ref_shim.load()
    weights = make_weights()
    gen_rtn(r, weights)
    gen_mse_trace(r, weights)
    gen_minmax(r)
    gen_bias(r)
    gen_gptq(r)
    gen_hqq(r)
    meta = {"numpy": np.__version__, "reference": "AyoubMDL/onnx_quantize v0.3.0",
            "power_dispatch": "host-dependent (SVML on AVX512_SKX hosts)"}
    with open(os.path.join(OUT, "META.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
