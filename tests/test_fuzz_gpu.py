"""Randomised differential test: product (CUDA, through the reference-shaped API) against the
pinned oracle over random configurations and awkward shapes — odd N, K not a multiple of 8, group
size equal to K or not dividing the tile, single rows/columns, constant and sign-definite columns,
huge outliers, zeros.  Everything must be bit-identical (RTN, with and without the MSE search)."""
import numpy as np
import pytest

import onnx_quantize_b200 as q
from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
from oracle import np_oracle as O
from tests.helpers import as_i8, bits, stable_seed

pytestmark = pytest.mark.gpu
QT = {"int4": q.QuantType.QInt4, "uint4": q.QuantType.QUInt4, "int8": q.QuantType.QInt8,
      "uint8": q.QuantType.QUInt8}


def _weights(rng, k, n, kind):
    w = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
    if kind == 1:
        w[rng.integers(0, k, 4), rng.integers(0, n, 4)] *= 1000.0          # outliers
    elif kind == 2:
        w[:, : max(1, n // 3)] = np.abs(w[:, : max(1, n // 3)])              # sign-definite columns
        w[:, -1] = 0.0                                                       # an all-zero column
    elif kind == 3:
        w[:] = np.float32(0.37)                                              # constant
    elif kind == 4:
        w *= np.float32(1e-30)                                               # tiny magnitudes
    elif kind == 5:
        w = np.round(w * 64).astype(np.float32) / 64                         # many exact ties
    return w


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        qt = rng.choice(list(QT))
        strategy = rng.choice(["tensor", "channel", "group"])
        k = int(rng.choice([1, 7, 16, 24, 96, 128, 200, 256, 384, 1000]))
        n = int(rng.choice([1, 3, 16, 17, 40, 64, 100, 128, 272, 1030]))
        gs = -1
        if strategy == "group":
            divs = [d for d in (1, 2, 8, 16, 24, 32, 64, 100, 128, k) if k % d == 0]
            gs = int(rng.choice(divs + [-1, 4 * k]))
        out.append((qt, strategy, k, n, gs, bool(rng.integers(2)), bool(rng.integers(4) == 0),
                    float(rng.choice([1.0, 0.9, 0.5])), bool(rng.integers(3) == 0), int(rng.integers(6))))
    return out


@pytest.mark.parametrize("case", _cases(2024, 160), ids=lambda c: "-".join(str(v) for v in c))
def test_rtn_random_configurations_bit_exact(cuda, case):
    qt, strategy, k, n, gs, sym, rr, clip, mse, kind = case
    if mse and k * n > 40000:
        mse = False                                   # keep the NumPy oracle's 20 passes quick
    rng = np.random.default_rng(stable_seed(case))
    w = _weights(rng, k, n, kind)
    got = _rtn_quantize(w, QT[qt], q.QuantizationStrategy(strategy), gs, sym, rr, clip, mse,
                        np.dtype(np.float32), QT[qt].np_dtype)
    want = O.rtn_quantize(w, qt, strategy, gs, sym, rr, clip, mse)
    assert got[1].shape == want[1].shape and got[2].shape == want[2].shape
    assert np.array_equal(bits(got[1]), bits(want[1])), "scale"
    assert np.array_equal(as_i8(got[2], qt), as_i8(want[2], qt)), "zero point"
    assert np.array_equal(as_i8(got[0], qt), as_i8(want[0], qt)), "codes"


def _gptq_cases(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        qt = rng.choice(list(QT))
        strategy = rng.choice(["tensor", "channel", "group"])
        k = int(rng.choice([32, 64, 96, 160, 256]))
        n = int(rng.choice([1, 8, 24, 40, 130]))
        gs = int(rng.choice([d for d in (8, 16, 32, 64, k) if k % d == 0] + [-1]))
        if strategy == "group" and gs == -1:
            gs = k
        out.append((qt, strategy, k, n, gs, bool(rng.integers(2)), bool(rng.integers(4) == 0),
                    float(rng.choice([1.0, 0.9])), bool(rng.integers(4) == 0), bool(rng.integers(2)),
                    int(rng.choice([16, 32, 48, 128, 256])), int(rng.integers(2))))
    return out


@pytest.mark.parametrize("case", _gptq_cases(7, 80), ids=lambda c: "-".join(str(v) for v in c))
def test_gptq_reference_mode_random_configurations_bit_exact(cuda, case):
    """GPTQ as the reference computes it (mode="reference"): codes, scale bits and zero points
    identical to the oracle for random types / strategies / group and block sizes / act-order /
    MSE, with dead input channels."""
    from onnx_quantize_b200.core._algorithms.gptq import _gptq
    qt, strategy, k, n, gs, sym, rr, clip, mse, actorder, bs, dead = case
    rng = np.random.default_rng(stable_seed(case))
    w = (rng.standard_normal((k, n)) * 0.05).astype(np.float32)
    x = rng.standard_normal((8, 16, k)).astype(np.float32) * rng.uniform(0.5, 2.0, k).astype(np.float32)
    if dead:
        x[..., rng.integers(0, k, 2)] = 0.0
    h, _ = O.accumulate_hessian(x, np.zeros((k, k), np.float32), 0)
    got = _gptq(w, h, QT[qt], q.QuantizationStrategy(strategy), gs, sym, rr, clip, bs, 0.01, actorder, mse,
                np.dtype(np.float32), QT[qt].np_dtype, mode="reference")
    want = O.gptq(w, h, qt, strategy, gs, sym, rr, clip, bs, 0.01, actorder, mse, O.np_dtype(qt), "reference")
    assert got[1].shape == want[1].shape and got[2].shape == want[2].shape
    assert np.array_equal(as_i8(got[0], qt), as_i8(want[0], qt)), "codes"
    assert np.array_equal(bits(got[1]), bits(want[1])), "scale"
    assert np.array_equal(as_i8(got[2], qt), as_i8(want[2], qt)), "zero point"
