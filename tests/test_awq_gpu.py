"""AWQ scale / clip search on the device (onnx_quantize_b200/awq.py) against the oracle
(np_oracle.awq_*, pinned live to the reference's AwqPass in tests/test_oracle_golden.py).

The device path scores candidates through the Gram matrix XᵀX instead of through X·W, so losses
agree to floating-point contraction error (~1e-4 relative with the 3xTF32 Gram) — plus a discrete
effect: the candidate scale vectors come from pow(), whose last bit differs between CUDA and NumPy,
and a weight that lands on the other side of a rounding point flips one code, which moves that
candidate's loss by up to ~1e-3 relative.  The chosen grid point must be the oracle's unless two
candidates are closer than that."""
import numpy as np
import pytest
import torch

import onnx_quantize_b200 as q
from onnx_quantize_b200 import awq as A
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
QT = {"int4": q.QuantType.QInt4, "uint4": q.QuantType.QUInt4, "int8": q.QuantType.QInt8,
      "uint8": q.QuantType.QUInt8}


def _data(seed, k, n):
    rng = np.random.default_rng(seed)
    w = (rng.standard_normal((k, n)) * 0.05).astype(np.float32)
    x = (rng.standard_normal((8, 40, k)) * rng.uniform(0.2, 3.0, k)).astype(np.float32)
    return w, x


def test_statistics_and_weight_scale(cuda):
    w, x = _data(0, 256, 96)
    st = A.AwqStatistics(256, precision="fp32")
    for b in np.array_split(x, 3):              # streamed in batches
        st.add(b)
    assert st.tokens == 8 * 40
    np.testing.assert_allclose(st.activation_scale.cpu().numpy(), O.awq_activation_scale(x), rtol=2e-6)
    x2 = x.reshape(-1, 256).astype(np.float64)
    np.testing.assert_allclose(st.gram.cpu().numpy(), x2.T @ x2, rtol=0, atol=2e-6 * np.abs(x2.T @ x2).max())
    for strategy, gs in (("group", 64), ("channel", -1), ("tensor", -1)):
        got = A.weight_scale(torch.from_numpy(w).to(cuda), strategy, gs).cpu().numpy()
        np.testing.assert_allclose(got, O.awq_weight_scale(w, strategy, gs), rtol=2e-6)


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("qt,strategy,gs,sym,k,n", [("uint4", "group", 64, False, 256, 96),
                                                      ("int4", "group", 128, True, 512, 128),
                                                      ("int8", "channel", -1, True, 128, 48),
                                                      ("uint8", "tensor", -1, False, 128, 40)])
def test_scale_and_clip_search_match_oracle(cuda, precision, qt, strategy, gs, sym, k, n):
    w, x = _data(k + n, k, n)
    st = A.AwqStatistics(k, precision=precision)
    st.add(x)
    res = A.awq_search(w, st, QT[qt], strategy, gs, sym, False, clip_search=True, precision=precision)
    s_ref, l_ref = O.awq_scale_search(w, x, qt, strategy, gs, sym)
    np.testing.assert_allclose(res.losses, l_ref, rtol=2e-3)
    i_ref, i_got = int(np.argmin(l_ref)), int(np.argmin(res.losses))
    assert i_got == i_ref or abs(l_ref[i_got] - l_ref[i_ref]) <= 2e-3 * l_ref[i_ref]
    if i_got == i_ref:
        np.testing.assert_allclose(res.best_scale, s_ref, rtol=1e-5)
        s = res.best_scale.reshape(-1, 1)
        c_ref, cl_ref = O.awq_clip_search((w * s).astype(np.float32), (x / s.reshape(1, 1, -1)).astype(np.float32),
                                          qt, strategy, gs, sym)
        np.testing.assert_allclose(res.clip_losses, cl_ref, rtol=2e-3)
        j_ref, j_got = int(np.argmin(cl_ref)), int(np.argmin(res.clip_losses))
        assert j_got == j_ref or abs(cl_ref[j_got] - cl_ref[j_ref]) <= 2e-3 * cl_ref[j_ref]
        if j_got == j_ref:
            assert res.best_clip_ratio == c_ref


@pytest.mark.parametrize("alpha", [0.5, 0.8])
def test_smooth_quant_matches_oracle(cuda, alpha):
    """SmoothQuant scale migration (onnx_quantize_b200/smooth_quant.py): max|x| streamed in batches,
    max|w| per input channel, scale and fused weights to pow()'s last bit of the oracle."""
    from onnx_quantize_b200.smooth_quant import SmoothQuantStatistics, smooth_quant
    rng = np.random.default_rng(4)
    k, n = 320, 136
    w = (rng.standard_normal((k, n)) * 0.05).astype(np.float32)
    x = (rng.standard_normal((9, 33, k)) * rng.uniform(0.1, 4, k)).astype(np.float32)
    x[..., 11] = 0.0
    st = SmoothQuantStatistics(k)
    for b in np.array_split(x, 4):
        st.add(b)
    assert np.array_equal(st.abs_max.cpu().numpy(), np.max(np.abs(x.reshape(-1, k)), axis=0))
    s, w2 = smooth_quant(w, st, alpha)
    s_ref, w_ref = O.smooth_quant(w, x, alpha)
    np.testing.assert_allclose(s, s_ref, rtol=1e-6)
    np.testing.assert_allclose(w2, w_ref, rtol=1e-6)
    assert np.array_equal(w2, (s.reshape(-1, 1) * w).astype(np.float32))      # the fusion itself is exact
