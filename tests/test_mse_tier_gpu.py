"""The two-tier MSE search: (1) the measured error bound of the approximate |d|^2.4 that its
proof rests on, (2) identity of its results with the exact evaluation of every candidate at the
benchmark's shapes (size-independent property: same codes, scales, zero points, packed bytes)."""
import numpy as np
import pytest
import torch

from onnx_quantize_b200 import device_api as D
from oracle import np_oracle as O
from tests.helpers import bits, stable_seed

pytestmark = pytest.mark.gpu


def test_pow_approx_error_bound(cuda):
    """delta budget of rtn_fused.cuh::kTierTau assumes <= 1e-5 relative error over the range of
    residuals a quantization error can take (|d| from 1e-12 to 1e4)."""
    g = torch.Generator(device=cuda)
    g.manual_seed(0)
    worst = 0.0
    for lo, hi in ((-12, -6), (-6, -3), (-3, 0), (0, 4)):
        e = torch.rand(8_000_000, device=cuda, generator=g, dtype=torch.float64) * (hi - lo) + lo
        x = (10.0 ** e).to(torch.float32)
        approx = D.debug_pow_approx(x).to(torch.float64)
        exact = x.to(torch.float64) ** float(np.float32(2.4))
        rel = ((approx - exact).abs() / exact).max().item()
        worst = max(worst, rel)
    assert worst < 1e-5, worst
    z = D.debug_pow_approx(torch.zeros(4, device=cuda))
    assert torch.all(z == 0)


@pytest.mark.parametrize("qt,sym,gs,shape", [
    ("uint4", False, 128, (4096, 4096)), ("uint4", False, 128, (14336, 1024)),
    ("int4", True, 128, (2048, 4096)), ("uint4", True, 64, (1024, 2048)),
    ("int8", False, 32, (1024, 1024)), ("uint8", False, 16, (512, 2048)), ("int4", False, 16, (512, 512)),
])
def test_two_tier_equals_exact(cuda, qt, sym, gs, shape):
    g = torch.Generator(device=cuda)
    g.manual_seed(stable_seed(qt, gs, shape))
    w = torch.randn(shape, device=cuda, generator=g) * 0.02
    w.view(-1)[torch.randint(0, w.numel(), (2000,), device=cuda, generator=g)] *= 20
    layout = "matmul_nbits" if qt.startswith("u") else ("packed_flat" if "4" in qt else "kn")
    a = D.rtn_quantize(w, qt, "group", gs, sym, False, 1.0, True, layout=layout, return_info=True)
    b = D.rtn_quantize(w, qt, "group", gs, sym, False, 1.0, "exact", layout=layout, return_info=True)
    for ta, tb in zip(a[:3], b[:3]):
        assert torch.equal(ta, tb)
    ia, ib = a[3].tolist(), b[3].tolist()
    assert ia[0] == ib[0] and (ia[1] & ~ib[1]) == 0   # same stop index; proven improvements are real


def test_two_tier_tiny_inputs_take_the_exact_route(cuda):
    """With a single tile the 'some row improved' mask cannot be proven full: the fallback kernel
    must reproduce the reference's early stop."""
    rng = np.random.default_rng(1)
    for k, n in ((16, 16), (32, 16), (128, 16), (64, 32)):
        w = (rng.standard_normal((k, n)) * 0.1).astype(np.float32)
        wt = torch.from_numpy(w).to(cuda)
        q, s, z, info = D.rtn_quantize(wt, "uint4", "group", 16, False, False, 1.0, True, return_info=True)
        qo, so, zo = O.rtn_quantize(w, "uint4", "group", 16, False, False, 1.0, True)
        rows = O.to_rows(w, "group", 16)
        _, _, trace = O.mse_min_max(rows, "uint4", "group", False, False, return_trace=True)
        assert int(info[0]) == len(trace) - 1
        assert np.array_equal(bits(s.cpu().numpy()), bits(so.reshape(-1)))
        assert np.array_equal(q.cpu().numpy(), np.asarray(qo).view(np.uint8))


def test_two_tier_adversarial_ties(cuda):
    """Groups engineered so that several candidates have (nearly) equal errors: constant groups,
    two-valued groups, groups of exact grid points, all-zero groups."""
    rng = np.random.default_rng(2)
    k, n = 1024, 256
    w = np.zeros((k, n), np.float32)
    w[:, 0:64] = rng.choice([-1.0, 1.0], size=(k, 64)).astype(np.float32) * 0.5      # two-valued
    w[:, 64:128] = (rng.integers(-7, 9, size=(k, 64)) * 0.125).astype(np.float32)    # grid points
    w[:, 128:192] = 0.37                                                               # constant
    w[:, 192:] = (rng.standard_normal((k, 64)) * 1e-20).astype(np.float32)           # underflowing errors
    wt = torch.from_numpy(w).to(cuda)
    for qt, sym in (("uint4", False), ("int4", True), ("int8", False)):
        q, s, z = D.rtn_quantize(wt, qt, "group", 128, sym, False, 1.0, True)
        qo, so, zo = O.rtn_quantize(w, qt, "group", 128, sym, False, 1.0, True)
        assert np.array_equal(bits(s.cpu().numpy()), bits(so.reshape(-1))), qt
        assert np.array_equal(q.cpu().numpy(), np.asarray(qo).view(np.uint8)), qt


# ---- CHANNEL strategy: two tiers over sequentially summed columns (csrc/mse_channel.cuh) ------------
@pytest.mark.parametrize("qt,sym,shape", [("int8", True, (1024, 160)), ("uint8", False, (4096, 64)),
                                          ("int4", True, (2048, 96)), ("uint4", False, (512, 320)),
                                          ("int8", False, (200, 33))])
def test_channel_two_tier_matches_oracle(cuda, qt, sym, shape):
    """Codes, scales and zero points of CHANNEL + MSE equal the oracle's (every candidate exact, NumPy's
    sequential order down the F-ordered column view) although only the pairs the interval
    classification could not decide are evaluated exactly."""
    rng = np.random.default_rng(stable_seed(qt, sym, shape))
    w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    w[rng.integers(0, shape[0], 40), rng.integers(0, shape[1], 40)] *= 15       # outliers: later optimum
    w[:, 3] = 0.0                                                                # all-zero column
    w[:, 5] *= np.float32(1e-15)                                                 # below the approximation's floor
    w[:, 7] = np.float32(0.25)                                                   # constant column
    wt = torch.from_numpy(w).to(cuda)
    q, s, z, info = D.rtn_quantize(wt, qt, "channel", -1, sym, False, 1.0, True, return_info=True)
    qo, so, zo = O.rtn_quantize(w, qt, "channel", -1, sym, False, 1.0, True)
    _, _, trace = O.mse_min_max(O.to_rows(w, "channel"), qt, "channel", sym, False, return_trace=True)
    assert int(info[0]) == len(trace) - 1                                        # same early-stop step
    assert np.array_equal(bits(s.cpu().numpy()), bits(np.asarray(so).reshape(-1)))
    assert np.array_equal(z.cpu().numpy(), np.asarray(zo).reshape(-1).view(np.uint8))
    assert np.array_equal(q.cpu().numpy(), np.asarray(qo).view(np.uint8))


def test_channel_two_tier_equals_all_exact_at_4096(cuda):
    """4096 x 4096 int8 / int4 per-channel with the search: the error table API evaluates all 20 x N pairs
    exactly (the round-1 route); the improvement masks derived from it must give the parameters the
    two-tier route returns."""
    g = torch.Generator(device=cuda)
    g.manual_seed(77)
    w = torch.randn((4096, 4096), device=cuda, generator=g) * 0.02
    for qt, sym in (("int8", True), ("int4", True), ("uint4", False)):
        _, s, z = D.rtn_quantize(w, qt, "channel", -1, sym, False, 1.0, True)
        err = D.mse_error_table(w, qt, "channel", -1, sym, False)                # (20, N) exact sums
        e = err.cpu().numpy()
        best = np.full(e.shape[1], np.inf, np.float32)
        masks = np.zeros(e.shape[1], np.uint32)
        for i in range(20):
            imp = e[i] < best
            best = np.where(imp, e[i], best)
            masks |= imp.astype(np.uint32) << np.uint32(i)
        stalls, stop = 0, 19
        any_imp = [bool(((masks >> np.uint32(i)) & 1).any()) for i in range(20)]
        for i in range(20):
            stalls += 0 if any_imp[i] else 1
            if stalls >= 5:
                stop = i
                break
        masks &= np.uint32((2 << stop) - 1)
        best_i = np.where(masks != 0, np.floor(np.log2(np.maximum(masks, 1))).astype(np.int64), 0)
        lo, hi = D.row_ranges(w, qt, "channel", -1, sym, False, 1.0, mse=False)
        lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
        p = (1.0 - best_i / 100.0).astype(np.float32)
        so, zo = O.qparams(p * np.minimum(lo, 0), p * np.maximum(hi, 0), qt, sym, False)
        assert np.array_equal(bits(s.cpu().numpy()), bits(np.asarray(so).reshape(-1))), qt


@pytest.mark.parametrize("magnitude", [1e-9, 1e-11, 1e-13, 1e-15, 1e-25, 1e-36])
def test_two_tier_small_magnitudes(cuda, magnitude):
    """Quantization steps so small that |d|^2.4 leaves the range on which the approximation's error
    bound was measured (|d| >= 1e-12; lg2.approx / ex2.approx flush denormals): the result must still
    be the reference's, for the group kernels and the channel route alike."""
    rng = np.random.default_rng(stable_seed("tiny", magnitude))
    w = (rng.standard_normal((512, 128)) * magnitude).astype(np.float32)
    wt = torch.from_numpy(w).to(cuda)
    for qt, sym, strategy, gs in (("uint4", False, "group", 128), ("int4", True, "group", 64), ("int8", False, "group", 32),
                                  ("int8", True, "channel", -1), ("uint4", False, "channel", -1)):
        q, s, z = D.rtn_quantize(wt, qt, strategy, gs, sym, False, 1.0, True)
        qo, so, zo = O.rtn_quantize(w, qt, strategy, gs, sym, False, 1.0, True)
        assert np.array_equal(bits(s.cpu().numpy()), bits(np.asarray(so).reshape(-1))), (qt, strategy, magnitude)
        assert np.array_equal(q.cpu().numpy(), np.asarray(qo).view(np.uint8)), (qt, strategy, magnitude)
