"""Hessian accumulation on tcgen05 (kind::tf32): against a float64 torch reference (this is the one
floating-point contraction of the path; tolerance stated per precision mode)."""
import numpy as np
import pytest
import torch

from onnx_quantize_b200.hessian import HessianAccumulator, hessian_accumulate
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

# max |H - H64| / max|H64| per precision mode
TOL = {"fp32": 2e-6, "tf32x3": 2e-5, "bf16x3": 2e-5, "tf32": 3e-3}


def _ref(x, alpha):
    x64 = x.double().reshape(-1, x.shape[-1])
    return alpha * (x64.T @ x64)


@pytest.mark.parametrize("precision", ["fp32", "tf32", "tf32x3", "bf16x3"])
@pytest.mark.parametrize("t,k", [(256, 128), (4096, 512), (1000, 1152), (8192, 1024), (77, 96), (300, 100),
                                  (20000, 256)])
def test_hessian_matches_float64(cuda, precision, t, k):
    g = torch.Generator(device=cuda)
    g.manual_seed(t * 7 + k)
    x = torch.randn((t, k), device=cuda, generator=g)
    x[:, 3] *= 10
    x[:, 5] = 0
    h = torch.full((k, k), 0.5, device=cuda)
    hessian_accumulate(x, h, alpha=2.0 / 7, beta=0.25, precision=precision)
    want = _ref(x, 2.0 / 7) + 0.125
    err = ((h.double() - want).abs().max() / want.abs().max()).item()
    assert err < TOL[precision], err
    assert torch.equal(h, h.T)          # exactly symmetric by construction
    assert torch.all(h[5, :] == 0.125)  # dead channel stays exactly beta*H


def test_streaming_accumulator_matches_oracle(cuda):
    rng = np.random.default_rng(0)
    k = 256
    batches = [rng.standard_normal((4, 33, k)).astype(np.float32) for _ in range(5)]
    acc = HessianAccumulator(k, precision="tf32x3")
    h = np.zeros((k, k), np.float32)
    n = 0
    for b in batches:
        acc.add(b)
        h, n = O.accumulate_hessian(b, h, n)
    assert acc.num_samples == n
    got = acc.h.cpu().numpy()
    assert np.abs(got - h).max() / np.abs(h).max() < 2e-5


@pytest.mark.parametrize("pairs", ["0", "1"])
@pytest.mark.parametrize("t,k", [(4096, 512), (3000, 1152), (8192, 6912), (2048, 96), (5000, 800)])
def test_bf16x3_single_cta_and_cta_pair_kernels(cuda, monkeypatch, pairs, t, k):
    """Both tcgen05 kernels of the BF16x3 mode — 128 x 256 tiles on one CTA (cta_group::1) and
    256 x 256 tiles on a CTA pair (cta_group::2, each CTA staging its half of both operands) — against
    float64, whatever the size heuristic would pick; ragged K (not a multiple of 256) and T."""
    monkeypatch.setenv("B200Q_HESSIAN_PAIRS", pairs)
    g = torch.Generator(device=cuda)
    g.manual_seed(t + k)
    x = torch.randn((t, k), device=cuda, generator=g)
    x[:, 7] = 0
    h = torch.full((k, k), 0.25, device=cuda)
    hessian_accumulate(x, h, alpha=2.0 / 16, beta=1.0, precision="bf16x3")
    want = _ref(x, 2.0 / 16) + 0.25
    err = ((h.double() - want).abs().max() / want.abs().max()).item()
    assert err < TOL["bf16x3"], err
    assert torch.equal(h, h.T)
    assert torch.all(h[7, :] == 0.25)
