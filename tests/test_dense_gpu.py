"""Token-major dense layer on tcgen05 through the BF16x3 split (csrc/dense_bf16.cu, dense.py) against
float64.  Error budget: the split drops <= 3 * 2^-18 per product, and the tensor core adds into its
fp32 TMEM accumulator with truncation, a bias of about n * 2^-25 for a chain of n MMAs (measured,
see hessian.cu) — n = 3K/16 here, i.e. 2.3e-5 at K = 4096 (the 3xTF32 gemm_tn it replaces: 3K/8).
Stated tolerance: max |Y - Y64| <= 2e-5 * max|Y64| for K <= 1024, 6e-5 for K up to 4096."""
import pytest
import torch

from onnx_quantize_b200 import dense as DN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,k,n,bias,relu", [(200, 96, 64, True, True), (1000, 4096, 256, False, False),
                                             (129, 100, 32, True, False), (4096, 512, 4096, True, True),
                                             (64, 64, 288, False, True), (300, 1156, 96, True, False)])
def test_dense_forward_matches_float64(cuda, m, k, n, bias, relu):
    g = torch.Generator(device=cuda)
    g.manual_seed(m + k + n)
    x = torch.randn((m, k), device=cuda, generator=g)
    x[:, 1] *= 30
    w = torch.randn((k, n), device=cuda, generator=g) / k ** 0.5
    b = torch.randn((n,), device=cuda, generator=g) if bias else None
    assert DN.supported(k, n)
    y = DN.dense_forward(x, DN.Planes.of_weight(w), b, relu)
    want = x.double() @ w.double()
    if bias:
        want = want + b.double()
    scale = want.abs().max().item()
    if relu:
        want = want.clamp(min=0)
    assert y.shape == (m, n)
    assert ((y.double() - want).abs().max().item() / scale) < (2e-5 if k <= 1024 else 6e-5)


def test_planes_are_reused_and_symmetric_row_operand(cuda):
    """A symmetric matrix as the row operand (the AWQ Gram product) and buffer reuse."""
    g = torch.Generator(device=cuda)
    g.manual_seed(3)
    a = torch.randn((256, 256), device=cuda, generator=g)
    s = a @ a.T
    d = torch.randn((256, 64), device=cuda, generator=g)
    ps = DN.Planes.of_rows(s)
    pd = DN.Planes.of_weight(d)
    y = DN.forward_planes(ps, pd, alpha=0.5)
    want = 0.5 * (s.double() @ d.double())
    assert ((y.double() - want).abs().max() / want.abs().max()).item() < 2e-5
    pd2 = DN.Planes.of_weight(2 * d, into=pd)
    assert pd2 is pd
    y2 = DN.forward_planes(ps, pd, alpha=0.5)
    assert ((y2.double() - 2 * want).abs().max() / want.abs().max()).item() < 4e-5


def test_unsupported_shapes_are_refused(cuda):
    x = torch.zeros((8, 16), device=cuda)
    w = torch.zeros((16, 20), device=cuda)          # N not a multiple of 32
    assert not DN.supported(16, 20)
    with pytest.raises(NotImplementedError):
        DN.dense_forward(x, DN.Planes.of_weight(w))
