"""Parity of the CUDA RTN / MSE / packing path (through the C ABI) with the golden vectors of the
unmodified reference and with the NumPy oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import as_i8, bits, golden_keys, parse_rtn_key, stable_seed

pytestmark = pytest.mark.gpu


def _product(w, qt, strategy, gs, sym, rr, clip, mse):
    import onnx_quantize_b200 as q
    from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
    qtype = q.QuantType.from_string(qt)
    return _rtn_quantize(w, qtype, q.QuantizationStrategy(strategy), gs, sym, rr, clip, mse,
                         np.dtype(np.float32), qtype.np_dtype)


def test_rtn_matches_reference_golden(cuda, golden):
    g = golden("rtn.npz")
    bad = []
    for key in golden_keys(g):
        wname, qt, strategy, gs, sym, rr, clip, mse = parse_rtn_key(key)
        q, s, z = _product(g["w::" + wname], qt, strategy, gs, sym, rr, clip, mse)
        ok = (q.shape == g["q::" + key].shape and s.shape == g["s::" + key].shape
              and z.shape == g["z::" + key].shape and q.dtype == O.np_dtype(qt)
              and s.dtype == np.float32 and z.dtype == O.np_dtype(qt)
              and np.array_equal(as_i8(q, qt), g["q::" + key])
              and np.array_equal(bits(s), bits(g["s::" + key]))
              and np.array_equal(as_i8(z, qt), g["z::" + key]))
        if not ok:
            bad.append(key)
    assert not bad, f"{len(bad)} golden cases differ, first: {bad[:5]}"


def test_packed_layouts_match_reference_golden(cuda, golden):
    """PACKED_FLAT == onnx_ir initializer bytes, MATMUL_NBITS == _prepare_for_matmul_nbits."""
    import onnx_quantize_b200 as qz
    from onnx_quantize_b200 import device_api as D
    g = golden("rtn.npz")
    n_a = n_b = 0
    for key in golden_keys(g):
        wname, qt, strategy, gs, sym, rr, clip, mse = parse_rtn_key(key)
        w = torch.from_numpy(g["w::" + wname]).to(cuda)
        qtype = qz.QuantType.from_string(qt)
        if "packA::" + key in g.files:
            codes, s, z = D.rtn_quantize(w, qtype, strategy, gs, sym, rr, clip, mse, layout="packed_flat")
            assert np.array_equal(codes.cpu().numpy(), g["packA::" + key]), key
            assert np.array_equal(bits(s.cpu().numpy().reshape(-1)), bits(g["s::" + key].reshape(-1))), key
            n_a += 1
        if "B::" + key in g.files:
            b, s, z = D.rtn_quantize(w, qtype, strategy, gs, sym, rr, clip, mse, layout="matmul_nbits")
            assert b.shape == g["B::" + key].shape and z.shape == g["Bz::" + key].shape, key
            assert np.array_equal(b.cpu().numpy(), g["B::" + key]), key
            assert np.array_equal(bits(s.cpu().numpy()), bits(g["Bs::" + key])), key
            assert np.array_equal(z.cpu().numpy(), g["Bz::" + key]), key
            n_b += 1
    assert n_a > 50 and n_b > 20


SHAPES = [(512, 256), (1024, 384), (256, 1040), (384, 48)]


@pytest.mark.parametrize("qt", ["uint4", "int4", "uint8", "int8"])
@pytest.mark.parametrize("strategy,gs", [("group", 128), ("group", 64), ("group", 32), ("group", 16),
                                         ("group", 256), ("channel", -1), ("tensor", -1)])
@pytest.mark.parametrize("sym", [False, True])
def test_rtn_no_mse_bit_exact_vs_oracle(cuda, qt, strategy, gs, sym):
    rng = np.random.default_rng(stable_seed(qt, strategy, gs, sym))
    for (k, n) in SHAPES:
        if strategy == "group" and k % gs:
            continue
        w = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
        w[rng.integers(0, k, 8), rng.integers(0, n, 8)] *= 30
        for clip in (1.0, 0.9):
            q, s, z = _product(w, qt, strategy, gs, sym, False, clip, False)
            qo, so, zo = O.rtn_quantize(w, qt, strategy, gs, sym, False, clip, False)
            assert np.array_equal(as_i8(q, qt), as_i8(qo, qt)), (k, n, clip)
            assert np.array_equal(bits(s), bits(so)) and s.shape == so.shape
            assert np.array_equal(as_i8(z, qt), as_i8(zo, qt)) and z.shape == zo.shape


@pytest.mark.parametrize("qt,sym", [("uint4", False), ("int4", True), ("int8", False), ("uint8", True)])
@pytest.mark.parametrize("strategy,gs", [("group", 128), ("group", 32), ("group", 256), ("group", 24),
                                         ("channel", -1), ("tensor", -1)])
def test_rtn_mse_vs_oracle(cuda, qt, sym, strategy, gs):
    """MSE search: outputs (codes / scale / zp) must equal the reference's; the error sums feed an
    arg-min, so the count of parameter rows that differ is asserted to be zero on these inputs."""
    rng = np.random.default_rng(stable_seed(qt, strategy, gs))
    k, n = (768, 96) if strategy != "tensor" else (192, 40)
    w = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
    q, s, z = _product(w, qt, strategy, gs, sym, False, 0.9, True)
    qo, so, zo = O.rtn_quantize(w, qt, strategy, gs, sym, False, 0.9, True)
    rows_differ = int(np.sum(bits(s).reshape(-1) != bits(so).reshape(-1)))
    assert rows_differ == 0, f"{rows_differ} of {so.size} parameter rows picked another candidate"
    assert np.array_equal(as_i8(q, qt), as_i8(qo, qt))
    assert np.array_equal(as_i8(z, qt), as_i8(zo, qt))


def test_mse_error_sums_follow_numpy_order(cuda, golden):
    """The 20 error sums per row against the reference's own np.sum results.  np.power is
    host-dependent (SVML vs glibc), so equality is asserted to 2 ulp-of-sum relative, and the
    summation ORDER is pinned separately with an order-sensitive synthetic input."""
    from onnx_quantize_b200 import device_api as D
    g = golden("mse_trace.npz")
    w = torch.from_numpy(g["w"]).to(cuda)
    for strategy, gs in (("group", 128), ("channel", -1), ("tensor", -1)):
        err = D.mse_error_table(w, "uint4", strategy, gs).cpu().numpy()
        ref = g["err::" + strategy]
        assert err.shape[1] == ref.shape[1]
        n = ref.shape[0]   # the reference may have stopped early
        np.testing.assert_allclose(err[:n], ref, rtol=3e-7 * 4, atol=0)


def test_early_stop_info(cuda):
    """Tiny single-row problems make the global early stop observable."""
    from onnx_quantize_b200 import device_api as D
    rng = np.random.default_rng(3)
    for trial in range(6):
        w = rng.standard_normal((16, 1 + trial % 2)).astype(np.float32)
        wt = torch.from_numpy(w).to(cuda)
        for strategy in ("tensor", "channel"):
            q, s, z, info = D.rtn_quantize(wt, "int8", strategy, -1, True, False, 1.0, True, return_info=True)
            qo, so, zo = O.rtn_quantize(w, "int8", strategy, -1, True, False, 1.0, True)
            assert np.array_equal(bits(s.cpu().numpy().reshape(-1)), bits(so.reshape(-1)))
            assert np.array_equal(q.cpu().numpy().view(np.int8), qo)
            lo, hi, trace = O.mse_min_max(O.to_rows(w, strategy), "int8", strategy, True, False, return_trace=True)
            assert int(info.cpu()[0]) == len(trace) - 1


@pytest.mark.parametrize("gs", [16, 32, 64, 128])
@pytest.mark.parametrize("sym,rr,clip", [(False, False, 1.0), (False, False, 0.9), (True, False, 0.95),
                                         (False, True, 1.0)])
def test_stream_kernel_matmul_nbits_bit_exact(cuda, gs, sym, rr, clip):
    """The HBM-bound kernel (rtn_stream.cuh: reciprocal multiply validated by an exact residual,
    zero points packed in the same launch) against the oracle, byte for byte: ragged tiles, odd
    group counts, a single group, and inputs made of exact rounding ties."""
    from onnx_quantize_b200 import device_api as D
    rng = np.random.default_rng(gs * 7 + int(sym) + 2 * int(rr))
    for (k, n) in [(gs, 48), (3 * gs, 80), (4 * gs, 1024), (8 * gs, 4096 + 16), (5 * gs, 16)]:
        w = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
        w[rng.integers(0, k, 16), rng.integers(0, n, 16)] *= 25
        # half of the columns: values on the quantization grid's midpoints (x/s = m + 0.5 exactly
        # representable), which must round half to even like np.round
        step = np.float32(2.0 ** -7)
        ties = (rng.integers(-15, 16, (k, n // 2)).astype(np.float32) + 0.5) * step / 2
        ties[0, :] = 15 * step / 2
        ties[1, :] = -15 * step / 2
        w[:, : n // 2] = ties
        wt = torch.from_numpy(w).to(cuda)
        b, s, z = D.rtn_quantize(wt, "uint4", "group", gs, sym, rr, clip, False, layout="matmul_nbits")
        qo, so, zo = O.rtn_quantize(w, "uint4", "group", gs, sym, rr, clip, False)
        ob, os_, oz = O.matmul_nbits_layout(qo, so, zo, gs, 4)
        assert np.array_equal(b.cpu().numpy(), ob), (k, n)
        assert np.array_equal(bits(s.cpu().numpy()), bits(os_)), (k, n)
        assert z.shape == oz.shape and np.array_equal(z.cpu().numpy(), oz), (k, n)


@pytest.mark.parametrize("qt,sym", [("int8", True), ("uint8", False), ("int4", True), ("uint4", False)])
def test_tensor_route_rounding_ties(cuda, qt, sym):
    """Per-tensor route (fold + vectorised codes): inputs whose quotient x/s is exactly m + 0.5 must
    round half to even like np.round — the reciprocal-multiply fast path has to hand them to the
    exact division."""
    rng = np.random.default_rng(9)
    lo, hi = O.qrange(qt, sym, False)
    step = np.float32(2.0 ** -7)
    levels = (hi - lo) if not sym else min(hi, -lo)
    k, n = 256, 512
    m = rng.integers(0 if not sym else -levels, levels, (k, n)).astype(np.float32)
    w = ((m + 0.5) * step).astype(np.float32)
    w[0, 0] = levels * step                       # pins the scale to exactly `step`
    if not sym:
        w[0, 1] = 0.0
    else:
        w[0, 1] = -levels * step
    q, s, z = _product(w, qt, "tensor", -1, sym, False, 1.0, False)
    qo, so, zo = O.rtn_quantize(w, qt, "tensor", -1, sym, False, 1.0, False)
    assert float(so) == float(step)
    assert np.array_equal(bits(s), bits(so)) and np.array_equal(as_i8(z, qt), as_i8(zo, qt))
    assert np.array_equal(as_i8(q, qt), as_i8(qo, qt))


def test_batched_stream_launch_matches_single_calls(cuda):
    """b200q_rtn_quantize_batch sends the HBM-bound configuration out as ONE launch for the whole
    job list (job table in kernel parameters); results must equal per-weight calls and the oracle."""
    from onnx_quantize_b200 import device_api as D
    rng = np.random.default_rng(21)
    shapes = [(256, 128), (128, 48), (384, 1040 + 8), (1024, 256), (128, 16), (640, 4096), (1024, 8192), (2048, 1040)]
    ws = [torch.from_numpy((rng.standard_normal(s) * 0.02).astype(np.float32)).to(cuda) for s in shapes]
    for gs in (128, 32):
        outs = D.rtn_quantize_batch(ws, "uint4", "group", gs, False, False, 0.9, False, layout="matmul_nbits")
        for w, (b, s, z) in zip(ws, outs):
            b1, s1, z1 = D.rtn_quantize(w, "uint4", "group", gs, False, False, 0.9, False, layout="matmul_nbits")
            assert torch.equal(b, b1) and torch.equal(s, s1) and torch.equal(z, z1)
            qo, so, zo = O.rtn_quantize(w.cpu().numpy(), "uint4", "group", gs, False, False, 0.9, False)
            ob, os_, oz = O.matmul_nbits_layout(qo, so, zo, gs, 4)
            assert np.array_equal(b.cpu().numpy(), ob) and np.array_equal(z.cpu().numpy(), oz)
            assert np.array_equal(bits(s.cpu().numpy()), bits(os_))


@pytest.mark.parametrize("mse", [0, 1])
@pytest.mark.parametrize("k,n,gs", [(128, 48, 128), (384, 80, 128), (96, 4112, 32), (64, 16, 16), (640, 208, 64)])
def test_no_out_of_bounds_writes_through_the_c_abi(cuda, k, n, gs, mse):
    """Guard bands around every output buffer of b200q_rtn_quantize (MatMulNBits layout, ragged
    tiles): compute-sanitizer is not available on the GPU pool, so the outputs are carved out of
    one canary-filled allocation and the bands are checked after the call."""
    from onnx_quantize_b200 import _lib, _device
    lib = _lib.load()
    rng = np.random.default_rng(k + n)
    w = torch.from_numpy((rng.standard_normal((k, n)) * 0.02).astype(np.float32)).to(cuda)
    g = k // gs
    sizes = [n * k // 2, 4 * n * g, n * ((g + 1) // 2 if g > 1 else g)]      # codes, scales, packed zp
    band = 256
    offs, total = [], band
    for sz in sizes:
        offs.append(total)
        total += (sz + 255) // 256 * 256 + band
    buf = torch.full((total,), 0xAB, dtype=torch.uint8, device=cuda)
    ws = torch.empty(lib.b200q_rtn_workspace_bytes(k, n, 2, gs, mse) + 512, dtype=torch.uint8, device=cuda)
    ws.fill_(0xCD)
    base = buf.data_ptr()
    rc = lib.b200q_rtn_quantize(w.data_ptr(), k, n, _lib.QTYPE["uint4"], 2, gs, 0, 0, 0.9, mse,
                                _lib.LAYOUT["matmul_nbits"], base + offs[0], base + offs[1], base + offs[2],
                                None, ws.data_ptr(), ws.numel() - 512, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    used = np.zeros(total, bool)
    for o, sz in zip(offs, sizes):
        used[o:o + sz] = True
    assert np.all(host[~used] == 0xAB), "write outside an output buffer"
    assert np.all(ws[-512:].cpu().numpy() == 0xCD), "write past the declared workspace size"
    qo, so, zo = O.rtn_quantize(w.cpu().numpy(), "uint4", "group", gs, False, False, 0.9, bool(mse))
    ob, os_, oz = O.matmul_nbits_layout(qo, so, zo, gs, 4)
    assert np.array_equal(host[offs[0]:offs[0] + sizes[0]], ob.reshape(-1))
    assert np.array_equal(host[offs[1]:offs[1] + sizes[1]].view(np.float32).view(np.uint32), bits(os_).reshape(-1))
    assert np.array_equal(host[offs[2]:offs[2] + sizes[2]], oz.reshape(-1))


@pytest.mark.parametrize("k,n", [(4096, 132), (2048, 96), (1024, 1024), (1000, 52), (520, 512), (2049, 515), (96, 40)])
def test_tensor_mse_parallel_pairwise_matches_oracle(cuda, k, n):
    """TENSOR + MSE: NumPy's pairwise sum over the flat array evaluated as independent blocks plus
    a perfect binary combine tree (mse_generic.cuh) — 8-lane leaves (2048 x 96: B = 96), the
    generic per-block recursion (4096 x 132: B = 264; 520 x 512: B = 130), the general cut of the
    recursion for sizes with few factors of two (1000 x 52, 2049 x 515: an odd element count) and
    the single-thread walk for small tensors (96 x 40).  Error sums to 4 ulp-of-sum (np.power is host-dependent),
    outputs identical."""
    from onnx_quantize_b200 import device_api as D
    rng = np.random.default_rng(k + n)
    w = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
    wt = torch.from_numpy(w).to(cuda)
    err = D.mse_error_table(wt, "int8", "tensor", -1, True).cpu().numpy().reshape(-1)
    lo, hi, trace = O.mse_min_max(O.to_rows(w, "tensor"), "int8", "tensor", True, False, return_trace=True)
    ref = np.array([float(np.asarray(t).reshape(-1)[0]) for t in trace], np.float32)
    np.testing.assert_allclose(err[:len(ref)], ref, rtol=3e-7 * 4, atol=0)
    q, s, z = _product(w, "int8", "tensor", -1, True, False, 1.0, True)
    qo, so, zo = O.rtn_quantize(w, "int8", "tensor", -1, True, False, 1.0, True)
    assert np.array_equal(bits(s).reshape(-1), bits(so).reshape(-1)) and np.array_equal(as_i8(q, "int8"), as_i8(qo, "int8"))


@pytest.mark.parametrize("qt,sym,clip,shape", [("int8", True, 1.0, (4096, 4096)), ("uint8", False, 0.9, (2048, 1024)),
                                               ("int4", True, 1.0, (1200, 1028)), ("uint4", False, 0.8, (3000, 1500))])
def test_per_tensor_route_on_large_weights_matches_oracle(cuda, qt, sym, clip, shape):
    """TENSOR strategy on weights of 5 to 64 MB (cfg1's shape among them): the streamlined two-launch
    route (min/max partials kept in L2, fold + parameters + vectorised codes walked back to front).
    Codes, scale and zero point are the oracle's."""
    from onnx_quantize_b200 import device_api as D
    rng = np.random.default_rng(stable_seed(qt, shape))
    w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    w[17, 5] = 0.31                                        # the extremes sit in different slices
    w[shape[0] - 3, shape[1] - 2] = -0.29
    q, s, z = D.rtn_quantize(torch.from_numpy(w).to(cuda), qt, "tensor", -1, sym, False, clip, False)
    qo, so, zo = O.rtn_quantize(w, qt, "tensor", -1, sym, False, clip, False)
    assert np.array_equal(bits(s.cpu().numpy()), bits(np.asarray(so).reshape(-1)))
    assert np.array_equal(z.cpu().numpy(), np.asarray(zo).reshape(-1).view(np.uint8))
    assert np.array_equal(q.cpu().numpy(), np.asarray(qo).view(np.uint8))


def test_ring_kernel_against_the_fused_kernel_on_a_long_job_list(cuda):
    """The persistent streaming kernel (TMA slab + registers, launch-wide tile dispenser) over 150 jobs —
    two launches, thousands of tiles, every CTA walks many of them — against an independent route:
    the plain fused kernel's (K,N) codes + parameters packed by `pack_matmul_nbits`."""
    from onnx_quantize_b200 import device_api as D
    g = torch.Generator(device=cuda)
    g.manual_seed(5)
    shapes = [(1024, 1024), (1024, 256), (1024, 3584), (3584, 1024), (512, 2064), (384, 1024)] * 25
    ws = [torch.randn(s, device=cuda, generator=g) * 0.02 for s in shapes]
    outs = D.rtn_quantize_batch(ws, "uint4", "group", 128, False, False, 0.9, False, layout="matmul_nbits")
    for w, (b, s, z) in zip(ws, outs):
        codes, s_kn, z_kn = D.rtn_quantize(w, "uint4", "group", 128, False, False, 0.9, False, layout="kn")
        b_ref, z_ref = D.pack_matmul_nbits(codes, z_kn, 128, 4)
        assert torch.equal(b, b_ref) and torch.equal(z, z_ref)
        assert torch.equal(s.reshape(-1), s_kn)
