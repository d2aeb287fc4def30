"""Guard bands around the outputs and workspaces of the entry points added in the second half of
round 1, called through the raw C ABI with ragged shapes (compute-sanitizer is closed on the GPU
pool): every output is carved out of a canary-filled allocation, the declared workspace is followed
by a canary tail, and both are checked after the call."""
import numpy as np
import pytest
import torch

from onnx_quantize_b200 import _device, _lib

pytestmark = pytest.mark.gpu
BAND = 512


class Arena:
    def __init__(self, sizes, cuda):
        self.sizes, self.offs, total = list(sizes), [], BAND
        for sz in self.sizes:
            self.offs.append(total)
            total += (sz + 255) // 256 * 256 + BAND
        self.buf = torch.full((total,), 0xAB, dtype=torch.uint8, device=cuda)

    def ptr(self, i):
        return self.buf.data_ptr() + self.offs[i]

    def check(self):
        host = self.buf.cpu().numpy()
        used = np.zeros(host.shape[0], bool)
        for o, sz in zip(self.offs, self.sizes):
            used[o:o + sz] = True
        assert np.all(host[~used] == 0xAB), "write outside an output buffer"
        return [host[o:o + sz] for o, sz in zip(self.offs, self.sizes)]


def workspace(nbytes, cuda):
    ws = torch.full((nbytes + 1024,), 0xCD, dtype=torch.uint8, device=cuda)
    return ws, lambda: bool(torch.all(ws[nbytes:] == 0xCD))


@pytest.mark.parametrize("k,n,gs,iters", [(128, 40, 64, 6), (96, 132, -1, 4), (512, 24, 256, 5), (64, 1, 16, 3)])
def test_hqq_stays_inside_its_buffers(cuda, k, n, gs, iters):
    lib = _lib.load()
    rng = np.random.default_rng(k + n)
    w = torch.from_numpy((rng.standard_normal((k, n)) * 0.05).astype(np.float32)).to(cuda)
    g = 1 if gs == -1 else k // gs
    rows = n * g
    a = Arena([k * n, 4 * rows, 4 * rows, 4, 8 * iters], cuda)
    nb = lib.b200q_hqq_workspace_bytes(k, n, gs, 0, iters)
    ws, tail_ok = workspace(nb, cuda)
    rc = lib.b200q_hqq_quantize(w.data_ptr(), k, n, 1, gs, 0, 1.0, 0, 0.7, 10.0, 1.01, iters, 1, a.ptr(0), a.ptr(1),
                                a.ptr(2), a.ptr(3), a.ptr(4), ws.data_ptr(), nb, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    codes = a.check()[0]
    assert tail_ok() and codes.max() <= 15


@pytest.mark.parametrize("m,k,n", [(130, 100, 32), (1, 64, 96), (257, 68, 288), (128, 256, 256)])
def test_dense_forward_stays_inside_its_buffers(cuda, m, k, n):
    lib = _lib.load()
    g = torch.Generator(device=cuda)
    g.manual_seed(m + k + n)
    x = torch.randn((m, k), device=cuda, generator=g)
    w = torch.randn((k, n), device=cuda, generator=g)
    pa, pb = lib.b200q_dense_planes_bytes(m, k), lib.b200q_dense_planes_bytes(n, k)
    a = Arena([pa, pb, 4 * m * n], cuda)
    st = _device.stream_ptr()
    assert lib.b200q_dense_split_rows(x.data_ptr(), m, k, a.ptr(0), pa, st) == 0, lib.b200q_last_error()
    assert lib.b200q_dense_split_transposed(w.data_ptr(), k, n, a.ptr(1), pb, st) == 0, lib.b200q_last_error()
    assert lib.b200q_dense_forward_planes(a.ptr(0), m, a.ptr(1), n, k, 1.0, None, 0, a.ptr(2), n, st) == 0, \
        lib.b200q_last_error()
    torch.cuda.synchronize()
    y = torch.from_numpy(a.check()[2].view(np.float32).reshape(m, n).copy())
    want = x.double().cpu() @ w.double().cpu()
    assert ((y.double() - want).abs().max() / want.abs().max()).item() < 2e-5


@pytest.mark.parametrize("t,k", [(100, 96), (4100, 160), (64, 32), (700, 1152)])
def test_hessian_bf16x3_stays_inside_its_buffers(cuda, t, k):
    lib = _lib.load()
    g = torch.Generator(device=cuda)
    g.manual_seed(t + k)
    x = torch.randn((t, k), device=cuda, generator=g)
    a = Arena([4 * k * k], cuda)
    nb = lib.b200q_hessian_workspace_bytes(t, k, _lib.PRECISION["bf16x3"])
    ws, tail_ok = workspace(nb, cuda)
    rc = lib.b200q_hessian_accumulate(x.data_ptr(), t, k, 1.0 / t, 0.0, a.ptr(0), _lib.PRECISION["bf16x3"],
                                      ws.data_ptr(), nb, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    h = torch.from_numpy(a.check()[0].view(np.float32).reshape(k, k).copy())
    want = (x.double().T @ x.double()).cpu() / t
    assert tail_ok() and ((h.double() - want).abs().max() / want.abs().max()).item() < 2e-5


@pytest.mark.parametrize("k,n,strategy,gs", [(256, 36, 1, -1), (512, 132, 1, -1), (768, 20, 2, 256), (96, 8, 2, 32),
                                             (320, 260, 2, 160)])
def test_slab_route_stays_inside_its_buffers(cuda, k, n, strategy, gs):
    """CHANNEL / groups of >= 32 rows with N % 4 == 0: the slab kernels (rtn_generic.cuh)."""
    from oracle import np_oracle as O
    lib = _lib.load()
    rng = np.random.default_rng(k * 7 + n)
    wn = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
    w = torch.from_numpy(wn).to(cuda)
    rows = n if strategy == 1 else n * (k // gs)
    a = Arena([k * n, 4 * rows, rows], cuda)
    nb = lib.b200q_rtn_workspace_bytes(k, n, strategy, gs, 0)
    ws, tail_ok = workspace(nb, cuda)
    rc = lib.b200q_rtn_quantize(w.data_ptr(), k, n, _lib.QTYPE["int8"], strategy, gs, 1, 0, 1.0, 0, _lib.LAYOUT["kn"],
                                a.ptr(0), a.ptr(1), a.ptr(2), None, ws.data_ptr(), nb, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    codes, scale, zp = a.check()
    assert tail_ok()
    qo, so, zo = O.rtn_quantize(wn, "int8", "channel" if strategy == 1 else "group", gs, True, False, 1.0, False)
    assert np.array_equal(codes.view(np.int8).reshape(k, n), qo.astype(np.int8))
    assert np.array_equal(scale.view(np.uint32), np.ascontiguousarray(so, np.float32).reshape(-1).view(np.uint32))


# ---- round 2 kernels --------------------------------------------------------------------------------
def test_ring_stream_kernel_stays_inside_its_buffers(cuda):
    """The persistent streaming kernel through b200q_rtn_quantize_batch: ragged last column tiles (TMA
    zero-fills beyond N, nothing may be stored there), odd group counts (padded zero-point nibble),
    CTAs that walk several tiles; every job's three outputs and the workspace (tile counters) sit
    between canaries."""
    from oracle import np_oracle as O
    lib = _lib.load()
    rng = np.random.default_rng(3)
    shapes = [(384, 1040), (128, 48), (640, 4112), (256, 16), (1152, 2064), (128, 128)]
    gs = 128
    ws_np = [(rng.standard_normal(s) * 0.02).astype(np.float32) for s in shapes]
    wts = [torch.from_numpy(w).to(cuda) for w in ws_np]
    sizes = []
    for k, n in shapes:
        g = k // gs
        sizes += [n * g * (gs // 2), 4 * n * g, n * ((g + 1) // 2 if g > 1 else g)]
    a = Arena(sizes, cuda)
    jobs = (_lib.RtnJob * len(shapes))()
    for i, ((k, n), w) in enumerate(zip(shapes, wts)):
        jobs[i] = _lib.RtnJob(w.data_ptr(), k, n, a.ptr(3 * i), a.ptr(3 * i + 1), a.ptr(3 * i + 2), None)
    nb = lib.b200q_rtn_batch_workspace_bytes(jobs, len(shapes), 2, gs, 0)
    assert nb > 0
    ws, tail_ok = workspace(nb, cuda)
    rc = lib.b200q_rtn_quantize_batch(jobs, len(shapes), _lib.QTYPE["uint4"], 2, gs, 0, 0, 0.9, 0,
                                      _lib.LAYOUT["matmul_nbits"], ws.data_ptr(), nb, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    outs = a.check()
    assert tail_ok()
    for i, w in enumerate(ws_np):
        qo, so, zo = O.rtn_quantize(w, "uint4", "group", gs, False, False, 0.9, False)
        ob, os_, oz = O.matmul_nbits_layout(qo, so, zo, gs, 4)
        assert np.array_equal(outs[3 * i], ob.reshape(-1)) and np.array_equal(outs[3 * i + 2], oz.reshape(-1))
        assert np.array_equal(outs[3 * i + 1].view(np.uint32), np.ascontiguousarray(os_, np.float32).reshape(-1).view(np.uint32))


@pytest.mark.parametrize("k,n,qt,strategy,gs,mse", [(300, 70, "int8", 1, -1, 1), (1000, 33, "uint4", 1, -1, 1),
                                                    (257, 67, "uint4", 2, -1, 0), (96, 50, "int4", 2, 24, 0),
                                                    (520, 131, "int8", 1, -1, 0), (144, 37, "uint8", 2, 48, 0)])
def test_channel_mse_and_ragged_routes_stay_inside_their_buffers(cuda, k, n, qt, strategy, gs, mse):
    """CHANNEL + MSE in two tiers (slab partials, pair list, exact tier) and the column walkers of
    ragged shapes (N % 4 != 0, group sizes that are not multiples of 32)."""
    from oracle import np_oracle as O
    lib = _lib.load()
    rng = np.random.default_rng(k + 3 * n)
    wn = (rng.standard_normal((k, n)) * 0.02).astype(np.float32)
    w = torch.from_numpy(wn).to(cuda)
    g = 1 if strategy == 1 or gs == -1 else k // gs
    rows = n * g
    a = Arena([k * n, 4 * rows, rows], cuda)
    nb = lib.b200q_rtn_workspace_bytes(k, n, strategy, gs, mse)
    ws, tail_ok = workspace(nb, cuda)
    sym = int(qt.startswith("int"))
    rc = lib.b200q_rtn_quantize(w.data_ptr(), k, n, _lib.QTYPE[qt], strategy, gs, sym, 0, 1.0, mse, _lib.LAYOUT["kn"],
                                a.ptr(0), a.ptr(1), a.ptr(2), None, ws.data_ptr(), nb, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    codes, scale, zp = a.check()
    assert tail_ok()
    qo, so, zo = O.rtn_quantize(wn, qt, "channel" if strategy == 1 else "group", gs, bool(sym), False, 1.0, bool(mse))
    assert np.array_equal(codes.reshape(k, n), np.asarray(qo).view(np.uint8))
    assert np.array_equal(scale.view(np.uint32), np.ascontiguousarray(so, np.float32).reshape(-1).view(np.uint32))


@pytest.mark.parametrize("pairs", ["0", "1"])
@pytest.mark.parametrize("t,k", [(100, 96), (1100, 352), (2050, 800)])
def test_both_hessian_kernels_stay_inside_their_buffers(cuda, monkeypatch, pairs, t, k):
    monkeypatch.setenv("B200Q_HESSIAN_PAIRS", pairs)
    lib = _lib.load()
    g = torch.Generator(device=cuda)
    g.manual_seed(t + k)
    x = torch.randn((t, k), device=cuda, generator=g)
    a = Arena([4 * k * k], cuda)
    nb = lib.b200q_hessian_workspace_bytes(t, k, _lib.PRECISION["bf16x3"])
    ws, tail_ok = workspace(nb, cuda)
    rc = lib.b200q_hessian_accumulate(x.data_ptr(), t, k, 1.0 / t, 0.0, a.ptr(0), _lib.PRECISION["bf16x3"],
                                      ws.data_ptr(), nb, _device.stream_ptr())
    assert rc == 0, lib.b200q_last_error()
    torch.cuda.synchronize()
    h = torch.from_numpy(a.check()[0].view(np.float32).reshape(k, k).copy())
    want = (x.double().T @ x.double()).cpu() / t
    assert tail_ok() and ((h.double() - want).abs().max() / want.abs().max()).item() < 2e-5
