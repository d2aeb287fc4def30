"""The NumPy-facing mirrors of the reference API (utils / pack / _common / calibration), run on the
GPU and compared with the golden vectors, the oracle and the reference's known answers.  These
read like the reference's own tests (test/core/algorithms/test_rtn.py, test/core/test_pack.py,
test/qrules/test_common.py, test/core/calibration/test_minmax_calibrator.py)."""
import math

import numpy as np
import pytest
import torch

import onnx_quantize_b200 as oq
from onnx_quantize_b200 import QuantizationStrategy, QuantType
from onnx_quantize_b200.core._algorithms.rtn import _quantize_bias, _rtn_quantize
from onnx_quantize_b200.core._algorithms.utils import (
    _compute_min_max,
    _compute_min_max_mse,
    _compute_qparams,
    _compute_qparams_from_array,
    _dequantize_array,
    _fake_quantize_array,
    _preprocess_array,
    _quantize_array_from_qparams,
)
from onnx_quantize_b200.core._calibration.factory import get_calibrator
from onnx_quantize_b200.core._calibration.minmax import MinMaxCalibrator
from onnx_quantize_b200.core._pack import pack, unpack
from onnx_quantize_b200.qrules._common import _prepare_for_matmul_nbits
from oracle import np_oracle as O
from tests.helpers import as_i8, bits
from tests.test_oracle_golden import PACK_TABLE, SCALE_ZP_TABLE

pytestmark = pytest.mark.gpu
QT = {"int4": QuantType.QInt4, "uint4": QuantType.QUInt4, "int8": QuantType.QInt8, "uint8": QuantType.QUInt8}


@pytest.mark.parametrize("vals,qt,sym,scale,zp", SCALE_ZP_TABLE)
@pytest.mark.parametrize("mse", [False, True])
def test_get_quantization_params_scalar(cuda, vals, qt, sym, scale, zp, mse):
    s, z = _compute_qparams_from_array(np.array(vals), QT[qt], QuantizationStrategy.TENSOR, -1, sym,
                                       False, 1.0, mse, np.float32, QT[qt].np_dtype)
    assert s > 0 and s.size == 1 and z.size == 1 and z.dtype == QT[qt].np_dtype
    np.testing.assert_allclose(s, np.array(scale, dtype=np.float32), rtol=1e-5)
    # the reference's table is evaluated on float64 inputs; the device path is float32 (as the
    # pipeline is), which moves exactly one tie: zp(-5,0,5) = round(-128 + 127.49999) = -1
    so, zo = O.qparams_from_rows(np.array(vals, dtype=np.float32), qt, "tensor", sym, False, 1.0, mse)
    assert np.array_equal(bits(s), bits(so)) and int(z) == int(zo)
    if vals != [-5.0, 0.0, 5.0] or sym:
        np.testing.assert_allclose(z, np.array(zp, dtype=np.float32), rtol=1e-5)


@pytest.mark.parametrize("mse", [False, True])
def test_qparams_per_channel_and_group_shapes(cuda, mse):
    a = np.array([[-5.0, 0.0, 10.0], [0.0, 0.0, 0.0]], dtype=np.float32)
    s, z = _compute_qparams_from_array(a, QuantType.QInt8, QuantizationStrategy.CHANNEL, -1, False,
                                       False, 1.0, mse, np.float32, np.dtype(np.int8))
    so, zo = O.qparams_from_rows(a, "int8", "channel", False, False, 1.0, mse)
    assert s.shape == (2, 1) and z.shape == (2, 1) and np.array_equal(bits(s), bits(so)) and np.array_equal(z, zo)
    w = np.ones((32, 64), dtype=np.float32)
    rows = _preprocess_array(w, QuantizationStrategy.GROUP, 16)
    s, z = _compute_qparams_from_array(rows, QuantType.QInt8, QuantizationStrategy.GROUP, 16, True, False,
                                       1.0, mse, np.float32, np.dtype(np.int8))
    assert s.shape == (64 * 2, 1) and np.all(s > 0)


@pytest.mark.parametrize("strategy,gs", [(QuantizationStrategy.TENSOR, -1), (QuantizationStrategy.CHANNEL, -1),
                                         (QuantizationStrategy.GROUP, 16)])
@pytest.mark.parametrize("reduce_range", [False, True])
def test_calculate_mse_min_max(cuda, rng, strategy, gs, reduce_range):
    x = rng.standard_normal((32, 64), dtype=np.float32)
    rows = _preprocess_array(x, strategy, gs)
    lo0, hi0 = _compute_min_max(rows, strategy=strategy, group_size=gs)
    lo, hi = _compute_min_max_mse(rows, QuantType.QInt8, strategy, gs, False, reduce_range, np.float32, np.int8)
    assert lo.shape == lo0.shape and np.all(lo >= lo0) and np.all(hi <= hi0) and np.all(lo <= hi)
    olo, ohi = O.mse_min_max(O.to_rows(x, strategy.value, gs), "int8", strategy.value, False, reduce_range)
    assert np.array_equal(bits(lo), bits(olo)) and np.array_equal(bits(hi), bits(ohi))
    olo0, ohi0 = O.row_min_max(O.to_rows(x, strategy.value, gs), strategy.value, 0.7)
    lo7, hi7 = _compute_min_max(rows, strategy, gs, clip_ratio=0.7)
    assert np.array_equal(bits(lo7), bits(olo0)) and np.array_equal(bits(hi7), bits(ohi0))


@pytest.mark.parametrize("qt,sym,rr", [("int8", False, False), ("int8", True, False), ("uint8", False, False),
                                       ("uint8", True, False), ("int8", False, True), ("uint4", False, False)])
@pytest.mark.parametrize("strategy,gs", [(QuantizationStrategy.TENSOR, -1), (QuantizationStrategy.CHANNEL, -1),
                                         (QuantizationStrategy.GROUP, 8)])
@pytest.mark.parametrize("mse", [False, True])
def test_quantize_array_invariants(cuda, rng, qt, sym, rr, strategy, gs, mse):
    """Shapes / dtypes / ranges / reconstruction bound of test_rtn.py:259-452."""
    k, n = 32, 64
    x = rng.standard_normal((k, n), dtype=np.float32)
    q, s, z = _rtn_quantize(x, QT[qt], strategy, gs, sym, rr, 1.0, mse, np.float32, QT[qt].np_dtype)
    assert q.shape == x.shape and q.dtype == QT[qt].np_dtype and s.dtype == np.float32 and z.dtype == QT[qt].np_dtype
    want = {QuantizationStrategy.TENSOR: (), QuantizationStrategy.CHANNEL: (n,),
            QuantizationStrategy.GROUP: (n * math.ceil(k / gs), 1)}[strategy]
    assert s.shape == want and z.shape == want
    lo, hi = QT[qt].qrange(sym, rr)
    qi, zi = q.astype(np.int32), z.astype(np.int32)
    assert qi.min() >= lo and qi.max() <= hi and zi.min() >= lo and zi.max() <= hi and np.all(s > 0)
    dq = _dequantize_array(q, s, z, preprocess=True, strategy=strategy, group_size=gs)
    assert dq.shape == x.shape and dq.dtype == np.float32
    assert np.array_equal(bits(dq), bits(O.dequantize_weight(q, s, z, strategy.value, gs)))
    if qt != "uint4" and not rr:
        assert np.max(np.abs(dq - x)) <= 2 * s.max()


def test_edge_cases(cuda):
    q, s, z = _rtn_quantize(np.zeros((4, 4), np.float32), QuantType.QInt8, QuantizationStrategy.TENSOR, -1,
                            False, False, 1.0, False, np.float32, np.int8)
    assert np.all(q == z) and s == 1.0
    x = np.full((3, 3), 5.0, np.float32)
    q, s, z = _rtn_quantize(x, QuantType.QInt8, QuantizationStrategy.TENSOR, -1, False, False, 1.0, False,
                            np.float32, np.int8)
    np.testing.assert_allclose(_dequantize_array(q, s, z), x, rtol=0.1)
    with pytest.raises(ValueError, match="does not divide"):
        _rtn_quantize(np.zeros((96, 8), np.float32), QuantType.QInt8, QuantizationStrategy.GROUP, 36, False,
                      False, 1.0, False, np.float32, np.int8)


def test_quantize_dequantize_fake_quantize_mirrors(cuda, rng):
    x = rng.standard_normal((24, 40), dtype=np.float32)
    for qt, sym in (("int8", True), ("uint4", False), ("int4", True)):
        rows = x.T   # per-channel rows (F-ordered view, like the reference's channel path)
        s, z = O.qparams_from_rows(rows, qt, "channel", sym, False, 1.0, False)
        want = O.quantize_rows(rows, s, z, qt, sym, False)
        got = _quantize_array_from_qparams(rows, s, z, QT[qt], sym, False)
        assert got.dtype == want.dtype and np.array_equal(as_i8(got, qt), as_i8(want, qt))
        assert np.array_equal(bits(_dequantize_array(got, s, z)), bits(O.dequantize(want, s, z)))
        fq = _fake_quantize_array(rows, s, z, QT[qt], sym, False)
        assert np.array_equal(bits(fq), bits(O.dequantize(want, s, z)))
        # one row with 0-d parameters, as the GPTQ row loop of the reference does (gptq.py:186-189)
        s1, z1 = O.qparams_from_rows(x, qt, "tensor", sym, False, 1.0, False)
        got = _quantize_array_from_qparams(x[3], s1, z1, QT[qt], sym, False)
        assert np.array_equal(as_i8(got, qt), as_i8(O.quantize_rows(x[3], s1, z1, qt, sym, False), qt))


def test_compute_qparams_matches_golden(cuda, golden):
    g = golden("minmax.npz")
    for m in (0.0, 0.5, 0.9):
        for qt, sym in (("int8", True), ("int8", False), ("uint8", False), ("uint8", True)):
            s, z = _compute_qparams(g[f"lo::{m}"], g[f"hi::{m}"], QT[qt], sym, False, np.float32, QT[qt].np_dtype)
            assert s.shape == () and np.array_equal(bits(s), bits(g[f"s::{m}|{qt}|{int(sym)}"]))
            assert z.dtype == QT[qt].np_dtype and np.array_equal(z, g[f"z::{m}|{qt}|{int(sym)}"])
    lo = -np.abs(np.random.default_rng(0).standard_normal(1000)).astype(np.float32) * 3
    hi = np.abs(np.random.default_rng(1).standard_normal(1000)).astype(np.float32) * 3
    for qt in QT:
        for sym in (False, True):
            for rr in (False, True):
                s, z = _compute_qparams(lo, hi, QT[qt], sym, rr, np.float32, QT[qt].np_dtype)
                so, zo = O.qparams(lo, hi, qt, sym, rr)
                assert np.array_equal(bits(s), bits(so)) and np.array_equal(as_i8(z, qt), as_i8(zo, qt))


def test_quantize_bias(cuda, rng, golden):
    bias = rng.random((16,)).astype(np.float32)
    ws = rng.random((16,)).astype(np.float32)
    q, s, z = _quantize_bias(bias, 1.5, ws)
    assert q.shape == bias.shape and q.dtype == np.int32 and z == 0
    np.testing.assert_array_equal(s, 1.5 * ws)
    g = golden("bias.npz")
    q, s, _ = _quantize_bias(g["bias"], 0.037, g["ws"])
    assert np.array_equal(q, g["q_vec"]) and np.array_equal(bits(s), bits(g["s_vec"]))
    q, s, _ = _quantize_bias(g["bias"], 0.037, g["ws"][:1])
    assert np.array_equal(q, g["q_one"]) and np.array_equal(bits(s), bits(g["s_one"]))


# ---- packing ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("qt,vals,packed", PACK_TABLE)
def test_pack_unpack_known_answers(cuda, qt, vals, packed):
    a = np.array(vals, dtype=np.int8 if qt == "int4" else np.uint8)
    p = pack(a, QT[qt])
    assert p.dtype == np.uint8 and p.tolist() == packed
    u = unpack(p, a.shape, QT[qt])
    assert u.dtype == a.dtype and np.array_equal(u, a)


def test_pack_shapes_and_passthrough(cuda):
    a = np.array([[[1, 2], [3, 4]], [[5, 6], [7, -1]]], dtype=np.int8)
    assert np.array_equal(unpack(pack(a, QuantType.QInt4), a.shape, QuantType.QInt4), a)
    big = np.random.default_rng(0).integers(-8, 8, size=1001, dtype=np.int8)
    assert np.array_equal(unpack(pack(big, QuantType.QInt4), big.shape, QuantType.QInt4), big)
    assert np.array_equal(pack(big, QuantType.QInt4), O.pack4_flat(big, "int4"))
    for qt, dt in ((QuantType.QInt8, np.int8), (QuantType.QUInt8, np.uint8), (QuantType.QInt32, np.int32)):
        arr = np.array([1, 2, 3, 4, 5], dtype=dt)
        np.testing.assert_array_equal(pack(arr, qt), arr)
        np.testing.assert_array_equal(unpack(arr, arr.shape, qt), arr)


def test_prepare_for_matmul_nbits_odd_num_blocks(cuda):
    gs, g, n = 16, 5, 4
    qconfig = oq.QConfig(weights=oq.QWeightArgs(dtype="uint4", strategy="group", group_size=gs))
    r = np.random.default_rng(0)
    w_q = r.integers(0, 16, size=(gs * g, n), dtype=np.uint8)
    w_scale = r.random(size=(n * g,)).astype(np.float32)
    w_zp = r.integers(0, 16, size=(n * g, 1), dtype=np.uint8)
    b, s, pz = _prepare_for_matmul_nbits(w_q, w_scale, w_zp, qconfig)
    ob, os_, oz = O.matmul_nbits_layout(w_q, w_scale, w_zp, gs, 4)
    assert pz.shape == (n, 3) and np.array_equal(pz, oz) and np.array_equal(b, ob) and np.array_equal(s, os_)
    nib = np.empty((n, 6), np.uint8)
    nib[:, ::2], nib[:, 1::2] = pz & 0xF, pz >> 4
    np.testing.assert_array_equal(nib[:, :g], w_zp.reshape(n, g))


def test_prepare_for_matmul_nbits_from_rtn_output(cuda, golden):
    g = golden("rtn.npz")
    for key in ("randn|uint4|group|128|0|0|0.9|1", "randn|uint8|group|32|0|0|1.0|0", "edge|uint4|group|32|0|0|1.0|0"):
        qt, gs = key.split("|")[1], int(key.split("|")[3])
        qc = oq.QConfig(weights=oq.QWeightArgs(dtype=qt, strategy="group", group_size=gs))
        q = g["q::" + key].view(QT[qt].np_dtype) if qt == "uint4" else g["q::" + key]
        z = g["z::" + key].view(QT[qt].np_dtype) if qt == "uint4" else g["z::" + key]
        b, s, pz = _prepare_for_matmul_nbits(q, g["s::" + key], z, qc)
        assert np.array_equal(b, g["B::" + key]) and np.array_equal(s, g["Bs::" + key]) and np.array_equal(pz, g["Bz::" + key])


# ---- MinMax calibrator ------------------------------------------------------------------------------
def test_minmax_calibrator_surface(cuda):
    c = MinMaxCalibrator()
    assert c.momentum == 0.0 and c.data == {}
    assert MinMaxCalibrator(momentum=0.99).momentum == 0.99
    for bad in (1.0, 1.5, -0.1):
        with pytest.raises(AssertionError, match="Momentum must be in"):
            MinMaxCalibrator(momentum=bad)
    with pytest.raises(KeyError, match="No calibration data collected for 'nonexistent'"):
        c.compute_range("nonexistent")
    with pytest.raises(TypeError, match="Invalid arguments for MinMaxCalibrator"):
        get_calibrator(oq.CalibrationMethod.MINMAX, bogus=1)


def test_minmax_collect_known_answers(cuda):
    c = MinMaxCalibrator()
    c.collect("test", np.array([1.0, 2.0, 3.0, 4.0, 5.0]))
    assert "test" in c.data and c.data["test"].min_val == 1.0 and c.data["test"].max_val == 5.0
    lo, hi = c.compute_range("test")
    assert lo == 0.0 and hi == 5.0 and lo.dtype == np.float32 and lo.shape == ()
    c = MinMaxCalibrator(momentum=0.0)
    for b in ([1.0, 2.0, 3.0], [-0.5, 4.0, 2.5], [1.5, 3.5, 5.5]):
        c.collect("t", np.array(b))
    assert c.data["t"].min_val == -0.5 and c.data["t"].max_val == 5.5
    c = MinMaxCalibrator(momentum=0.8)
    c.collect("t", np.array([-1.0, 2.0, 3.0]))
    assert c.data["t"].min_val == -1.0 and c.data["t"].max_val == 3.0
    c.collect("t", np.array([-0.5, 2.5, 4.0]))
    assert np.isclose(c.data["t"].min_val, -0.9) and np.isclose(c.data["t"].max_val, 3.2)
    c = MinMaxCalibrator()
    for name, vals in (("a", [-1.5, 2.0, 3.0]), ("b", [-1.0, 0.0, 1.0]), ("c", [10.0, 20.0, 30.0])):
        c.collect(name, np.array(vals))
    assert len(c.data) == 3 and c.data["c"].min_val == 10.0 and c.compute_range("c")[0] == 0.0
    c.collect("neg", np.array([-10.0, -5.0, -2.0, -1.0]))
    lo, hi = c.compute_range("neg")
    assert lo == -10.0 and hi == 0.0


def test_minmax_golden_and_ragged_sizes(cuda, golden):
    g = golden("minmax.npz")
    for m in (0.0, 0.5, 0.9):
        c = get_calibrator(oq.CalibrationMethod.MINMAX, momentum=m)
        for b in g["batches"]:
            c.collect("x", b)
        lo, hi = c.compute_range("x")
        assert np.array_equal(bits(lo), bits(g[f"lo::{m}"])) and np.array_equal(bits(hi), bits(g[f"hi::{m}"]))
    r = np.random.default_rng(5)
    from onnx_quantize_b200 import device_api as D
    for n in (1, 3, 4, 5, 127, 1024, 4099, 1 << 20, (1 << 22) + 7):
        x = r.standard_normal(n + 3).astype(np.float32) * 7
        for off in (0, 1, 3):   # unaligned starts
            t = torch.from_numpy(x).cuda()[off:off + n]
            pair = D.minmax_reduce(t.contiguous() if off == 0 else t).cpu().numpy()
            assert pair[0] == x[off:off + n].min() and pair[1] == x[off:off + n].max(), (n, off)
    c = MinMaxCalibrator(momentum=0.3)   # more batches than the pending-pair chunk
    o = O.MinMax(momentum=0.3)
    for i in range(150):
        b = r.standard_normal(257).astype(np.float32)
        c.collect("y", b)
        o.collect("y", b)
    assert np.array_equal(bits(c.compute_range("y")[0]), bits(o.compute_range("y")[0]))
    assert np.array_equal(bits(c.compute_range("y")[1]), bits(o.compute_range("y")[1]))


def test_large_pageable_arrays_take_the_staged_copies(cuda):
    """>= 8 MB inputs go up through the pinned staging chunks (several chunks, ragged tail) and the
    codes come back through pinned memory: results identical to the oracle, arrays writable."""
    import onnx_quantize_b200 as q
    from onnx_quantize_b200 import _device as dev
    from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
    from oracle import np_oracle as O
    rng = np.random.default_rng(21)
    w = (rng.standard_normal((4224, 4100)) * 0.02).astype(np.float32)        # 69 MB: three staging chunks
    up = dev.to_device_f32(w)
    assert torch.equal(up.cpu(), torch.from_numpy(w))
    assert torch.equal(dev.to_device_f32(w[::2]).cpu(), torch.from_numpy(np.ascontiguousarray(w[::2])))   # strided view
    codes, s, z = _rtn_quantize(w, q.QuantType.QUInt4, q.QuantizationStrategy.GROUP, 128, False, False, 0.9, False,
                                np.dtype(np.float32), q.QuantType.QUInt4.np_dtype)
    qo, so, zo = O.rtn_quantize(w, "uint4", "group", 128, False, False, 0.9, False)
    assert np.array_equal(codes.astype(np.uint8), qo.astype(np.uint8))
    assert np.array_equal(s.view(np.uint32), so.view(np.uint32)) and np.array_equal(z.astype(np.uint8), zo.astype(np.uint8))
    assert codes.flags.writeable and codes.shape == w.shape


def test_bulk_pipeline_with_pageable_weights(cuda):
    """quantize_weights_bulk on ordinary NumPy arrays (what an ONNX model's initializers are): the
    staged upload feeds the device slots; results equal the per-weight reference-facing call."""
    import onnx_quantize_b200 as q
    from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
    from onnx_quantize_b200.pipeline import RtnSpec, quantize_weights_bulk
    rng = np.random.default_rng(5)
    ws = [(rng.standard_normal(s) * 0.02).astype(np.float32) for s in ((2304, 1024), (256, 64), (1024, 2304), (128, 48))]
    spec = RtnSpec(q.QuantType.QUInt4, "group", 128, False, False, 0.9, True, "kn")
    got = quantize_weights_bulk(ws, spec)
    for w, (c, s, z) in zip(ws, got):
        qc, qs_, qz = _rtn_quantize(w, q.QuantType.QUInt4, q.QuantizationStrategy.GROUP, 128, False, False, 0.9, True,
                                    np.dtype(np.float32), q.QuantType.QUInt4.np_dtype)
        assert np.array_equal(c.reshape(w.shape), qc.astype(np.uint8))
        assert np.array_equal(s.reshape(-1).view(np.uint32), qs_.reshape(-1).view(np.uint32))


def test_calibrator_data_is_materialised_on_every_kind_of_read(cuda):
    """`.data` of the reference is a plain dict of real numbers; here it is filled lazily from the
    device, so every read path has to trigger the fold: get, copy, dict(), iteration, pickling — and
    entries a caller writes directly (the reference allows it) must work in compute_range."""
    import pickle

    from onnx_quantize_b200.core._calibration.base import CalibrationData
    from onnx_quantize_b200.core._calibration.minmax import MinMaxCalibrator
    c = MinMaxCalibrator()
    c.collect("a", np.array([[-2.0, 3.0]], np.float32))
    assert c.data.get("a").max_val == 3.0
    c.collect("a", np.array([[5.0]], np.float32))                    # pending again
    assert dict(c.data)["a"].max_val == 5.0
    c.collect("a", np.array([[7.0]], np.float32))
    assert c.data.copy()["a"].max_val == 7.0
    c.collect("a", np.array([[9.0]], np.float32))
    assert [v.max_val for v in c.data.values()] == [9.0]
    c.collect("a", np.array([[-11.0]], np.float32))
    assert pickle.loads(pickle.dumps(c.data))["a"].min_val == -11.0
    assert c.data.get("missing") is None and c.data.setdefault("a").min_val == -11.0
    c.data["manual"] = CalibrationData(np.float32(0.5), np.float32(2.0))
    lo, hi = c.compute_range("manual")
    assert lo == 0.0 and hi == 2.0
    with pytest.raises(KeyError):
        c.compute_range("nope")
