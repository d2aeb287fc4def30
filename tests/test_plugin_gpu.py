"""The drop-in boundary end to end: the registered algorithm plugins are called the
way the reference's rewriter calls them — `qconfig.weights.algorithm.quantize_weights(w, qconfig,
out=out)` (qrules/_common.py:133) with `w.const_value.numpy()` and `out.producer().meta["input"]`
— and must return what the reference's plugins return.  onnx_ir is not installed here, so `w` /
`out` are minimal stand-ins exposing exactly the attributes the plugins touch."""
import numpy as np
import pytest

import onnx_quantize_b200 as q
from onnx_quantize_b200.parallel import prequantized
from onnx_quantize_b200.parallel.shard import quantize_weights_sharded
from onnx_quantize_b200.pipeline import RtnSpec
from onnx_quantize_b200.qrules._common import _prepare_for_matmul_nbits, is_matmul_nbits_compatible
from oracle import np_oracle as O
from tests.helpers import as_i8, bits

pytestmark = pytest.mark.gpu


class _Const:
    def __init__(self, a):
        self._a = a

    def numpy(self):
        return self._a


class _Value:                      # ir.Value: .name, .const_value.numpy()
    def __init__(self, name, a=None, node=None):
        self.name, self.const_value, self._node = name, _Const(a), node

    def producer(self):
        return self._node


class _Node:                       # ir.Node: .meta
    def __init__(self, meta):
        self.meta = meta


@pytest.mark.parametrize("dtype,gs,sym,mse", [("uint4", 128, False, False), ("uint4", 128, False, True),
                                              ("int8", None, True, False), ("int4", 32, True, False)])
def test_rtn_plugin_matches_reference_semantics(cuda, rng, dtype, gs, sym, mse):
    w = (rng.standard_normal((256, 96)) * 0.02).astype(np.float32)
    w.setflags(write=False)                      # initializers arrive as read-only views
    cfg = q.QConfig(weights=q.QWeightArgs(dtype=dtype, group_size=gs, symmetric=sym, mse=mse, clip_ratio=0.9))
    back = q.QConfig(**cfg.model_dump())         # what the rewriter rebuilds from node.meta (qrules/base.py:57)
    got = back.weights.algorithm.quantize_weights(_Value("w0", w), back)
    strategy = back.weights.strategy.value
    want = O.rtn_quantize(w, dtype, strategy, gs or -1, sym, False, 0.9, mse)
    assert got[0].dtype == back.weights.dtype.np_dtype
    assert np.array_equal(as_i8(got[0], dtype), as_i8(want[0], dtype))
    assert got[1].shape == want[1].shape and np.array_equal(bits(got[1]), bits(want[1]))
    assert got[2].shape == want[2].shape and np.array_equal(as_i8(got[2], dtype), as_i8(want[2], dtype))
    if is_matmul_nbits_compatible(back, "w0"):
        b, s, z = _prepare_for_matmul_nbits(*got, back)
        ob, os_, oz = O.matmul_nbits_layout(want[0], want[1], want[2], gs, 4)
        assert np.array_equal(b, ob) and np.array_equal(bits(s), bits(os_)) and np.array_equal(z, oz)


def test_gptq_plugin_reads_calibration_input_from_node_meta(cuda, rng):
    k, n = 128, 48
    w = (rng.standard_normal((k, n)) * 0.05).astype(np.float32)
    x = rng.standard_normal((16, 12, k)).astype(np.float32)
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int4", group_size=64, symmetric=True,
                                          algorithm=q.GPTQConfig(block_size=64)))
    back = q.QConfig(**cfg.model_dump())
    assert isinstance(back.weights.algorithm, q.GPTQConfig)
    out = _Value("y", node=_Node({"input": x}))
    got = back.weights.algorithm.quantize_weights(_Value("w1", w), back, out=out)
    want = O.gptq_quantize(w, x, "int4", "group", 64, True, block_size=64, zp_dtype=O.np_dtype("int4"))
    assert np.array_equal(as_i8(got[0], "int4"), as_i8(want[0], "int4"))
    assert np.array_equal(bits(got[1]), bits(want[1])) and np.array_equal(as_i8(got[2], "int4"), as_i8(want[2], "int4"))
    with pytest.raises(AssertionError, match="Output value is required"):
        back.weights.algorithm.quantize_weights(_Value("w1", w), back)
    with pytest.raises(AssertionError, match="calibration data"):
        back.weights.algorithm.quantize_weights(_Value("w1", w), back, out=_Value("y", node=_Node({})))


SHARDED_SPECS = [("uint4", "group", 128, False), ("int8", "channel", -1, True), ("int4", "tensor", -1, True),
                 ("int4", "group", 64, True), ("uint8", "tensor", -1, False)]


@pytest.mark.parametrize("dtype,strategy,gs,sym", SHARDED_SPECS)
def test_sharded_prepass_publishes_results_for_the_plugins(cuda, rng, dtype, strategy, gs, sym):
    """Single-rank run of the multi-GPU pre-pass: every weight is quantized by the bulk pipeline and
    stored; the plugin then returns the stored triple, which must be EXACTLY what `_rtn_quantize`
    returns for the same request — dtype, shape and values, no reinterpretation on the way."""
    from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
    weights = {f"layer{i}.w": (rng.standard_normal(s) * 0.02).astype(np.float32)
               for i, s in enumerate([(256, 64), (128, 128), (384, 32)])}
    qt = q.QuantType.from_string(dtype)
    spec = RtnSpec(qt, strategy, gs, sym, False, 0.9, False, "kn")
    cfg = q.QConfig(weights=q.QWeightArgs(dtype=dtype, symmetric=sym, clip_ratio=0.9, strategy=strategy,
                                          group_size=gs if strategy == "group" else (-1 if strategy == "channel" else None)))
    with prequantized.scope():
        merged = quantize_weights_sharded(weights, spec)
        assert sorted(merged) == sorted(weights)
        for name, w in weights.items():
            calls = []
            value = _Value(name, w)
            triple = cfg.weights.algorithm.quantize_weights(value, cfg)
            assert triple is merged[name]                              # served from the store
            direct = _rtn_quantize(w, qt, cfg.weights.strategy, gs, sym, False, 0.9, False, np.dtype(np.float32),
                                   cfg.weights.zp_dtype)
            want = O.rtn_quantize(w, dtype, strategy, gs, sym, False, 0.9, False)
            for got, ref, orc in zip(triple, direct, want):
                assert got.dtype == ref.dtype == orc.dtype and got.shape == ref.shape == orc.shape
                assert np.array_equal(np.asarray(got).astype(np.float32), np.asarray(orc).astype(np.float32))
            assert np.array_equal(bits(triple[1]), bits(want[1]))
            if is_matmul_nbits_compatible(cfg, name):              # straight into the reference's packer
                b, s, z = _prepare_for_matmul_nbits(*triple, cfg)
                ob, os_, oz = O.matmul_nbits_layout(want[0], want[1], want[2], gs, 4)
                assert b.dtype == ob.dtype and np.array_equal(b, ob)
                assert np.array_equal(bits(s), bits(os_)) and np.array_equal(z, oz)
    assert prequantized.lookup(_Value("layer0.w", weights["layer0.w"]), cfg.weights, "rtn") is None   # scope ended


def test_prepass_entries_are_only_served_for_the_request_they_were_computed_for(cuda, rng):
    """Same initializer name, different array (what AWQ / SmoothQuant leave behind) or different
    weight arguments → the plugin recomputes; the GPTQ plugin never reads RTN entries."""
    w = (rng.standard_normal((256, 64)) * 0.02).astype(np.float32)
    spec = RtnSpec(q.QuantType.QUInt4, "group", 128, False, False, 1.0, False, "kn")
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="uint4", group_size=128))
    with prequantized.scope():
        merged = quantize_weights_sharded({"w": w}, spec)
        assert cfg.weights.algorithm.quantize_weights(_Value("w", w), cfg) is merged["w"]
        w2 = w * np.linspace(0.5, 2.0, 256, dtype=np.float32)[:, None]           # rescaled rows, same name
        got = cfg.weights.algorithm.quantize_weights(_Value("w", w2), cfg)
        assert got is not merged["w"]
        want = O.rtn_quantize(w2, "uint4", "group", 128)
        assert np.array_equal(as_i8(got[0], "uint4"), as_i8(want[0], "uint4"))
        for other in (q.QWeightArgs(dtype="uint4", group_size=64), q.QWeightArgs(dtype="uint4", group_size=128, mse=True),
                      q.QWeightArgs(dtype="int4", group_size=128, symmetric=True),
                      q.QWeightArgs(dtype="uint4", group_size=128, clip_ratio=0.9)):
            assert prequantized.lookup(_Value("w", w), other, "rtn") is None
        assert prequantized.lookup(_Value("w", w), cfg.weights, "gptq") is None
        x = rng.standard_normal((8, 16, 256)).astype(np.float32)
        gcfg = q.QConfig(weights=q.QWeightArgs(dtype="uint4", group_size=128, algorithm=q.GPTQConfig()))
        g = gcfg.weights.algorithm.quantize_weights(_Value("w", w), gcfg, out=_Value("y", node=_Node({"input": x})))
        assert g is not merged["w"]
    with pytest.raises(ValueError, match="layout='kn'"):
        quantize_weights_sharded({"w": w}, RtnSpec(q.QuantType.QUInt4, "group", 128, layout="matmul_nbits"))


def test_array_level_functions_accept_the_reference_packages_own_enums(cuda, rng):
    """Patched into the reference (`integration.patched_reference`), `_rtn_quantize` /
    `_gptq_quantize` receive the REFERENCE's `QuantType` / `QuantizationStrategy` members — other
    classes with the same member names (the reference cannot travel to the GPU box, so stand-ins
    with its definitions are used: core/_dtypes.py:33-41, core/_qconfig.py:31-36)."""
    import enum

    from onnx_quantize_b200.core._algorithms.gptq import _gptq_quantize
    from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize

    class QuantizationStrategy(str, enum.Enum):          # noqa: N801 - the reference's class name
        TENSOR = "tensor"
        CHANNEL = "channel"
        GROUP = "group"

    class QuantType(enum.Enum):
        QInt4 = 22
        QUInt4 = 21
        QInt8 = 3
        QUInt8 = 2

    w = (rng.standard_normal((128, 48)) * 0.05).astype(np.float32)
    got = _rtn_quantize(w, QuantType.QUInt4, QuantizationStrategy.GROUP, 64, False, False, 1.0, False,
                        np.dtype(np.float32), np.dtype(q.QuantType.QUInt4.np_dtype))
    want = O.rtn_quantize(w, "uint4", "group", 64)
    assert np.array_equal(as_i8(got[0], "uint4"), as_i8(want[0], "uint4")) and np.array_equal(bits(got[1]), bits(want[1]))
    x = rng.standard_normal((8, 16, 128)).astype(np.float32)
    got = _gptq_quantize(w, x, quant_type=QuantType.QInt8, strategy=QuantizationStrategy.CHANNEL, group_size=-1,
                         is_symmetric=True, zp_dtype=np.dtype(np.int8))
    want = O.gptq_quantize(w, x, "int8", "channel", -1, True)
    assert np.array_equal(as_i8(got[0], "int8"), as_i8(want[0], "int8")) and np.array_equal(bits(got[1]), bits(want[1]))
    with pytest.raises(AssertionError):
        _rtn_quantize(w, QuantType.QUInt4, "group", 64, False, False, 1.0, False, np.dtype(np.float32), None)


def test_gptq_plugin_shares_hessian_and_factor_between_nodes_with_the_same_input(cuda, rng):
    """q/k/v read one activation: the reference stores the SAME array object in their
    node.meta["input"] (calibrate.py:296-307).  The plugin must upload it, contract it and factorize
    once — and return what three independent computations return."""
    from onnx_quantize_b200 import _lib
    from onnx_quantize_b200.core._algorithms.gptq import calibration_cache

    lib = _lib.load()
    k = 256
    x = rng.standard_normal((16, 24, k)).astype(np.float32)
    ws = [(rng.standard_normal((k, n)) * 0.05).astype(np.float32) for n in (64, 32, 32)]
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int4", group_size=128, symmetric=True,
                                          algorithm=q.GPTQConfig(mode="propagate")))
    calibration_cache.clear()
    h0, m0, fh0, fm0 = (calibration_cache.hits, calibration_cache.misses, calibration_cache.factor_hits,
                        calibration_cache.factor_misses)
    launches, shared = [], []
    for i, w in enumerate(ws):
        before = lib.b200q_launch_count()
        shared.append(cfg.weights.algorithm.quantize_weights(_Value(f"w{i}", w), cfg,
                                                             out=_Value("y", node=_Node({"input": x}))))
        launches.append(lib.b200q_launch_count() - before)
    assert (calibration_cache.misses - m0, calibration_cache.hits - h0) == (1, 2)
    assert (calibration_cache.factor_misses - fm0, calibration_cache.factor_hits - fh0) == (1, 2)
    assert launches[1] < launches[0] and launches[2] < launches[0]      # no Hessian / factor launches again
    for w, got in zip(ws, shared):                                       # same results as without sharing
        calibration_cache.clear()
        alone = cfg.weights.algorithm.quantize_weights(_Value("w", w), cfg,
                                                       out=_Value("y", node=_Node({"input": x.copy()})))
        assert np.array_equal(as_i8(got[0], "int4"), as_i8(alone[0], "int4"))
        assert np.array_equal(bits(got[1]), bits(alone[1]))
    # a different array object with other contents, or an in-place edit, is a different input
    calibration_cache.clear()
    m1 = calibration_cache.misses
    cfg.weights.algorithm.quantize_weights(_Value("w", ws[0]), cfg, out=_Value("y", node=_Node({"input": x})))
    x *= np.float32(1.5)
    cfg.weights.algorithm.quantize_weights(_Value("w", ws[0]), cfg, out=_Value("y", node=_Node({"input": x})))
    assert calibration_cache.misses - m1 == 2
    calibration_cache.clear()


def test_streamed_hessian_from_host_chunks_matches_one_shot(cuda, rng, monkeypatch):
    """Host activations larger than one staging chunk are folded chunk by chunk (alpha = 2/n, beta = 1):
    same H as one call on the whole array, to float32 accumulation order."""
    import torch

    from onnx_quantize_b200.core._algorithms import gptq as gq
    from onnx_quantize_b200.hessian import hessian_accumulate
    k = 256
    x = rng.standard_normal((12, 700, k)).astype(np.float32)
    monkeypatch.setattr(gq, "_HESSIAN_CHUNK_BYTES", 1024 * 4 * k)       # 1024-row chunks -> 9 chunks, ragged tail
    h = gq._hessian_from_host(x, k, "bf16x3")
    want = torch.zeros((k, k), device=cuda)
    hessian_accumulate(torch.from_numpy(x).to(cuda), want, alpha=2.0 / 12, beta=0.0, precision="bf16x3")
    assert ((h - want).abs().max() / want.abs().max()).item() < 1e-5
    ref, _ = O.accumulate_hessian(x, np.zeros((k, k), np.float32), 0)
    assert np.abs(h.cpu().numpy() - ref).max() / np.abs(ref).max() < 2e-5
