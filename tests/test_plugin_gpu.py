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


def test_sharded_prepass_publishes_results_for_the_plugins(cuda, rng):
    """Single-rank run of the multi-GPU pre-pass: every weight is quantized by the bulk pipeline,
    stored by initializer name, and the plugin returns the stored triple without recomputing."""
    weights = {f"layer{i}.w": (rng.standard_normal(s) * 0.02).astype(np.float32)
               for i, s in enumerate([(256, 64), (128, 128), (384, 32)])}
    spec = RtnSpec(q.QuantType.QUInt4, "group", 128, False, False, 1.0, False, "kn")
    prequantized.clear()
    try:
        merged = quantize_weights_sharded(weights, spec)
        assert sorted(merged) == sorted(weights)
        cfg = q.QConfig(weights=q.QWeightArgs(dtype="uint4", group_size=128))
        for name, w in weights.items():
            triple = cfg.weights.algorithm.quantize_weights(_Value(name, None), cfg)   # no array needed
            want = O.rtn_quantize(w, "uint4", "group", 128)
            assert np.array_equal(np.asarray(triple[0]).view(np.uint8), as_i8(want[0], "uint4"))
            assert np.array_equal(bits(np.asarray(triple[1]).reshape(-1)), bits(want[1].reshape(-1)))
    finally:
        prequantized.clear()
