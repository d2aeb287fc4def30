"""Host logic of the drop-in seam that needs no GPU: `integration.patched_reference` rebinding the
LIVE reference's numeric entry points (only when /root/reference is present — the GPU boxes do not
have it), and the keyed store of pre-pass results."""
import sys

import numpy as np
import pytest

import onnx_quantize_b200 as q
from onnx_quantize_b200.parallel import prequantized as P
from oracle import ref_shim


class _Const:
    def __init__(self, a):
        self._a = a

    def numpy(self):
        return self._a


class _Value:
    def __init__(self, name, a):
        self.name, self.const_value = name, _Const(a)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_patched_reference_rebinds_every_holder_and_restores():
    import importlib
    import os

    from onnx_quantize_b200 import integration
    from onnx_quantize_b200.core._algorithms import gptq as my_gptq
    from onnx_quantize_b200.core._algorithms import rtn as my_rtn
    from onnx_quantize_b200.core._calibration import minmax as my_minmax

    r = ref_shim.load()
    ref_shim._namespace("onnx_quantize.pre_passes", os.path.join(ref_shim._SRC, "pre_passes"))
    awq = importlib.import_module("onnx_quantize.pre_passes.awq")      # binds _rtn_quantize at import (awq.py:10)
    ref_pkg = sys.modules["onnx_quantize"]
    orig_rtn, orig_gptq, orig_cal = r.rtn._rtn_quantize, r.gptq._gptq_quantize, r.minmax.MinMaxCalibrator
    assert awq._rtn_quantize is orig_rtn
    with integration.patched_reference(ref_pkg):
        assert r.rtn._rtn_quantize is my_rtn._rtn_quantize
        assert awq._rtn_quantize is my_rtn._rtn_quantize               # the early binder too
        assert r.gptq._gptq_quantize is my_gptq._gptq_quantize
        assert r.minmax.MinMaxCalibrator is my_minmax.MinMaxCalibrator
        assert all(v is not orig_cal for v in r.calib_factory._CALIBRATORS.values())
        assert any(v is my_minmax.MinMaxCalibrator for v in r.calib_factory._CALIBRATORS.values())
        # the reference's own plugin now lands in this package's code, which refuses to run
        # without a CUDA device instead of falling back to NumPy
        import torch
        if not torch.cuda.is_available():
            cfg = r.QConfig(weights=dict(dtype="uint4", group_size=32))
            w = np.ones((64, 8), np.float32)
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                cfg.weights.algorithm.quantize_weights(_Value("w", w), cfg)
    assert r.rtn._rtn_quantize is orig_rtn and awq._rtn_quantize is orig_rtn
    assert r.gptq._gptq_quantize is orig_gptq and r.minmax.MinMaxCalibrator is orig_cal
    assert any(v is orig_cal for v in r.calib_factory._CALIBRATORS.values())


def test_prequantized_store_is_keyed_by_request_and_array():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((512, 96)).astype(np.float32)
    wa = q.QWeightArgs(dtype="uint4", group_size=128)
    triple = (object(), object(), object())
    P.clear()
    with P.scope():
        P.put("w", P.request_digest(wa, "rtn", w), triple)
        assert P.lookup(_Value("w", w), wa, "rtn") is triple
        assert P.lookup(_Value("w", w.copy()), wa, "rtn") is triple            # same content, other object
        assert P.lookup(_Value("other", w), wa, "rtn") is None
        assert P.lookup(_Value("w", w), wa, "gptq") is None
        assert P.lookup(_Value("w", w), q.QWeightArgs(dtype="uint4", group_size=64), "rtn") is None
        assert P.lookup(_Value("w", w), q.QWeightArgs(dtype="uint4", group_size=128, mse=True), "rtn") is None
        assert P.lookup(_Value("w", w), q.QWeightArgs(dtype="int4", group_size=128), "rtn") is None
        scaled = w * np.float32(1.0001)
        assert P.lookup(_Value("w", scaled), wa, "rtn") is None               # AWQ / SmoothQuant under the same name
        rows = w.copy()
        rows[7] *= 2
        assert P.weight_fingerprint(rows) != P.weight_fingerprint(w)
        assert P.weight_fingerprint(w.T) != P.weight_fingerprint(w)           # shape is part of it
    assert P.lookup(_Value("w", w), wa, "rtn") is None                       # dropped with the scope


def test_rtn_spec_and_weight_args_give_the_same_digest():
    from onnx_quantize_b200.pipeline import RtnSpec
    w = np.ones((256, 8), np.float32)
    for kw in (dict(dtype="uint4", group_size=128), dict(dtype="int8", group_size=-1, symmetric=True),
               dict(dtype="int4", symmetric=True, clip_ratio=0.9, mse=True), dict(dtype="uint8", reduce_range=True)):
        wa = q.QWeightArgs(**kw)
        spec = RtnSpec.from_weight_args(wa)
        assert P.request_digest(spec.as_weight_args(), "rtn", w) == P.request_digest(wa, "rtn", w)
