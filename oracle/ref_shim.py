"""Import shim for the *live* reference numeric core (TEST INFRASTRUCTURE ONLY).

The reference package (AyoubMDL/onnx_quantize, mounted read-only at /root/reference in the
build container) cannot be imported as a whole here because onnx / onnx_ir / onnxscript /
onnxruntime are not installed.  Its numeric core only touches ``onnx_ir.DataType`` though, so
this module

  1. installs a tiny stand-in ``onnx_ir`` module (``DataType`` with ``.numpy()`` / ``.bitwidth``
     plus the few names that are evaluated in annotations), and
  2. registers namespace packages for ``onnx_quantize`` and its sub-packages whose ``__path__``
     points into the reference tree, so the package ``__init__`` files (which import onnxscript)
     never run.

Nothing from the reference is copied; the reference source files are executed where they lie.
It exists so that ``oracle/gen_golden.py`` can produce the fixtures under ``tests/golden/`` and
so that the CPU test-suite can cross-check ``oracle/np_oracle.py`` against the real thing when
``/root/reference`` is present.  It is never imported by the product package, by ``-m gpu``
tests, by ``smoke()`` or by ``bench.py`` (the GPU boxes have no /root/reference).
"""
from __future__ import annotations

import enum
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200Q_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "src", "onnx_quantize")


def available() -> bool:
    return os.path.isdir(_SRC)


def _install_onnx_ir_stub() -> None:
    if "onnx_ir" in sys.modules:
        return
    import ml_dtypes
    import numpy as np

    class DataType(enum.IntEnum):
        # numeric values follow onnx.TensorProto.DataType
        UINT8 = 2
        INT8 = 3
        INT32 = 6
        UINT32 = 12
        UINT4 = 21
        INT4 = 22

        def numpy(self):
            return {
                DataType.UINT8: np.dtype(np.uint8),
                DataType.INT8: np.dtype(np.int8),
                DataType.INT32: np.dtype(np.int32),
                DataType.UINT32: np.dtype(np.uint32),
                DataType.UINT4: np.dtype(ml_dtypes.uint4),
                DataType.INT4: np.dtype(ml_dtypes.int4),
            }[self]

        @property
        def bitwidth(self):
            return 4 if self in (DataType.UINT4, DataType.INT4) else (
                8 if self in (DataType.UINT8, DataType.INT8) else 32)

    ir = types.ModuleType("onnx_ir")
    ir.DataType = DataType

    class _Value:  # only used in annotations / tiny helpers of qrules/_common.py
        def __init__(self, name=None, shape=None, const_value=None):
            self.name, self.shape, self.const_value = name, shape, const_value

    class _Tensor:
        def __init__(self, array):
            self._a = array

        def numpy(self):
            return self._a

    ir.Value = _Value
    ir.val = lambda name, shape=None, const_value=None, **kw: _Value(name, shape, const_value)
    ir.tensor = lambda a, **kw: _Tensor(a)
    ir.tape = types.ModuleType("onnx_ir.tape")
    ir.tape.Tape = object
    ir.passes = types.ModuleType("onnx_ir.passes")
    ir.passes.InPlacePass = object
    ir.passes.PassResult = lambda model, modified=False: (model, modified)
    ir.Node = object
    ir.Model = object
    ir.convenience = types.ModuleType("onnx_ir.convenience")
    ir.convenience.get_const_tensor = lambda v: getattr(v, "const_value", None)
    # the node keeps its Value object: carrying the new tensor over is what "replace all uses" means here
    ir.convenience.replace_all_uses_with = lambda old, new: setattr(old, "const_value", new.const_value)
    ir.AttrInt64 = lambda name, value: types.SimpleNamespace(as_int=lambda: value)
    sys.modules["onnx_ir.convenience"] = ir.convenience
    sys.modules["onnx_ir"] = ir
    sys.modules["onnx_ir.tape"] = ir.tape
    sys.modules["onnx_ir.passes"] = ir.passes


def _namespace(name: str, path: str) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod


_loaded = None


def load():
    """Return a namespace object exposing the reference's hot-path functions."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_onnx_ir_stub()
    _namespace("onnx_quantize", _SRC)
    _namespace("onnx_quantize.core", os.path.join(_SRC, "core"))
    _namespace("onnx_quantize.core._algorithms", os.path.join(_SRC, "core", "_algorithms"))
    _namespace("onnx_quantize.core._calibration", os.path.join(_SRC, "core", "_calibration"))
    _namespace("onnx_quantize.qrules", os.path.join(_SRC, "qrules"))

    ns = types.SimpleNamespace()
    ns.dtypes = importlib.import_module("onnx_quantize.core._dtypes")
    ns.qconfig = importlib.import_module("onnx_quantize.core._qconfig")
    ns.utils = importlib.import_module("onnx_quantize.core._algorithms.utils")
    ns.rtn = importlib.import_module("onnx_quantize.core._algorithms.rtn")
    ns.gptq = importlib.import_module("onnx_quantize.core._algorithms.gptq")
    ns.hqq = importlib.import_module("onnx_quantize.core._algorithms.hqq")
    ns.pack = importlib.import_module("onnx_quantize.core._pack")
    ns.minmax = importlib.import_module("onnx_quantize.core._calibration.minmax")
    ns.calib_base = importlib.import_module("onnx_quantize.core._calibration.base")
    ns.calib_factory = importlib.import_module("onnx_quantize.core._calibration.factory")
    ns.common = importlib.import_module("onnx_quantize.qrules._common")
    ns.QuantType = ns.dtypes.QuantType
    ns.QuantizationStrategy = ns.qconfig.QuantizationStrategy
    ns.QConfig = ns.qconfig.QConfig
    ns.QWeightArgs = ns.qconfig.QWeightArgs
    _loaded = ns
    return ns


def run_reference_awq(w, x, qconfig_kwargs: dict, clip_search: bool = False):
    """Run the reference's ``AwqPass._apply_awq`` (and ``_apply_awq_clip``) — unmodified, executed
    where it lies — on a stand-in MatMul node holding weight ``w`` (K,N) and calibration input
    ``x``.  Returns ``(best_scale (K,), best_clip_ratio or None)``."""
    import numpy as np

    r = load()
    _namespace("onnx_quantize.pre_passes", os.path.join(_SRC, "pre_passes"))
    awq = importlib.import_module("onnx_quantize.pre_passes.awq")
    ir = sys.modules["onnx_ir"]
    qconfig = r.QConfig(**qconfig_kwargs)
    w_val = ir.val("w", const_value=ir.tensor(np.array(w, copy=True)))
    node = types.SimpleNamespace(
        op_type="MatMul", domain="", inputs=[ir.val("x"), w_val], attributes={},
        outputs=[types.SimpleNamespace(name="y")],
        meta={"qconfig": qconfig.model_dump(), "input": np.array(x, copy=True)})
    model = types.SimpleNamespace(graph=types.SimpleNamespace(initializers={}))
    p = awq.AwqPass(clip_search=clip_search, target_op_types=("MatMul", "Gemm"))
    captured = {}
    p._insert_mul_node_before = lambda n, m, scale_init: captured.__setitem__("inv", scale_init.const_value.numpy())
    assert p._apply_awq(node, model)
    best_scale = 1.0 / captured["inv"]
    best_clip = None
    if clip_search:
        # runs on the node as _apply_awq left it: weights scaled by best_scale, inputs divided by it
        assert p._apply_awq_clip(node)
        best_clip = r.QConfig(**node.meta["qconfig"]).weights.clip_ratio
    return best_scale, best_clip


def run_reference_smooth_quant(w, x, alpha: float = 0.5):
    """The reference's ``SmoothQuantPass._smooth_quant_node`` on a stand-in MatMul node →
    ``(scale (K,), updated weights)``."""
    import numpy as np

    r = load()
    _namespace("onnx_quantize.pre_passes", os.path.join(_SRC, "pre_passes"))
    sq = importlib.import_module("onnx_quantize.pre_passes.smooth_quant")
    ir = sys.modules["onnx_ir"]
    qconfig = r.QConfig(weights=dict(dtype="int8"), preprocessors=[dict(preprocessing_type="smooth_quant", alpha=alpha)])
    w_val = ir.val("w", const_value=ir.tensor(np.array(w, copy=True)))
    node = types.SimpleNamespace(op_type="MatMul", domain="", inputs=[ir.val("x"), w_val], attributes={},
                                 outputs=[types.SimpleNamespace(name="y")],
                                 meta={"qconfig": qconfig.model_dump(), "input": np.array(x, copy=True)})
    model = types.SimpleNamespace(graph=types.SimpleNamespace(initializers={}))
    p = sq.SmoothQuantPass(alpha=alpha, target_op_types=("MatMul", "Gemm"))
    captured = {}
    p._insert_mul_node_before = lambda n, m, scale_init: captured.__setitem__("inv", scale_init.const_value.numpy())
    assert p._smooth_quant_node(node, model)
    return 1.0 / captured["inv"], model.graph.initializers["w"].const_value.numpy()
