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
