"""``quantize(model, qconfig)`` — the public graph-level entry point.

The ONNX-IR graph rewriting (pre-passes, qrules, qfunctions) is NOT re-implemented here: per the
project scope it stays the reference's own code.  This function hands the model to the reference
pipeline (``onnx_quantize.quantize``) after swapping the reference's numeric plugins for the GPU
ones of this package (``INTEGRATION.md`` shows the three-line patch a maintainer would apply
instead).  It needs the ONNX stack (onnx, onnx_ir, onnxscript) and the reference package to be
importable and raises ImportError otherwise — the array-level API of this package does not.
"""
from __future__ import annotations

__all__ = ["quantize"]


def quantize(model, qconfig):
    """Quantize an ONNX model with the reference graph pipeline and this package's GPU numerics."""
    try:
        import onnx_quantize as _ref  # the reference package (graph side)
    except ImportError as e:
        raise ImportError(
            "quantize(model, qconfig) drives the reference's ONNX graph pipeline and needs "
            "`onnx_quantize` with onnx / onnx_ir / onnxscript installed; the array-level API "
            "(_rtn_quantize, _gptq_quantize, MinMaxCalibrator, ...) works without them.") from e
    from onnx_quantize_b200 import integration

    with integration.patched_reference(_ref):
        return _ref.quantize(model, qconfig)
