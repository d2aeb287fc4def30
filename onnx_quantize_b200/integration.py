"""Swapping the reference's numeric plugins for the GPU ones (used by ``quantize()`` when the
reference package and the ONNX stack are importable)."""
from __future__ import annotations

import contextlib


@contextlib.contextmanager
def patched_reference(ref):
    """Within the context the reference's ``_rtn_quantize`` / ``_gptq_quantize`` / ``_hqq_quantize`` /
    ``MinMaxCalibrator`` resolve to this package's implementations."""
    import importlib

    from onnx_quantize_b200.core._algorithms import gptq as my_gptq
    from onnx_quantize_b200.core._algorithms import hqq as my_hqq
    from onnx_quantize_b200.core._algorithms import rtn as my_rtn
    from onnx_quantize_b200.core._calibration import minmax as my_minmax

    r_rtn = importlib.import_module(ref.__name__ + ".core._algorithms.rtn")
    r_gptq = importlib.import_module(ref.__name__ + ".core._algorithms.gptq")
    r_hqq = importlib.import_module(ref.__name__ + ".core._algorithms.hqq")
    r_fact = importlib.import_module(ref.__name__ + ".core._calibration.factory")
    saved = (r_rtn._rtn_quantize, r_gptq._gptq_quantize, dict(r_fact._CALIBRATORS), r_hqq._hqq_quantize)
    r_rtn._rtn_quantize = my_rtn._rtn_quantize
    r_gptq._gptq_quantize = my_gptq._gptq_quantize
    r_hqq._hqq_quantize = my_hqq._hqq_quantize
    for key in list(r_fact._CALIBRATORS):
        if getattr(key, "value", key) == "minmax":
            r_fact._CALIBRATORS[key] = my_minmax.MinMaxCalibrator
    try:
        yield
    finally:
        r_rtn._rtn_quantize, r_gptq._gptq_quantize, r_hqq._hqq_quantize = saved[0], saved[1], saved[3]
        r_fact._CALIBRATORS.clear()
        r_fact._CALIBRATORS.update(saved[2])
