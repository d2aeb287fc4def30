"""Swapping the reference's numeric plugins for the GPU ones (used by ``quantize()`` when the
reference package and the ONNX stack are importable)."""
from __future__ import annotations

import contextlib
import importlib
import sys


def _replacements(ref_name: str) -> list[tuple[object, object]]:
    """(reference object, this package's object) pairs — the array-level seams of SURVEY.md §8b."""
    from onnx_quantize_b200.core._algorithms import gptq as my_gptq
    from onnx_quantize_b200.core._algorithms import hqq as my_hqq
    from onnx_quantize_b200.core._algorithms import rtn as my_rtn
    from onnx_quantize_b200.core._calibration import minmax as my_minmax

    r_rtn = importlib.import_module(ref_name + ".core._algorithms.rtn")
    r_gptq = importlib.import_module(ref_name + ".core._algorithms.gptq")
    r_hqq = importlib.import_module(ref_name + ".core._algorithms.hqq")
    r_minmax = importlib.import_module(ref_name + ".core._calibration.minmax")
    return [(r_rtn._rtn_quantize, my_rtn._rtn_quantize),
            (r_rtn._quantize_bias, my_rtn._quantize_bias),
            (r_gptq._gptq_quantize, my_gptq._gptq_quantize),
            (r_hqq._hqq_quantize, my_hqq._hqq_quantize),
            (r_minmax.MinMaxCalibrator, my_minmax.MinMaxCalibrator)]


@contextlib.contextmanager
def patched_reference(ref):
    """Within the context the reference's ``_rtn_quantize`` / ``_quantize_bias`` / ``_gptq_quantize``
    / ``_hqq_quantize`` / ``MinMaxCalibrator`` resolve to this package's implementations — in EVERY
    loaded module of the reference that holds a binding to them, not only the defining one: the
    reference binds ``_rtn_quantize`` at import time in ``pre_passes/awq.py`` (:10, used :156, :226)
    and ``_quantize_bias`` in ``qrules/_qlinear/gemm_to_qgemm.py`` (:3), and registers the calibrator
    class in the ``_CALIBRATORS`` table of ``core/_calibration/factory.py``.  Modules of the
    reference imported later pick the patched objects up from the defining modules.  Results
    published by a multi-GPU pre-pass (``parallel.prequantized``) and the GPTQ plugin's device-side
    calibration cache are dropped at exit."""
    from onnx_quantize_b200.parallel import prequantized

    prefix = ref.__name__ + "."
    for sub in (".pre_passes.awq", ".qrules._qlinear.gemm_to_qgemm", ".core._calibration.factory"):
        try:                                        # make sure the early binders are loaded and seen
            importlib.import_module(ref.__name__ + sub)
        except ImportError:
            pass
    pairs = _replacements(ref.__name__)
    undo: list[tuple[object, object, object]] = []   # (container, key, original)
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == ref.__name__ or name.startswith(prefix)):
            continue
        for attr, value in list(vars(mod).items()):
            for old, new in pairs:
                if value is old:
                    undo.append((mod, attr, old))
                    setattr(mod, attr, new)
            if isinstance(value, dict) and attr.isupper():     # registries such as _CALIBRATORS
                for key, item in list(value.items()):
                    for old, new in pairs:
                        if item is old:
                            undo.append((value, key, old))
                            value[key] = new
    try:
        with prequantized.scope():
            yield
    finally:
        from onnx_quantize_b200.core._algorithms.gptq import calibration_cache

        calibration_cache.clear()                  # device Hessians / factors of the last calibration arrays
        for container, key, old in reversed(undo):
            if isinstance(container, dict):
                container[key] = old
            else:
                setattr(container, key, old)
