"""Which calibrator class serves which ``CalibrationMethod`` (reference: ``core/_calibration/
factory.py`` :10-32 — same ``get_calibrator(method, **kwargs)`` entry point, same exceptions:
``KeyError`` for an unknown method, ``TypeError`` naming the class for bad constructor arguments).
Here the table is filled through ``register_calibrator`` so that other device calibrators can be
added next to the min/max one."""
from __future__ import annotations

from typing import Any, Callable

from onnx_quantize_b200.core._calibration import base as _base

__all__ = ["get_calibrator", "register_calibrator"]

_CALIBRATORS: dict[_base.CalibrationMethod, type[_base.Calibrator]] = {}


def register_calibrator(method: _base.CalibrationMethod) -> Callable[[type], type]:
    def bind(cls: type) -> type:
        _CALIBRATORS[method] = cls
        return cls
    return bind


def get_calibrator(method: _base.CalibrationMethod = _base.CalibrationMethod.MINMAX, **kwargs: Any) -> _base.Calibrator:
    chosen = _CALIBRATORS[method]
    try:
        instance = chosen(**kwargs)
    except TypeError as bad_args:
        raise TypeError(f"Invalid arguments for {chosen.__name__}: {bad_args}") from bad_args
    return instance


def _register_builtin() -> None:
    from onnx_quantize_b200.core._calibration.minmax import MinMaxCalibrator

    register_calibrator(_base.CalibrationMethod.MINMAX)(MinMaxCalibrator)


_register_builtin()
