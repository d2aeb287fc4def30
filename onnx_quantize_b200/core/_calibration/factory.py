"""Calibrator registry — mirrors the reference's ``core/_calibration/factory.py`` (:10-32)."""
from __future__ import annotations

__all__ = ["get_calibrator"]

from typing import Any

from onnx_quantize_b200.core._calibration.base import CalibrationMethod, Calibrator
from onnx_quantize_b200.core._calibration.minmax import MinMaxCalibrator

_CALIBRATORS: dict[CalibrationMethod, type[Calibrator]] = {
    CalibrationMethod.MINMAX: MinMaxCalibrator,
}


def get_calibrator(method: CalibrationMethod = CalibrationMethod.MINMAX, **kwargs: Any) -> Calibrator:
    """Instantiate the calibrator registered for ``method`` with ``kwargs`` (e.g. ``momentum``)."""
    cls = _CALIBRATORS[method]
    try:
        return cls(**kwargs)
    except TypeError as e:
        raise TypeError(f"Invalid arguments for {cls.__name__}: {e}") from e
