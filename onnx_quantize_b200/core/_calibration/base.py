"""Calibration configuration and the ``Calibrator`` plugin interface.

Same public surface as the reference's ``core/_calibration/base.py`` (``ExecutionProvider``
:12-32, ``CalibrationMethod`` :35-38, ``CalibrationParams`` :41-97, ``CalibrationData`` :100-110,
``Calibrator`` :113-144): field names, defaults, accepted aliases and error texts are kept so a
serialized ``QConfig`` round-trips between the two packages.
"""
from __future__ import annotations

__all__ = ["CalibrationMethod", "CalibrationParams"]

import abc
import dataclasses
import enum
from typing import Any

from pydantic import BaseModel, ConfigDict, field_validator


class ExecutionProvider(str, enum.Enum):
    CPU = "CPUExecutionProvider"
    CUDA = "CUDAExecutionProvider"

    @classmethod
    def from_alias(cls, value: str) -> "ExecutionProvider":
        short = {"cpu": cls.CPU, "cuda": cls.CUDA, "gpu": cls.CUDA}
        hit = short.get(value.lower())
        if hit is not None:
            return hit
        for member in cls:
            if member.value == value:
                return member
        valid = sorted(set(short) | {m.value for m in cls})
        raise ValueError(f"Invalid execution provider '{value}'. Valid values are: {valid}")


class CalibrationMethod(enum.Enum):
    """Calibration method enum."""

    MINMAX = "minmax"


class CalibrationParams(BaseModel):
    """How activations are calibrated: method, sample budget, batching, EMA momentum, ORT provider."""

    model_config = ConfigDict(extra="forbid")

    method: CalibrationMethod | str = CalibrationMethod.MINMAX
    num_samples: int = 100
    batch_size: int = 10
    momentum: float = 0.0
    provider: ExecutionProvider | str = ExecutionProvider.CPU

    @field_validator("method", mode="before")
    @classmethod
    def _coerce_method(cls, value: Any):
        if not isinstance(value, str):
            return value
        for member in CalibrationMethod:
            if member.value == value:
                return member
        raise ValueError(f"Invalid calibration method '{value}'. Valid methods are: "
                         f"{[m.value for m in CalibrationMethod]}")

    @field_validator("provider", mode="before")
    @classmethod
    def _coerce_provider(cls, value: Any):
        return ExecutionProvider.from_alias(value) if isinstance(value, str) else value

    @field_validator("momentum", mode="after")
    @classmethod
    def _check_momentum(cls, value: float) -> float:
        if value < 0 or value >= 1:
            raise ValueError(f"Momentum must be in [0, 1), got {value}")
        return value

    @field_validator("num_samples", "batch_size", mode="after")
    @classmethod
    def _check_positive(cls, value: int, info) -> int:
        if value <= 0:
            raise ValueError(f"{info.field_name} must be positive, got {value}")
        return value


@dataclasses.dataclass
class CalibrationData:
    """Statistics of one calibrated tensor (here: 0-d float32 arrays once read back)."""

    min_val: Any
    max_val: Any


class Calibrator(abc.ABC):
    """Plugin interface: ``collect(name, array)`` per batch, ``compute_range(name)`` at the end."""

    def __init__(self) -> None:
        self.data: dict[str, CalibrationData] = {}

    @abc.abstractmethod
    def collect(self, name: str, array) -> None:
        """Fold one activation batch of tensor ``name`` into the running statistics."""

    @abc.abstractmethod
    def compute_range(self, name: str):
        """→ ``(min, max)`` as 0-d float32 arrays, zero included."""
