"""Quantized element types and their integer ranges.

Mirrors the public surface of the reference's ``core/_dtypes.py`` (``QuantType`` members,
``from_string``, ``np_dtype``, ``bitwidth``, ``qrange``; reference core/_dtypes.py:33-70).  The
enum values are ``onnx_ir.DataType`` members when onnx_ir is installed, otherwise a stand-in with
the same names, numeric codes (onnx TensorProto) and ``.numpy()`` / ``.bitwidth`` accessors, so
the package imports on a box without the ONNX stack.
"""
from __future__ import annotations

__all__ = ["QuantType"]

import enum

import ml_dtypes
import numpy as np

try:  # pragma: no cover - the ONNX stack is optional
    from onnx_ir import DataType as _DataType
except Exception:  # noqa: BLE001 - any import problem means "not available"

    class _DataType(enum.IntEnum):
        UINT8 = 2
        INT8 = 3
        INT32 = 6
        UINT32 = 12
        UINT4 = 21
        INT4 = 22

        def numpy(self) -> np.dtype:
            return np.dtype(_NUMPY_OF[self.name])

        @property
        def bitwidth(self) -> int:
            return _BITS_OF[self.name]


_NUMPY_OF = {"UINT8": np.uint8, "INT8": np.int8, "INT32": np.int32, "UINT32": np.uint32,
             "UINT4": ml_dtypes.uint4, "INT4": ml_dtypes.int4}
_BITS_OF = {"UINT8": 8, "INT8": 8, "INT32": 32, "UINT32": 32, "UINT4": 4, "INT4": 4}

# (full range, symmetric range or None, reduced range) per type name — reference
# core/_dtypes.py:8-31.  Signed symmetric ranges drop the most negative code; the "reduced"
# ranges are the reference's (note int8 → [-64, 64] and int32 → [-2^30, 2^30]).
_RANGES = {
    "UINT4": ((0, 15), None, (0, 7)),
    "INT4": ((-8, 7), (-7, 7), (-4, 3)),
    "UINT8": ((0, 255), None, (0, 127)),
    "INT8": ((-128, 127), (-127, 127), (-64, 64)),
    "UINT32": ((0, 2**32 - 1), None, (0, 2**31 - 1)),
    "INT32": ((-(2**31), 2**31 - 1), (-(2**31 - 1), 2**31 - 1), (-(2**30), 2**30)),
}


class QuantType(enum.Enum):
    """Enumeration of quantization types."""

    QInt4 = _DataType.INT4
    QUInt4 = _DataType.UINT4
    QInt8 = _DataType.INT8
    QUInt8 = _DataType.UINT8
    QInt32 = _DataType.INT32
    QUInt32 = _DataType.UINT32

    @classmethod
    def from_string(cls, value: str) -> "QuantType":
        names = {m.short_name: m for m in cls}
        try:
            return names[value.lower().strip()]
        except KeyError:
            raise ValueError(
                f"Invalid quantization type '{value}'. Expected one of: {', '.join(_ORDER)}"
            ) from None

    @classmethod
    def coerce(cls, value) -> "QuantType":
        """This enum's member for ``value``: a member of this enum, a type name, or the member of the
        same NAME of another package's ``QuantType`` — the reference's own enum is what arrives when
        these functions are patched into the reference (``integration.patched_reference``)."""
        if isinstance(value, cls):
            return value
        if isinstance(value, str):
            return cls.from_string(value)
        if isinstance(value, enum.Enum) and value.name in cls.__members__:
            return cls[value.name]
        raise TypeError(f"not a quantization type: {value!r}")

    @property
    def short_name(self) -> str:
        """'int4', 'uint4', 'int8', … — also the key the C ABI wrapper uses."""
        return self.value.name.lower()

    @property
    def np_dtype(self) -> np.dtype:
        return self.value.numpy()

    @property
    def bitwidth(self) -> int:
        return self.value.bitwidth

    @property
    def is_signed(self) -> bool:
        return self.value.name.startswith("INT")

    def qrange(self, is_symmetric: bool, reduce_range: bool = False) -> tuple[int, int]:
        """reduce_range wins; then the symmetric table (signed types only); else the full range."""
        full, sym, reduced = _RANGES[self.value.name]
        if reduce_range:
            return reduced
        if is_symmetric and sym is not None:
            return sym
        return full


_ORDER = ("int4", "uint4", "int8", "uint8", "int32", "uint32")
