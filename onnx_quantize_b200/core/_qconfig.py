"""Quantization configuration (public API) and the algorithm / pre-processing plugin registries.

This is the drop-in boundary of SURVEY.md §8b: the same pydantic models, field names, defaults,
coercions and validation messages as the reference's ``core/_qconfig.py`` (``QuantizationStrategy``
:31, ``QFormat`` :39, ``AlgorithmConfig`` :46, ``PreProcessingConfig`` :75,
``register_algorithm_config`` :104, ``_BaseArgs`` :170, ``QWeightArgs`` :271, ``QActivationArgs``
:304, ``QConfig`` :338), so ``QConfig(**qconfig.model_dump())`` round-trips and a config written
for one package constructs in the other.  The default algorithm is this package's
``RTNConfig`` whose ``quantize_weights`` runs on the GPU.
"""
from __future__ import annotations

import logging
from collections.abc import Sequence
from enum import Enum
from typing import TYPE_CHECKING, Any, ClassVar

import numpy as np
from pydantic import BaseModel, ConfigDict, Field, SerializeAsAny, field_validator, model_validator

from onnx_quantize_b200.core._calibration.base import CalibrationParams
from onnx_quantize_b200.core._dtypes import QuantType

if TYPE_CHECKING:  # pragma: no cover
    import onnx_ir as ir

logger = logging.getLogger(__name__)

_SUPPORTED_OP_TYPES = ("MatMul", "Gemm")
_FOUR_BIT = frozenset({QuantType.QInt4, QuantType.QUInt4})
_EIGHT_BIT = frozenset({QuantType.QInt8, QuantType.QUInt8})


class QuantizationStrategy(str, Enum):
    """Granularity of one (scale, zero-point) pair."""

    TENSOR = "tensor"
    CHANNEL = "channel"
    GROUP = "group"


def coerce_strategy(value) -> QuantizationStrategy:
    """A member of this package's enum for a member of ANY ``QuantizationStrategy`` enum (the
    reference's own class is what arrives when the array-level functions are patched into the
    reference).  Plain strings are refused, as the reference's ``isinstance`` assertions do."""
    if isinstance(value, QuantizationStrategy):
        return value
    assert isinstance(value, Enum) and type(value).__name__ == "QuantizationStrategy", \
        f"strategy must be a QuantizationStrategy, got {value!r}"
    return QuantizationStrategy(value.value)


class QFormat(str, Enum):
    """Graph representation of the quantized model."""

    QDQ = "qdq"
    QLINEAR = "qlinear"


# ------------------------------------------------------------------------------------------------
# plugin base classes + registries
# ------------------------------------------------------------------------------------------------
class AlgorithmConfig(BaseModel):
    """Base class of weight-quantization algorithm plugins.

    A plugin declares an ``algorithm_type: Literal[tag] = tag`` field, is decorated with
    :func:`register_algorithm_config`, sets ``requires_calibration`` when it needs input
    activations, implements :meth:`quantize_weights` and may override
    :meth:`validate_weight_args`.
    """

    requires_calibration: ClassVar[bool] = False

    def validate_weight_args(self, weight_args: "QWeightArgs") -> None:
        """Hook to validate/adjust the enclosing ``QWeightArgs``; the default accepts anything."""

    def quantize_weights(self, w: "ir.Value", qconfig: "QConfig", out: "ir.Value | None" = None
                         ) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Quantize weight ``w`` → ``(q_weight, scale, zero_point)``."""
        raise NotImplementedError(f"{type(self).__name__} must implement quantize_weights().")


class PreProcessingConfig(BaseModel):
    """Base class of pre-processing pass plugins (AWQ, SmoothQuant, … — graph side, not hot path)."""

    requires_calibration: ClassVar[bool] = True
    requires_post_calibration: ClassVar[bool] = True

    def build_pass(self, qconfig: "QConfig"):
        raise NotImplementedError(f"{type(self).__name__} must implement build_pass().")


_ALGORITHM_REGISTRY: dict[str, type[AlgorithmConfig]] = {}
_PREPROCESSING_REGISTRY: dict[str, type[PreProcessingConfig]] = {}


def _register(cls, tag_field: str, registry: dict, what: str):
    field = cls.model_fields.get(tag_field)
    if field is None:
        raise TypeError(f"{cls.__name__} must declare {what} field to be registered.")
    registry[field.default] = cls
    return cls


def register_algorithm_config(cls: type[AlgorithmConfig]) -> type[AlgorithmConfig]:
    """Class decorator: register under the default of the ``algorithm_type`` field."""
    return _register(cls, "algorithm_type", _ALGORITHM_REGISTRY, "an 'algorithm_type'")


def register_preprocessing_config(cls: type[PreProcessingConfig]) -> type[PreProcessingConfig]:
    """Class decorator: register under the default of the ``preprocessing_type`` field."""
    return _register(cls, "preprocessing_type", _PREPROCESSING_REGISTRY, "a 'preprocessing_type'")


def _default_algorithm_config() -> AlgorithmConfig:
    from onnx_quantize_b200.core._algorithms.rtn import RTNConfig

    return RTNConfig()


def _from_registry(value: Any, base: type, tag_field: str, registry: dict):
    """instance → itself; mapping → registered subclass chosen by its tag; anything else as is."""
    if isinstance(value, base) or not isinstance(value, dict):
        return value
    tag = value.get(tag_field)
    if tag not in registry:
        raise ValueError(f"Unknown {tag_field} {tag!r}. Registered: {sorted(registry)}")
    return registry[tag](**value)


def _resolve_algorithm_config(value: Any):
    if value is None:
        return _default_algorithm_config()
    return _from_registry(value, AlgorithmConfig, "algorithm_type", _ALGORITHM_REGISTRY)


def _resolve_preprocessing_config(value: Any):
    return _from_registry(value, PreProcessingConfig, "preprocessing_type", _PREPROCESSING_REGISTRY)


# ------------------------------------------------------------------------------------------------
# per-tensor-kind arguments
# ------------------------------------------------------------------------------------------------
_BAD_GROUP = ("Invalid group size {0}. Use group_size > 0 for strategy='group' and "
              "group_size = -1 for '{1}'")


class _BaseArgs(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True)

    dtype: QuantType | str = QuantType.QInt8
    symmetric: bool = False
    group_size: int | None = Field(
        default=None, description=">0: group quant, -1: channel quant, None: tensor quant")
    strategy: QuantizationStrategy | str | None = None
    scale_dtype: np.dtype = Field(default=np.dtype(np.float32))
    zp_dtype: np.dtype = Field(default=None, init=False)   # filled in by the model validator
    reduce_range: bool = False

    @field_validator("dtype", mode="before")
    @classmethod
    def _coerce_dtype(cls, value: Any):
        return QuantType.from_string(value) if isinstance(value, str) else value

    @field_validator("group_size", mode="before")
    @classmethod
    def _check_group_size(cls, value: Any):
        if value is not None and value < -1:
            raise ValueError(_BAD_GROUP.format(value, "per_channel"))
        return value

    @field_validator("strategy", mode="before")
    @classmethod
    def _coerce_strategy(cls, value: Any):
        return QuantizationStrategy(value.lower()) if isinstance(value, str) else value

    @field_validator("scale_dtype", mode="before")
    @classmethod
    def _coerce_scale_dtype(cls, value: Any):
        return value if isinstance(value, np.dtype) else np.dtype(value)

    @field_validator("scale_dtype", mode="after")
    @classmethod
    def _check_scale_dtype(cls, value: np.dtype):
        if value != np.float32:
            raise ValueError("Only float32 scale dtype is currently supported.")
        return value

    @model_validator(mode="after")
    def validate_model_after(self):
        strategy, gs = self.strategy, self.group_size
        if strategy is None:   # infer from group_size: None → tensor, -1 → channel, >0 → group
            if gs is None:
                strategy = QuantizationStrategy.TENSOR
            elif gs == -1:
                strategy = QuantizationStrategy.CHANNEL
            elif gs > 0:
                strategy = QuantizationStrategy.GROUP
            else:
                raise ValueError(_BAD_GROUP.format(gs, "channel"))
        if strategy == QuantizationStrategy.GROUP and (gs is None or gs <= 0):
            raise ValueError(
                f"strategy {strategy} requires group_size to be set to a positive value.")
        if gs is not None and gs > 0 and strategy != QuantizationStrategy.GROUP:
            raise ValueError("group_size requires strategy to be set to 'group'.")
        if self.zp_dtype is None:
            self.zp_dtype = self.dtype.np_dtype
        self.strategy = strategy
        return self


class QWeightArgs(_BaseArgs):
    """Weight quantization parameters: ``clip_ratio`` (0,1], ``mse`` search, ``algorithm`` plugin."""

    clip_ratio: float = 1.0
    mse: bool = False
    algorithm: SerializeAsAny[AlgorithmConfig] = Field(default_factory=_default_algorithm_config)

    @field_validator("algorithm", mode="before")
    @classmethod
    def _coerce_algorithm(cls, value: Any):
        return _resolve_algorithm_config(value)

    @field_validator("clip_ratio", mode="after")
    @classmethod
    def _check_clip_ratio(cls, value: float) -> float:
        if not 0.0 < value <= 1.0:
            raise ValueError(f"clip_ratio must be in (0.0, 1.0], got {value}")
        return value

    @model_validator(mode="after")
    def validate_model_after(self):
        self.algorithm.validate_weight_args(self)   # plugin-specific constraints first
        return super().validate_model_after()


class QActivationArgs(_BaseArgs):
    """Activation quantization parameters; ``is_static`` selects calibrated vs dynamic."""

    is_static: bool = True

    @field_validator("strategy", mode="after")
    @classmethod
    def _tensor_only(cls, value):
        if value not in (None, QuantizationStrategy.TENSOR):
            raise NotImplementedError("Activation quantization only supports 'tensor' strategy.")
        return QuantizationStrategy.TENSOR

    @field_validator("dtype", mode="after")
    @classmethod
    def _no_four_bit(cls, value):
        if value in _FOUR_BIT:
            raise NotImplementedError("4-bit quantization is not supported for activations.")
        return value

    @model_validator(mode="after")
    def validate_model_after(self):
        if not self.is_static and self.dtype != QuantType.QUInt8:
            raise NotImplementedError("Dynamic activation quantization only supports uint8 dtype.")
        return super().validate_model_after()


# ------------------------------------------------------------------------------------------------
# top-level configuration
# ------------------------------------------------------------------------------------------------
class QConfig(BaseModel):
    r"""All quantization parameters of one ``quantize(model, qconfig)`` call.

    ``target_op_types`` (MatMul/Gemm), ``weights`` / ``input_activations`` /
    ``output_activations``, ``format`` (qdq | qlinear), ``calibration_params`` /
    ``calibration_data`` (an array for the first model input or a dict name → array),
    ``preprocessors`` and ``ignore`` (regexes matched against node names with ``re.search``).
    """

    model_config = ConfigDict(extra="forbid", arbitrary_types_allowed=True)

    target_op_types: Sequence[str] = Field(default_factory=lambda: _SUPPORTED_OP_TYPES)
    weights: QWeightArgs | None = None
    input_activations: QActivationArgs | None = None
    output_activations: QActivationArgs | None = None
    format: QFormat | str = QFormat.QDQ
    calibration_params: CalibrationParams | None = Field(default_factory=CalibrationParams)
    calibration_data: np.ndarray | dict[str, np.ndarray] | None = None
    preprocessors: Sequence[SerializeAsAny[PreProcessingConfig]] = Field(default_factory=tuple)
    ignore: Sequence[str] = Field(default_factory=tuple)

    @field_validator("target_op_types", mode="before")
    @classmethod
    def _dedupe_op_types(cls, value):
        return tuple(sorted(set(value)))

    @field_validator("ignore", mode="before")
    @classmethod
    def _normalize_ignore(cls, value):
        if value is None:
            return ()
        return (value,) if isinstance(value, str) else tuple(value)

    @field_validator("preprocessors", mode="before")
    @classmethod
    def _coerce_preprocessors(cls, value):
        return () if value is None else tuple(_resolve_preprocessing_config(v) for v in value)

    @field_validator("format", mode="before")
    @classmethod
    def _coerce_format(cls, value):
        if not isinstance(value, str):
            return value
        try:
            return QFormat(value.lower())
        except ValueError:
            raise ValueError(f"Invalid quantization format '{value}'. Valid formats are: "
                             f"{[f.value for f in QFormat]}") from None

    @field_validator("calibration_params", mode="before")
    @classmethod
    def _coerce_calibration_params(cls, value):
        return CalibrationParams(**value) if isinstance(value, dict) else value

    def _check_qlinear_format_constraints(self) -> None:
        acts = (("input", self.input_activations), ("output", self.output_activations))
        if any(a is None for _, a in acts):
            raise ValueError("QLinear format requires both input and output activation quantization.")
        if not all(a.is_static for _, a in acts):
            raise ValueError("QLinear format requires both input and output activations "
                             "quantization to be static.")
        if self.weights.strategy == QuantizationStrategy.GROUP:
            raise NotImplementedError("QLinear format does not support grouped weight quantization.")
        if self.weights.dtype not in _EIGHT_BIT:
            raise ValueError(
                f"QLinear format supports only int8/uint8 for weights, got {self.weights.dtype}.")
        for kind, a in acts:
            if a.dtype not in _EIGHT_BIT:
                raise ValueError(f"QLinear format supports only int8/uint8 for {kind} activations, "
                                 f"got {a.dtype}.")

    @model_validator(mode="after")
    def validate_model_after(self):
        for op_type in self.target_op_types:
            if op_type not in _SUPPORTED_OP_TYPES:
                raise ValueError(f"Unsupported operator type '{op_type}' in target_op_types. "
                                 f"Supported operator types are: {_SUPPORTED_OP_TYPES}")
        acts = (self.input_activations, self.output_activations)
        if self.weights is None:
            if all(a is None for a in acts):
                return self   # nothing to quantize
            raise ValueError("Activation only quantization is not supported.")
        weights_only = all(a is None for a in acts)
        if not weights_only:
            if self.weights.dtype in _FOUR_BIT:
                raise NotImplementedError(
                    "4-bit quantization is only supported for weights_only quantization.")
            if self.weights.strategy == QuantizationStrategy.GROUP:
                raise NotImplementedError(
                    "Group quantization is only supported for weights_only quantization.")
        if all(a is not None for a in acts) and acts[0].is_static != acts[1].is_static:
            raise NotImplementedError(
                "Both input and output activations must be either both static or dynamic.")
        if self.format == QFormat.QLINEAR:
            self._check_qlinear_format_constraints()
        return self
