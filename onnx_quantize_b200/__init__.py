"""onnx_quantize_b200 — the numeric hot path of AyoubMDL/onnx_quantize on NVIDIA B200 (sm_100a).

Public names follow the reference package (``onnx_quantize/__init__.py``): ``QConfig``,
``QWeightArgs``, ``QActivationArgs``, ``QuantType``, ``QuantizationStrategy``, ``QFormat``,
``CalibrationParams``, ``CalibrationMethod``, ``RTNConfig``, ``GPTQConfig``, ``HqqConfig``, ``quantize``,
``set_log_level``.  All arithmetic runs in ``lib/libb200quant.so`` (hand-written CUDA, C ABI in
``include/b200q.h``); there is no CPU fallback.
"""
from onnx_quantize_b200._logging import *  # noqa: F401,F403
from onnx_quantize_b200.core._algorithms.gptq import GPTQConfig  # noqa: F401
from onnx_quantize_b200.core._algorithms.hqq import HqqConfig  # noqa: F401
from onnx_quantize_b200.core._algorithms.rtn import RTNConfig  # noqa: F401
from onnx_quantize_b200.core._calibration.base import *  # noqa: F401,F403
from onnx_quantize_b200.core._dtypes import *  # noqa: F401,F403
from onnx_quantize_b200.core._qconfig import (  # noqa: F401
    AlgorithmConfig,
    PreProcessingConfig,
    QActivationArgs,
    QConfig,
    QFormat,
    QuantizationStrategy,
    QWeightArgs,
    register_algorithm_config,
    register_preprocessing_config,
)
from onnx_quantize_b200.quantize import *  # noqa: F401,F403

__version__ = "0.1.0"
