"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/b200q.h declares; the host mirror of the reference API validates like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import onnx_quantize_b200 as q
from onnx_quantize_b200 import _lib
from onnx_quantize_b200.core._qconfig import _ALGORITHM_REGISTRY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "b200q.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 15
    missing = [name for name in declared if not hasattr(lib, name)]
    assert not missing, f"declared in include/b200q.h but not exported: {missing}"
    bound = set(_lib.exported_symbols())
    assert not [d for d in declared if d not in bound], "ctypes signatures missing for declared symbols"
    assert _lib.load().b200q_version() == 100
    assert _lib.load().b200q_status_string(1).decode() == "matrix is not positive definite"


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _rtn_quantize(np.zeros((16, 16), np.float32), q.QuantType.QInt8, q.QuantizationStrategy.TENSOR,
                      -1, False, False, 1.0, False, np.dtype(np.float32), np.dtype(np.int8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q.core._calibration.minmax.MinMaxCalibrator().collect("x", np.zeros(4, np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "onnx_quantize_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


# ---- public configuration API (mirrors test/core/test_qconfig.py, test_dtypes.py) -------------------
def test_quant_type_surface():
    assert q.QuantType.from_string(" UInt4 ") is q.QuantType.QUInt4
    assert q.QuantType.QInt4.qrange(True) == (-7, 7) and q.QuantType.QInt4.qrange(True, True) == (-4, 3)
    assert q.QuantType.QUInt8.qrange(True) == (0, 255) and q.QuantType.QInt8.qrange(False, True) == (-64, 64)
    assert q.QuantType.QInt32.qrange(True) == (-(2**31 - 1), 2**31 - 1)
    assert q.QuantType.QUInt4.bitwidth == 4 and q.QuantType.QInt8.np_dtype == np.int8
    assert q.QuantType.QUInt4.np_dtype.itemsize == 1
    with pytest.raises(ValueError, match="Invalid quantization type"):
        q.QuantType.from_string("fp8")


def test_qweight_args_inference_and_validation():
    assert q.QWeightArgs().strategy == q.QuantizationStrategy.TENSOR
    assert q.QWeightArgs(group_size=-1).strategy == q.QuantizationStrategy.CHANNEL
    a = q.QWeightArgs(dtype="uint4", group_size=128)
    assert a.strategy == q.QuantizationStrategy.GROUP and a.zp_dtype == q.QuantType.QUInt4.np_dtype
    assert isinstance(a.algorithm, q.RTNConfig)
    for bad in (0.0, -0.1, 1.5):
        with pytest.raises(ValueError, match="clip_ratio must be in"):
            q.QWeightArgs(clip_ratio=bad)
    with pytest.raises(ValueError, match="Invalid group size"):
        q.QWeightArgs(group_size=-2)
    with pytest.raises(ValueError, match="group_size requires strategy to be set to 'group'"):
        q.QWeightArgs(group_size=32, strategy="channel")
    with pytest.raises(ValueError, match="strategy .* requires group_size"):
        q.QWeightArgs(strategy="group")
    with pytest.raises(ValueError, match="Only float32 scale dtype"):
        q.QWeightArgs(scale_dtype=np.float16)


def test_activation_args_validation():
    assert q.QActivationArgs().strategy == q.QuantizationStrategy.TENSOR
    with pytest.raises(NotImplementedError, match="only supports 'tensor' strategy"):
        q.QActivationArgs(strategy="channel")
    with pytest.raises(NotImplementedError, match="4-bit quantization is not supported"):
        q.QActivationArgs(dtype="int4")
    with pytest.raises(NotImplementedError, match="Dynamic activation quantization only supports uint8"):
        q.QActivationArgs(dtype="int8", is_static=False)


def test_qconfig_validation_messages():
    w8 = q.QWeightArgs(dtype="int8")
    act = q.QActivationArgs(dtype="uint8")
    assert q.QConfig().weights is None
    assert q.QConfig(target_op_types=["Gemm", "MatMul", "Gemm"]).target_op_types == ("Gemm", "MatMul")
    assert q.QConfig(ignore="lm_head").ignore == ("lm_head",) and q.QConfig(ignore=None).ignore == ()
    with pytest.raises(ValueError, match="Unsupported operator type.*Conv"):
        q.QConfig(target_op_types=["Conv"])
    with pytest.raises(ValueError, match="Activation only quantization is not supported"):
        q.QConfig(input_activations=act)
    with pytest.raises(NotImplementedError, match="4-bit quantization is only supported"):
        q.QConfig(weights=q.QWeightArgs(dtype="uint4"), input_activations=act)
    with pytest.raises(NotImplementedError, match="Group quantization is only supported"):
        q.QConfig(weights=q.QWeightArgs(dtype="int8", group_size=32), output_activations=act)
    with pytest.raises(NotImplementedError, match="Both input and output activations must be either both"):
        q.QConfig(weights=w8, input_activations=act, output_activations=q.QActivationArgs(dtype="uint8", is_static=False))
    with pytest.raises(ValueError, match="Invalid quantization format"):
        q.QConfig(format="nope")
    with pytest.raises(ValueError, match="QLinear format requires both input and output"):
        q.QConfig(weights=w8, input_activations=act, format="qlinear")
    ok = q.QConfig(weights=w8, input_activations=act, output_activations=act, format="qlinear")
    assert ok.format == q.QFormat.QLINEAR
    assert q.QConfig(calibration_params={"momentum": 0.5, "provider": "gpu"}).calibration_params.momentum == 0.5
    with pytest.raises(ValueError, match="Momentum must be in"):
        q.CalibrationParams(momentum=1.0)
    with pytest.raises(ValueError, match="Invalid calibration method"):
        q.CalibrationParams(method="histogram")


def test_plugin_registry_round_trip():
    assert set(_ALGORITHM_REGISTRY) >= {"rtn", "gptq"}
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int4", group_size=128, symmetric=True,
                                          algorithm=q.GPTQConfig(block_size=64, actorder=True)))
    back = q.QConfig(**cfg.model_dump())
    assert isinstance(back.weights.algorithm, q.GPTQConfig) and back.weights.algorithm.block_size == 64
    assert q.GPTQConfig.requires_calibration and not q.RTNConfig.requires_calibration
    with pytest.raises(ValueError, match="Unknown algorithm_type"):
        q.QWeightArgs(algorithm={"algorithm_type": "hqq2"})

    @q.register_algorithm_config
    class Mine(q.AlgorithmConfig):
        algorithm_type: "typing.Literal['mine']" = "mine"   # noqa: F821
    assert _ALGORITHM_REGISTRY.pop("mine") is Mine
    with pytest.raises(TypeError, match="must declare an 'algorithm_type' field"):
        q.register_algorithm_config(type("Bad", (q.AlgorithmConfig,), {}))


def test_output_shapes_match_the_oracle_layouts():
    """device_api.output_shapes (pure host logic; sizes the staging buffers of the bulk pipeline)
    against the arrays the oracle produces for the same configuration."""
    from onnx_quantize_b200 import device_api as D
    from oracle import np_oracle as O
    rng = np.random.default_rng(0)
    for k, n, qt, st, gs in ((256, 48, "uint4", "group", 128), (128, 36, "uint4", "group", 32),
                             (64, 20, "int8", "channel", -1), (96, 8, "int4", "tensor", -1),
                             (128, 16, "uint8", "group", 64), (128, 12, "uint4", "group", -1)):
        w = rng.standard_normal((k, n)).astype(np.float32)
        q, s, z = O.rtn_quantize(w, qt, st, gs)
        shp = D.output_shapes(k, n, qt, st, gs, "kn")
        assert tuple(shp[0]) == q.shape and int(np.prod(shp[1])) == np.size(s) == int(np.prod(shp[2]))
        flat = D.output_shapes(k, n, qt, st, gs, "packed_flat")
        assert flat[0] == ((k * n + 1) // 2,)
        if st == "group" and qt in ("uint4", "uint8"):
            bits = 4 if qt == "uint4" else 8
            g = gs if gs > 0 else k
            b, bs, bz = O.matmul_nbits_layout(q, s, z, g, bits)
            m = D.output_shapes(k, n, qt, st, gs, "matmul_nbits")
            assert tuple(m[0]) == b.shape and tuple(m[1]) == bs.shape and tuple(m[2]) == np.reshape(bz, (n, -1)).shape
