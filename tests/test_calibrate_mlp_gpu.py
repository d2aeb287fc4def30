"""On-device static calibration of a MatMul / Gemm (+Relu) chain (onnx_quantize_b200/calibrate_mlp.py)
against a NumPy restatement of the reference flow: `_prepare_calibration_data` batching, a float32
forward per batch, `MinMaxCalibrator.collect` per batch and tensor, `compute_range`,
`_compute_qparams` (core/_calibration/calibrate.py:150-179, :254-285).  The first layer's input is
the calibration data itself, so its range / scale / zero point are bit-identical; deeper tensors
come out of a tensor-core product (3xTF32) instead of NumPy's sgemm: ranges agree to 2e-5."""
import numpy as np
import pytest

import onnx_quantize_b200 as q
from onnx_quantize_b200.calibrate_mlp import DenseLayer, calibrate_mlp, prepare_calibration_data
from oracle import np_oracle as O
from tests.helpers import bits

pytestmark = pytest.mark.gpu


def _mlp(rng):
    dims = [96, 160, 64, 32]
    layers = []
    for i in range(3):
        w = (rng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i])).astype(np.float32)
        b = (rng.standard_normal(dims[i + 1]) * 0.1).astype(np.float32) if i != 1 else None
        layers.append(DenseLayer(f"fc{i}", w, b, "relu" if i < 2 else None))
    return layers


def _oracle(layers, data, params, in_qt, out_qt, sym):
    batches = data[: min(params.num_samples, data.shape[0])]
    nb = max(batches.shape[0] // params.batch_size, 1) if params.batch_size < batches.shape[0] else 1
    per = params.batch_size if params.batch_size < batches.shape[0] else batches.shape[0]
    cal = O.MinMax(params.momentum)
    for b in range(nb):
        x = batches[b * per:(b + 1) * per]
        for l in layers:
            cal.collect(l.name + "/input", x)
            x = np.matmul(x, l.weight)
            if l.bias is not None:
                x = x + l.bias
            if l.activation == "relu":
                x = np.maximum(x, 0)
            cal.collect(l.name + "/output", x)
    out = {}
    for l in layers:
        for kind, qt in (("input", in_qt), ("output", out_qt)):
            lo, hi = cal.compute_range(f"{l.name}/{kind}")
            s, z = O.qparams(lo, hi, qt, sym, False)
            out[(l.name, kind)] = (lo, hi, s, z)
    return out


@pytest.mark.parametrize("momentum", [0.0, 0.9])
@pytest.mark.parametrize("in_qt,out_qt,sym", [("uint8", "uint8", False), ("int8", "int8", True)])
def test_mlp_calibration_matches_reference_flow(cuda, rng, momentum, in_qt, out_qt, sym):
    layers = _mlp(rng)
    data = rng.standard_normal((37, 6, 96)).astype(np.float32)
    params = q.CalibrationParams(num_samples=35, batch_size=10, momentum=momentum)
    assert prepare_calibration_data(data, 10, 35).shape == (3, 10, 6, 96)      # remainder dropped
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int8", symmetric=True),
                    input_activations=q.QActivationArgs(dtype=in_qt, symmetric=sym, is_static=True),
                    output_activations=q.QActivationArgs(dtype=out_qt, symmetric=sym, is_static=True),
                    calibration_params=params)
    got = calibrate_mlp(layers, data, cfg)
    want = _oracle(layers, data, params, in_qt, out_qt, sym)
    for l in layers:
        for kind in ("input", "output"):
            lo, hi, s, z = want[(l.name, kind)]
            g = got[l.name]
            glo, ghi = g[f"{kind}_range"]
            if l.name == "fc0" and kind == "input":
                assert np.array_equal(bits(glo), bits(lo)) and np.array_equal(bits(ghi), bits(hi))
                assert np.array_equal(bits(g["input_scale"]), bits(s)) and int(g["input_zero_point"]) == int(z)
                continue
            span = float(hi - lo)
            assert abs(float(glo) - float(lo)) <= 2e-5 * span and abs(float(ghi) - float(hi)) <= 2e-5 * span
            np.testing.assert_allclose(g[f"{kind}_scale"], s, rtol=5e-5)
            assert abs(int(g[f"{kind}_zero_point"]) - int(z)) <= 1
            assert g[f"{kind}_scale"].dtype == np.float32 and g[f"{kind}_scale"].shape == ()


def test_only_requested_tensors_are_calibrated(cuda, rng):
    layers = _mlp(rng)
    data = rng.standard_normal((8, 4, 96)).astype(np.float32)
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int8"), input_activations=q.QActivationArgs(dtype="uint8", is_static=True))
    got = calibrate_mlp(layers, data, cfg, q.CalibrationParams(num_samples=8, batch_size=100))
    assert all("input_scale" in v and "output_scale" not in v for v in got.values()) and len(got) == 3
