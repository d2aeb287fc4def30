"""Static activation calibration of a MatMul / Gemm (+Relu) chain entirely on the GPU.

The reference calibrates by running the ONNX model in an ONNX Runtime session, keeping every
activation of every batch in host lists, and feeding them to the calibrator afterwards
(``core/_calibration/calibrate.py``: ``_prepare_calibration_data`` :150-179, ``_collect_activations``
:204-251, ``_set_qparams`` :254-285).  For the chains it targets (MatMul / Gemm nodes) the same
statistics come from a forward pass on the device: every layer is one tensor-core product — the
token-major BF16x3 kernel of ``dense.py`` (bias / ReLU in its epilogue, the output is the next
layer's input as it is) or, for shapes it does not take and for ``precision != "bf16x3"``, the
feature-major ``gemm_tn`` route — the calibrator's streaming min/max kernels read the activations
where they are, and nothing is ever copied to the host or kept: only ``(min, max)`` per tensor
survives.  With several ranks the batches are
sharded (``parallel.calibration``) and the ranges combined with one all-reduce.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200 import dense as DN
from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.core._algorithms.utils import _compute_qparams
from onnx_quantize_b200.core._calibration.base import CalibrationParams
from onnx_quantize_b200.core._calibration.factory import get_calibrator
from onnx_quantize_b200.parallel import calibration as PC
from onnx_quantize_b200.parallel.shard import world


@dataclass
class DenseLayer:
    """One MatMul (``bias is None``) or Gemm node: ``y = act(x @ weight + bias)``, weight (K,N)."""

    name: str
    weight: object
    bias: object = None
    activation: str | None = None        # None | "relu"


def prepare_calibration_data(data, batch_size: int, num_samples: int):
    """(num_batches, batch_size, ...) view of the first samples; the remainder that does not fill a
    batch is dropped — ``_prepare_calibration_data`` (calibrate.py:150-179)."""
    total = data.shape[0]
    num_samples = min(num_samples, total)
    data = data[:num_samples]
    if batch_size >= num_samples:
        return data.reshape((1, num_samples, *data.shape[1:]))
    nb = num_samples // batch_size
    return data[: nb * batch_size].reshape((nb, batch_size, *data.shape[1:]))


def _forward_feature_major(xt: torch.Tensor, w: torch.Tensor, bias, relu: bool, precision: str):
    """xt (K, T) → act(W^T xt + b) as (N, T)."""
    lib = _lib.load()
    y = G.gemm_tn(w, xt, None, 1.0, False, precision)
    _lib.check(lib.b200q_bias_act(y.data_ptr(), int(y.shape[0]), int(y.shape[1]), dev.ptr(bias),
                                  int(relu), dev.stream_ptr()), "b200q_bias_act")
    return y


def calibrate_mlp(layers, calibration_data, qconfig, params: CalibrationParams | None = None,
                  precision: str = "bf16x3", group=None) -> dict:
    """Ranges and quantization parameters of every layer's input (and output) activation.

    Returns ``{layer.name: {"input_scale", "input_zero_point", "input_range"[, "output_*"]}}`` —
    the values the reference stores in ``node.meta`` (calibrate.py:275).  ``qconfig.input_activations``
    / ``output_activations`` select what is calibrated, exactly as ``get_target_nodes`` does.
    """
    params = params or (qconfig.calibration_params if getattr(qconfig, "calibration_params", None)
                        else CalibrationParams())
    device = dev.require_cuda()
    lib = _lib.load()
    in_args, out_args = qconfig.input_activations, qconfig.output_activations
    calibrator = get_calibrator(params.method, momentum=params.momentum)
    ws = [dev.to_device_f32(l.weight) for l in layers]
    bs = [None if l.bias is None else dev.to_device_f32(l.bias) for l in layers]
    batches = prepare_calibration_data(calibration_data, params.batch_size, params.num_samples)
    rank, n_ranks = world()
    mine = PC.shard_batches(batches.shape[0], rank, n_ranks)
    names = []
    token_major = precision == "bf16x3" and all(DN.supported(int(w.shape[0]), int(w.shape[1])) for w in ws)
    w_planes = [DN.Planes.of_weight(w) for w in ws] if token_major else None
    x_planes: dict = {}
    for b in mine:
        x = dev.to_device_f32(batches[b])
        k0 = int(x.shape[-1])
        x2 = x.reshape(-1, k0)
        if token_major:
            for li, (layer, bias) in enumerate(zip(layers, bs)):
                if in_args is not None:
                    calibrator.collect(f"{layer.name}/input", x2)
                key = (int(x2.shape[0]), int(x2.shape[1]))
                x_planes[key] = DN.Planes.of_rows(x2, x_planes.get(key))      # one plane buffer per shape, reused
                x2 = DN.forward_planes(x_planes[key], w_planes[li], 1.0, bias, layer.activation == "relu")
                if out_args is not None:
                    calibrator.collect(f"{layer.name}/output", x2)
            continue
        xt = torch.empty((k0, x2.shape[0]), dtype=torch.float32, device=device)
        _lib.check(lib.b200q_transpose(x2.data_ptr(), int(x2.shape[0]), k0, xt.data_ptr(), dev.stream_ptr()),
                   "b200q_transpose")
        for layer, w, bias in zip(layers, ws, bs):
            if in_args is not None:
                calibrator.collect(f"{layer.name}/input", xt)
            xt = _forward_feature_major(xt, w, bias, layer.activation == "relu", precision)
            if out_args is not None:
                calibrator.collect(f"{layer.name}/output", xt)
    for layer in layers:
        if in_args is not None:
            names.append((layer.name, "input", in_args))
        if out_args is not None:
            names.append((layer.name, "output", out_args))
    # ---- combine the ranks' statistics -------------------------------------------------------
    if n_ranks > 1:
        if params.momentum == 0:
            ranges = torch.stack([calibrator.device_range(f"{ln}/{kind}") if f"{ln}/{kind}" in calibrator._dev
                                  else torch.tensor([float("inf"), float("-inf")], device=device)
                                  for ln, kind, _ in names])
            PC.allreduce_minmax(ranges, group)
        else:
            raise NotImplementedError("momentum > 0 with several ranks: gather the per-batch pairs with "
                                      "parallel.calibration.gather_batch_pairs and replay them in order")
        ranges = ranges.cpu().numpy()
    else:
        ranges = np.stack([calibrator.device_range(f"{ln}/{kind}").cpu().numpy() for ln, kind, _ in names]) \
            if names else np.zeros((0, 2), np.float32)
    out: dict = {}
    for (ln, kind, qa), (lo, hi) in zip(names, ranges):
        rmin = np.array(np.minimum(lo, 0), dtype=np.float32)          # minmax.py:84-87
        rmax = np.array(np.maximum(hi, 0), dtype=np.float32)
        scale, zp = _compute_qparams(rmin, rmax, qa.dtype, qa.symmetric, qa.reduce_range, qa.scale_dtype,
                                     qa.zp_dtype)                      # calibrate.py:276-284
        d = out.setdefault(ln, {})
        d[f"{kind}_scale"], d[f"{kind}_zero_point"], d[f"{kind}_range"] = scale, zp, (rmin, rmax)
    return out
