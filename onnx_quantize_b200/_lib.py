"""ctypes binding of libb200quant.so (the C ABI declared in include/b200q.h).

There is deliberately no fallback: if the shared library has not been built
(``python __graft_entry__.py`` / ``onnx_quantize_b200/csrc/build.py``) or no CUDA device is
present, every numeric entry point of this package raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200quant.so")

# enums of include/b200q.h
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4
NOT_POSITIVE_DEFINITE = 1
MARGINAL_PIVOT = 2
QTYPE = {"int4": 0, "uint4": 1, "int8": 2, "uint8": 3}
STRATEGY = {"tensor": 0, "channel": 1, "group": 2}
LAYOUT = {"kn": 0, "packed_flat": 1, "matmul_nbits": 2}
PRECISION = {"tf32": 0, "tf32x3": 1, "fp32": 2, "bf16x3": 3}
GPTQ_MODE = {"reference": 0, "propagate": 1}


class RtnJob(ctypes.Structure):
    """struct b200q_rtn_job of include/b200q.h."""

    _fields_ = [("W", ctypes.c_void_p), ("K", ctypes.c_int64), ("N", ctypes.c_int64),
                ("out_codes", ctypes.c_void_p), ("out_scale", ctypes.c_void_p),
                ("out_zp", ctypes.c_void_p), ("out_mse_info", ctypes.c_void_p)]


class B200QuantError(RuntimeError):
    """A libb200quant call failed (CUDA error, workspace, unsupported configuration)."""


_lock = threading.Lock()
_lib = None

_i64, _i32, _f32, _f64 = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double
_ptr, _sz = ctypes.c_void_p, ctypes.c_size_t

_SIGNATURES = {
    "b200q_version": (ctypes.c_int, []),
    "b200q_status_string": (ctypes.c_char_p, [_i32]),
    "b200q_last_error": (ctypes.c_char_p, []),
    "b200q_launch_count": (ctypes.c_longlong, []),
    "b200q_assume_inputs_resident": (_i32, [_i32]),
    "b200q_rtn_workspace_bytes": (_sz, [_i64, _i64, _i32, _i64, _i32]),
    "b200q_rtn_quantize": (_i32, [_ptr, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _f64, _i32, _i32,
                                   _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "b200q_rtn_batch_workspace_bytes": (_sz, [_ptr, _i64, _i32, _i64, _i32]),
    "b200q_rtn_quantize_batch": (_i32, [_ptr, _i64, _i32, _i32, _i64, _i32, _i32, _f64, _i32, _i32,
                                         _ptr, _sz, _ptr]),
    "b200q_mse_error_table": (_i32, [_ptr, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _ptr, _ptr,
                                      _sz, _ptr]),
    "b200q_debug_pow_approx": (_i32, [_ptr, _i64, _ptr, _ptr]),
    "b200q_row_ranges": (_i32, [_ptr, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _f64, _i32, _ptr,
                                 _ptr, _ptr, _sz, _ptr]),
    "b200q_quantize_with_qparams": (_i32, [_ptr, _i64, _i64, _i32, _i32, _i64, _i32, _i32, _ptr,
                                            _ptr, _ptr, _ptr]),
    "b200q_qparams": (_i32, [_ptr, _ptr, _i64, _i32, _i32, _i32, _ptr, _ptr, _ptr]),
    "b200q_dequantize": (_i32, [_ptr, _i64, _i64, _i32, _i32, _i64, _ptr, _ptr, _ptr, _ptr]),
    "b200q_quantize_bias": (_i32, [_ptr, _i64, _ptr, _i64, _f32, _ptr, _ptr, _ptr]),
    "b200q_pack4_flat": (_i32, [_ptr, _i64, _ptr, _ptr]),
    "b200q_unpack4_flat": (_i32, [_ptr, _i64, _ptr, _ptr]),
    "b200q_pack_matmul_nbits": (_i32, [_ptr, _i64, _i64, _i64, _i32, _ptr, _ptr, _ptr, _ptr]),
    "b200q_minmax_workspace_bytes": (_sz, [_i64]),
    "b200q_minmax_reduce": (_i32, [_ptr, _i64, _ptr, _ptr, _sz, _ptr]),
    "b200q_minmax_merge": (_i32, [_ptr, _ptr, _ptr, _i64, _f64, _ptr]),
    "b200q_minmax_partials_stride": (_sz, []),
    "b200q_minmax_partials": (_i32, [_ptr, _i64, _ptr, _ptr, _ptr]),
    "b200q_minmax_fold_merge": (_i32, [_ptr, _ptr, _ptr, _ptr, _i64, _f64, _ptr, _ptr, _ptr]),
}
_OPTIONAL_SIGNATURES = {
    "b200q_hessian_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "b200q_hessian_accumulate": (_i32, [_ptr, _i64, _i64, _f32, _f32, _ptr, _i32, _ptr, _sz, _ptr]),
    "b200q_hinv_workspace_bytes": (_sz, [_i64]),
    "b200q_hinv_cholesky_upper": (_i32, [_ptr, _i64, _f64, _i32, _ptr, _ptr, _ptr, _ptr, _i32, _ptr,
                                          _sz, _ptr]),
    "b200q_gptq_workspace_bytes": (_sz, [_i64, _i64, _i32, _i64, _i32, _i64]),
    "b200q_gptq_quantize": (_i32, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _i32, _i32, _i64, _i32, _i32,
                                    _f64, _i32, _i64, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _sz,
                                    _ptr]),
    "b200q_awq_workspace_bytes": (_sz, [_i64, _i64, _i32, _i64]),
    "b200q_awq_abs_sum": (_i32, [_ptr, _i64, _i64, _ptr, _ptr]),
    "b200q_awq_weight_scale": (_i32, [_ptr, _i64, _i64, _i32, _i64, _ptr, _ptr, _sz, _ptr]),
    "b200q_awq_loss": (_i32, [_ptr, _i64, _i64, _ptr, _ptr, _f64, _i32, _i32, _i64, _i32, _i32, _f64, _i32,
                               _ptr, _ptr, _sz, _ptr]),
    "b200q_dequantize_float_zp": (_i32, [_ptr, _i64, _i64, _i32, _i32, _i64, _ptr, _ptr, _ptr, _ptr]),
    "b200q_hqq_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "b200q_hqq_quantize": (_i32, [_ptr, _i64, _i64, _i32, _i64, _i32, _f64, _i32, _f64, _f64, _f64, _i32, _i32,
                                   _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "b200q_col_abs_max": (_i32, [_ptr, _i64, _i64, _ptr, _ptr]),
    "b200q_row_abs_max": (_i32, [_ptr, _i64, _i64, _ptr, _ptr]),
    "b200q_scale_rows": (_i32, [_ptr, _i64, _i64, _ptr, _ptr, _ptr]),
    "b200q_transpose": (_i32, [_ptr, _i64, _i64, _ptr, _ptr]),
    "b200q_dense_planes_bytes": (_sz, [_i64, _i64]),
    "b200q_dense_split_rows": (_i32, [_ptr, _i64, _i64, _ptr, _sz, _ptr]),
    "b200q_dense_split_transposed": (_i32, [_ptr, _i64, _i64, _ptr, _sz, _ptr]),
    "b200q_dense_forward_planes": (_i32, [_ptr, _i64, _ptr, _i64, _i64, _f32, _ptr, _i32, _ptr, _i64, _ptr]),
    "b200q_bias_act": (_i32, [_ptr, _i64, _i64, _ptr, _i32, _ptr]),
    "b200q_gemm_tn": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _f32, _i32, _i32,
                              _ptr]),
}


def exported_symbols():
    """Names include/b200q.h declares; used by the CPU test that the library exports them all."""
    return sorted(list(_SIGNATURES) + list(_OPTIONAL_SIGNATURES))


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(`python -c 'import __graft_entry__ as g; g.build()'` at the repo root). "
                "onnx_quantize_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in {**_SIGNATURES, **_OPTIONAL_SIGNATURES}.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                if name in _OPTIONAL_SIGNATURES:
                    continue
                raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it") from None
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int, what: str = "libb200quant call") -> int:
    """0 → ok; <0 → raise; >0 → numeric status handed back to the caller."""
    if status >= 0:
        return status
    lib = load()
    text = lib.b200q_last_error().decode() or lib.b200q_status_string(status).decode()
    if status == ERR_INVALID_ARG:
        raise ValueError(f"{what}: {text}")
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {text}")
    raise B200QuantError(f"{what}: {text}")
