"""4-bit nibble packing of initializers ("layout A"), on the GPU.

Mirrors the reference's ``core/_pack.py`` (``pack`` :85-98, ``unpack`` :101-115): flat row-major,
low nibble first, an odd tail padded with 0; int4 uses the two's-complement nibble.  Non-4-bit
types pass through as a dtype cast, as in the reference.
"""
from __future__ import annotations

__all__ = ["pack", "unpack"]

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._dtypes import QuantType

_FOUR_BIT = (QuantType.QInt4, QuantType.QUInt4)


def _nibble_bytes(array) -> torch.Tensor:
    """Any integer-valued array → device bytes holding value & 0xF."""
    a = np.asarray(array)
    if a.dtype in (QuantType.QInt4.np_dtype, QuantType.QUInt4.np_dtype):
        a = a.view(np.uint8)
    host = np.ascontiguousarray(a).astype(np.uint8, copy=False).reshape(-1)
    return torch.from_numpy(host.copy()).to(dev.require_cuda())


def pack(array, quant_type):
    """→ uint8 array of ceil(size/2) bytes for int4/uint4; a plain cast otherwise."""
    if quant_type not in _FOUR_BIT:
        return np.asarray(array).astype(quant_type.np_dtype)
    return D.pack4_flat(_nibble_bytes(array)).cpu().numpy()


def unpack(array, dims, quant_type):
    """Inverse of :func:`pack`: → int8 (int4) or uint8 (uint4) values of shape ``dims``."""
    if quant_type not in _FOUR_BIT:
        return np.asarray(array).astype(quant_type.np_dtype)
    packed = np.asarray(array)
    assert packed.dtype == np.uint8, "Input data must be of type uint8"
    n = int(np.prod(dims))
    nib = D.unpack4_flat(torch.from_numpy(np.ascontiguousarray(packed)).to(dev.require_cuda()), n)
    out = nib.cpu().numpy().reshape(dims)
    if quant_type == QuantType.QInt4:   # sign-extend the nibble
        return ((out ^ 8).astype(np.int8) - 8).astype(np.int8)
    return out
