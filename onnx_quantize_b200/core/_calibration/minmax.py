"""MinMax activation calibrator on the GPU — mirrors the reference's
``core/_calibration/minmax.py`` (``MinMaxCalibrator`` :11-87).

``collect`` is ONE launch per batch: the streaming reduction leaves its per-CTA partial (min, max)
pairs in a device-side slot; nothing is folded, synchronised or copied back until
``compute_range`` (or ``.data``) is read, at which point all pending slots are folded and merged
in batch order by one kernel — running min/max for ``momentum == 0`` (reference :63-64), the EMA
``m·old + (1−m)·cur`` in float32 otherwise (reference :55-60).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._calibration.base import CalibrationData, Calibrator

logger = logging.getLogger(__name__)

_PAIR_CHUNK = 16   # device slots for pending batches; folded when full


class _TensorState:
    def __init__(self, device):
        self.state = torch.zeros((2,), dtype=torch.float32, device=device)
        self.valid = torch.zeros((1,), dtype=torch.int32, device=device)
        self.slots = torch.empty((_PAIR_CHUNK, D.minmax_partials_stride(), 2), dtype=torch.float32,
                                 device=device)
        self.counts = torch.zeros((_PAIR_CHUNK,), dtype=torch.int32, device=device)
        self.pending = 0


class _LazyData(dict):
    """``calibrator.data``: name → CalibrationData, materialised from the device on EVERY read access
    (item, ``get``, iteration over values / items, ``copy``, ``dict(data)``, pickling), so that a
    reader never sees the placeholder of a pending batch; entries written directly by the user (the
    reference's ``.data`` is a plain dict) stay as they are."""

    def __init__(self, owner: "MinMaxCalibrator"):
        super().__init__()
        self._owner = owner

    def _sync_all(self) -> None:
        for name in list(dict.keys(self)):
            self._owner._sync(name)

    def __getitem__(self, name):
        self._owner._sync(name)
        return super().__getitem__(name)

    def get(self, name, default=None):
        if dict.__contains__(self, name):
            return self[name]
        return default

    def setdefault(self, name, default=None):
        if dict.__contains__(self, name):
            return self[name]
        return super().setdefault(name, default)

    def pop(self, name, *default):
        self._owner._sync(name)
        return super().pop(name, *default)

    def values(self):
        self._sync_all()
        return super().values()

    def items(self):
        self._sync_all()
        return super().items()

    def copy(self) -> dict:
        self._sync_all()
        return dict(dict.items(self))

    def __iter__(self):                  # also takes dict(data) / {**data} off CPython's raw-copy fast path
        self._sync_all()
        return super().__iter__()

    def __reduce__(self):                # pickles as the plain dict the reference holds
        return (dict, (self.copy(),))


class MinMaxCalibrator(Calibrator):
    """Tracks the global minimum / maximum of every named activation tensor.

    Args:
        momentum: EMA factor in [0, 1); 0 means strict running min/max.
    """

    def __init__(self, momentum: float = 0.0):
        super().__init__()
        assert 0 <= momentum < 1, "Momentum must be in the range [0, 1)."
        self.momentum = momentum
        self.data = _LazyData(self)
        self._dev: dict[str, _TensorState] = {}
        logger.debug(f"Initialized MinMaxCalibrator with momentum={momentum}")

    # -- device side ------------------------------------------------------------------------
    def _fold(self, st: _TensorState) -> None:
        if st.pending:
            D.minmax_fold_merge(st.state, st.valid, st.slots, st.counts, st.pending, self.momentum)
            st.pending = 0

    def collect(self, name: str, array, *, resident: bool = False) -> None:
        """Fold one activation batch (numpy array or CUDA tensor) into the statistics of ``name``.

        ``resident=True`` (an extension; the reference has no such argument) asserts that ``array``
        is a CUDA tensor whose contents were complete before the previous kernel on the current
        stream was enqueued — e.g. calibration batches already in HBM — so the reduction may overlap
        the tail of that kernel (programmatic dependent launch, ``b200q_assume_inputs_resident``)."""
        x = dev.to_device_f32(array)
        st = self._dev.get(name)
        if st is None:
            st = self._dev[name] = _TensorState(x.device)
            dict.__setitem__(self.data, name, CalibrationData(None, None))
        if st.pending == _PAIR_CHUNK:
            self._fold(st)
        with dev.inputs_resident(resident and x is array):
            D.minmax_partials(x.reshape(-1), st.slots[st.pending], st.counts[st.pending:st.pending + 1])
        st.pending += 1

    def device_range(self, name: str) -> torch.Tensor:
        """f32[2] = (min, max) on the device, all collected batches folded; no host sync."""
        if name not in self._dev:
            raise KeyError(f"No calibration data collected for '{name}'")
        st = self._dev[name]
        self._fold(st)
        return st.state

    # -- host side --------------------------------------------------------------------------
    def _sync(self, name: str) -> None:
        if name in self._dev:
            lo, hi = self.device_range(name).cpu().numpy()
            entry = dict.__getitem__(self.data, name)
            entry.min_val, entry.max_val = np.float32(lo), np.float32(hi)

    def compute_range(self, name: str) -> tuple[np.ndarray, np.ndarray]:
        """(min, max) with zero included, as 0-d float32 arrays (reference minmax.py:66-87)."""
        if name not in self._dev and not dict.__contains__(self.data, name):   # .data may be filled directly
            raise KeyError(f"No calibration data collected for '{name}'")
        d = self.data[name]
        return (np.array(np.minimum(d.min_val, 0), dtype=np.float32),
                np.array(np.maximum(d.max_val, 0), dtype=np.float32))
