"""torch plumbing around the C ABI: device buffers, streams, workspace.  Nothing numeric lives here.

PyTorch is used only as the owner of device memory and streams; the product kernels are in
libb200quant.so.  Every helper raises when no CUDA device is available — there is no CPU path.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "onnx_quantize_b200 needs a CUDA device (NVIDIA B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def to_device_f32(x, *, name: str = "array") -> torch.Tensor:
    """numpy array / torch tensor → contiguous float32 CUDA tensor (no copy if already one)."""
    dev = require_cuda()
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        if not a.flags.writeable:   # torch.from_numpy warns on read-only views; the data is only read
            a = a.copy() if a.size < (1 << 20) else np.require(a, requirements="W")
        t = torch.from_numpy(a)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.device.type != "cuda":
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


def bind_host_to_gpu_numa_node(device_index: int | None = None) -> list[int] | None:
    """Pin the calling process to the CPU cores local to the GPU's PCIe root (its NUMA node), so that
    pinned host buffers allocated afterwards are first-touched on that node and host<->device
    copies do not cross the socket interconnect — with one process per GPU on an 8-GPU box the
    eight H2D streams otherwise all pull from one socket's memory.  Best effort: returns the core
    list, or None when the topology cannot be read (no /sys entry, restricted container)."""
    import os

    try:
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        p = torch.cuda.get_device_properties(idx)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            text = f.read().strip()
        cpus: list[int] = []
        for part in text.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.extend(range(int(lo), int(hi) + 1))
            elif part:
                cpus.append(int(part))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - topology is optional information
        return None


@contextlib.contextmanager
def inputs_resident(on: bool = True):
    """Within the context the caller asserts that every input tensor handed to the library was
    complete before the previous kernel on the current stream was enqueued (weights / calibration
    batches already in HBM), so input-only kernels may overlap that kernel's tail — see
    ``b200q_assume_inputs_resident`` in include/b200q.h."""
    from onnx_quantize_b200 import _lib

    lib = _lib.load()
    prev = lib.b200q_assume_inputs_resident(int(bool(on)))
    try:
        yield
    finally:
        lib.b200q_assume_inputs_resident(prev)


_WORKSPACES: dict[tuple[int, int], torch.Tensor] = {}


def workspace(nbytes: int) -> torch.Tensor:
    """A reusable scratch buffer per (device, stream); grows on demand, never shrinks."""
    dev = require_cuda()
    key = (dev.index, stream_ptr())
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = buf
    return buf


def release_workspaces() -> None:
    _WORKSPACES.clear()
