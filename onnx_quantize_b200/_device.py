"""torch plumbing around the C ABI: device buffers, streams, workspace.  Nothing numeric lives here.

PyTorch is used only as the owner of device memory and streams; the product kernels are in
libb200quant.so.  Every helper raises when no CUDA device is available — there is no CPU path.
"""
from __future__ import annotations

import contextlib
import threading
from concurrent.futures import ThreadPoolExecutor

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
        if t.is_contiguous() and not t.is_pinned() and t.numel() * 4 >= _STAGE_MIN_BYTES:
            return _upload_pageable(t, dev)
        t = t.to(dev, non_blocking=True)
    return t.contiguous()


# ---- pageable host arrays <-> device -----------------------------------------------------------------
# The reference-facing calls receive ordinary (pageable) NumPy arrays and return NumPy arrays.  A plain
# `.cuda()` of pageable memory runs at ~11 GB/s (the driver stages it single-threaded) and `.cpu()`
# into a fresh pageable tensor at ~2 GB/s (first-touch page faults) — 49 ms around a 0.8 ms kernel
# for a 4096 x 14336 weight (tools/prof_plugin_path.py).  Uploads therefore go through two pinned
# staging chunks filled by a few host threads (NumPy copies release the GIL) while the previous
# chunk's DMA is in flight; downloads land in pinned memory that backs the returned array.
_STAGE_MIN_BYTES = 8 << 20
_STAGE_CHUNK_ELEMS = (32 << 20) // 4



def _default_stage_threads() -> int:
    """Host threads that move bytes between pageable arrays and the pinned staging chunks.  Measured
    on the 16-core B200 hosts (tools/prof_host_copy.py, tools/prof_upload.py): pageable -> pinned
    runs at 35-60 GB/s from 4 threads on (it varies that much from box to box), pinned -> a FRESH
    pageable array (first-touch page faults) at 17-37 GB/s; with the chunk DMA pipelined behind the
    fill, 8 threads gave the shortest upload of a 235 MB weight (7.1 ms; 4 threads 8.3, 12 threads
    8.4)."""
    import os

    env = os.environ.get("B200Q_STAGE_THREADS")
    if env:
        return max(1, int(env))
    try:
        avail = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        avail = os.cpu_count() or 4
    return max(2, min(8, avail - 2))


_STAGE_THREADS = _default_stage_threads()
_stage_lock = threading.Lock()
_stage: dict[int, tuple] = {}
_stage_pool = None


def stage_pool() -> ThreadPoolExecutor:
    global _stage_pool
    if _stage_pool is None:
        _stage_pool = ThreadPoolExecutor(max_workers=_STAGE_THREADS, thread_name_prefix="b200q-stage")
    return _stage_pool


def parallel_copy(dst: np.ndarray, src: np.ndarray, piece: int = 1 << 20) -> list:
    """Submit ``dst[:] = src`` (flat arrays of one dtype) to the staging threads in pieces of
    ``piece`` elements; returns the futures (NumPy copies release the GIL)."""
    pool = stage_pool()
    n = src.shape[0]
    return [pool.submit(np.copyto, dst[s0:min(s0 + piece, n)], src[s0:min(s0 + piece, n)])
            for s0 in range(0, n, piece)]


def _staging(dev: torch.device):
    stage_pool()
    st = _stage.get(dev.index)
    if st is None:
        bufs = [torch.empty((_STAGE_CHUNK_ELEMS,), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        st = _stage[dev.index] = (bufs, [b.numpy() for b in bufs], evs)
    return st


def _upload_pageable(t: torch.Tensor, dev: torch.device, out: torch.Tensor | None = None) -> torch.Tensor:
    """Contiguous pageable float32 CPU tensor → CUDA tensor (``out`` or a new one) on the current stream."""
    if out is None:
        out = torch.empty(t.shape, dtype=torch.float32, device=dev)
    src = t.reshape(-1).numpy()
    dst = out.view(-1)
    total = src.shape[0]
    with _stage_lock:
        bufs, views, evs = _staging(dev)
        for i, off in enumerate(range(0, total, _STAGE_CHUNK_ELEMS)):
            b = i & 1
            evs[b].synchronize()                          # the DMA that last read this chunk is done
            m = min(_STAGE_CHUNK_ELEMS, total - off)
            step = -(-m // _STAGE_THREADS)
            parts = [(s, min(s + step, m)) for s in range(0, m, step)]
            list(_stage_pool.map(lambda se, b=b, off=off: np.copyto(views[b][se[0]:se[1]],
                                                                      src[off + se[0]:off + se[1]]), parts))
            dst[off:off + m].copy_(bufs[b][:m], non_blocking=True)
            evs[b].record()
    return out


def upload_into(src: torch.Tensor, dst: torch.Tensor) -> None:
    """``dst.copy_(src)`` for a float32 CPU tensor on the current stream: asynchronous DMA for pinned
    sources, the staged multi-threaded path for large pageable ones."""
    if src.is_pinned() or not src.is_contiguous() or src.numel() * 4 < _STAGE_MIN_BYTES:
        dst.copy_(src, non_blocking=True)
    else:
        _upload_pageable(src, dst.device, dst)


def to_numpy(t: torch.Tensor) -> np.ndarray:
    """CUDA tensor → fresh (pageable, caller-owned) NumPy array.  Large results come down by DMA
    into the pinned staging chunks and are copied out by a few host threads — the first-touch page
    faults of the new array are what makes a plain ``.cpu()`` slow — while the next chunk's DMA is in
    flight.  (Handing out pinned-backed arrays instead would need a new ``cudaHostAlloc`` for every
    result the caller keeps: slower than the copy.)"""
    if t.device.type != "cuda":
        return t.numpy()
    nbytes = t.numel() * t.element_size()
    if nbytes < _STAGE_MIN_BYTES:
        return t.cpu().numpy()
    src = t.contiguous().view(-1).view(torch.uint8)
    out = np.empty(t.shape, dtype=torch.empty((), dtype=t.dtype).numpy().dtype)
    dst = out.reshape(-1).view(np.uint8)
    chunk = _STAGE_CHUNK_ELEMS * 4
    stream = torch.cuda.current_stream(t.device)
    with _stage_lock:
        bufs, views, evs = _staging(t.device)
        bbufs = [b.view(torch.uint8) for b in bufs]
        bviews = [v.view(np.uint8) for v in views]
        offs = list(range(0, nbytes, chunk))

        def issue(i):
            b = i & 1
            m = min(chunk, nbytes - offs[i])
            bbufs[b][:m].copy_(src[offs[i]:offs[i] + m], non_blocking=True)
            evs[b].record(stream)

        for b in range(2):
            evs[b].synchronize()                      # staging chunks may still feed an earlier upload
        issue(0)
        for i in range(len(offs)):
            b = i & 1
            if i + 1 < len(offs):
                issue(i + 1)                          # the other chunk: its previous contents were copied out below
            evs[b].synchronize()
            m = min(chunk, nbytes - offs[i])
            step = -(-m // _STAGE_THREADS)
            parts = [(s0, min(s0 + step, m)) for s0 in range(0, m, step)]
            list(_stage_pool.map(lambda se, b=b, off=offs[i]: np.copyto(dst[off + se[0]:off + se[1]],
                                                                         bviews[b][se[0]:se[1]]), parts))
    return out


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
