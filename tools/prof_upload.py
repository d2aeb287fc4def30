"""Timeline of the staged upload of a pageable array (what `_device._upload_pageable` does): host
fill of a pinned chunk by T threads, DMA of the chunk, alone and pipelined, per chunk size."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

n = 4096 * 14336
src = (np.random.default_rng(0).standard_normal(n, dtype=np.float32))
dst = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for threads in (4, 8, 12):
    pool = ThreadPoolExecutor(threads)
    for chunk_mb in (8, 16, 32, 64):
        ce = (chunk_mb << 20) // 4
        bufs = [torch.empty(ce, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        views = [b.numpy() for b in bufs]
        evs = [torch.cuda.Event(), torch.cuda.Event()]

        def fill(b, off, m):
            step = -(-m // threads)
            parts = [(s, min(s + step, m)) for s in range(0, m, step)]
            list(pool.map(lambda p: np.copyto(views[b][p[0]:p[1]], src[off + p[0]:off + p[1]]), parts))

        def run(do_fill, do_dma):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            t_fill = t_wait = 0.0
            for i, off in enumerate(range(0, n, ce)):
                b = i & 1
                m = min(ce, n - off)
                a = time.perf_counter()
                evs[b].synchronize()
                c = time.perf_counter()
                if do_fill:
                    fill(b, off, m)
                d = time.perf_counter()
                if do_dma:
                    dst[off:off + m].copy_(bufs[b][:m], non_blocking=True)
                    evs[b].record()
                t_wait += c - a
                t_fill += d - c
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3, t_fill * 1e3, t_wait * 1e3

        for _ in range(2):
            run(True, True)
        both = run(True, True)
        fo = run(True, False)
        do = run(False, True)
        print(f"threads {threads:2d} chunk {chunk_mb:2d} MB: pipelined {both[0]:6.2f} ms (fill {both[1]:.2f}, wait {both[2]:.2f}) = {4*n/both[0]/1e6:5.1f} GB/s | "
              f"fill only {fo[0]:6.2f} ms = {4*n/fo[0]/1e6:5.1f} GB/s | DMA only {do[0]:6.2f} ms = {4*n/do[0]/1e6:5.1f} GB/s", flush=True)
        del bufs, views
    pool.shutdown()
# one plain pageable copy for comparison
t0 = time.perf_counter(); dst.copy_(torch.from_numpy(src)); torch.cuda.synchronize()
print(f"plain pageable .copy_: {(time.perf_counter()-t0)*1e3:.2f} ms")
# cudaHostRegister the source in place, then DMA
import ctypes
rt = torch.cuda.cudart()
t0 = time.perf_counter()
rc = rt.cudaHostRegister(src.ctypes.data, src.nbytes, 0)
t1 = time.perf_counter()
srct = torch.from_numpy(src)
dst.copy_(srct, non_blocking=True); torch.cuda.synchronize()
t2 = time.perf_counter()
rt.cudaHostUnregister(src.ctypes.data)
t3 = time.perf_counter()
print(f"cudaHostRegister rc={rc}: register {(t1-t0)*1e3:.2f} ms, copy {(t2-t1)*1e3:.2f} ms, unregister {(t3-t2)*1e3:.2f} ms")
print("ok")
