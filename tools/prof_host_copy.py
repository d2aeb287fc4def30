"""Host-side copy rates that bound the plugin seam (pageable NumPy in, NumPy out): multi-threaded
memcpy pageable -> pinned, first-touch of a fresh pageable array, pinned H2D / D2H DMA, alone and
both directions at once."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
n = 64 << 20                       # 256 MB of f32
src = np.random.default_rng(0).standard_normal(n, dtype=np.float32)
pin = torch.empty(n, dtype=torch.float32, pin_memory=True)
pv = pin.numpy()
for threads in (1, 2, 4, 8, 12, 16):
    pool = ThreadPoolExecutor(threads)
    step = -(-n // threads)
    parts = [(s, min(s + step, n)) for s in range(0, n, step)]
    list(pool.map(lambda p: np.copyto(pv[p[0]:p[1]], src[p[0]:p[1]]), parts))
    t0 = time.perf_counter()
    for _ in range(5):
        list(pool.map(lambda p: np.copyto(pv[p[0]:p[1]], src[p[0]:p[1]]), parts))
    dt = (time.perf_counter() - t0) / 5
    # first touch of a fresh pageable destination
    t0 = time.perf_counter()
    for _ in range(3):
        dst = np.empty(n, dtype=np.float32)
        list(pool.map(lambda p: np.copyto(dst[p[0]:p[1]], pv[p[0]:p[1]]), parts))
        del dst
    dt2 = (time.perf_counter() - t0) / 3
    print(f"threads {threads:2d}: pageable->pinned {4*n/dt/1e9:6.1f} GB/s   pinned->fresh pageable {4*n/dt2/1e9:6.1f} GB/s", flush=True)
    pool.shutdown()
d = torch.empty(n, dtype=torch.float32, device="cuda")
d2 = torch.empty(n, dtype=torch.float32, device="cuda")
pin2 = torch.empty(n, dtype=torch.float32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it
print(f"H2D pinned {4*n/t(lambda: d.copy_(pin, non_blocking=True))/1e9:.1f} GB/s")
print(f"D2H pinned {4*n/t(lambda: pin2.copy_(d2, non_blocking=True))/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(pin, non_blocking=True)
    with torch.cuda.stream(s2): pin2.copy_(d2, non_blocking=True)
print(f"H2D+D2H at once: {4*n/t(both)/1e9:.1f} GB/s each direction")
print("ok")
