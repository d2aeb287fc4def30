"""Timing of the routes ragged shapes take (N % 4 != 0, group sizes that are not multiples of 32,
channel strategy on them): rowstats_cols_kernel + quantize_cols_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D

g = torch.Generator(device="cuda"); g.manual_seed(0)
cases = [((4096, 4098), "int8", "channel", -1, True), ((4096, 4098), "uint4", "group", 128, False),
         ((4080, 4096), "uint4", "group", 48, False), ((4080, 4096), "int8", "group", 24, True),
         ((4096, 50257), "int8", "channel", -1, True), ((1152, 6913), "uint4", "group", 128, False),
         ((4096, 4096), "int8", "channel", -1, True), ((4096, 4096), "uint4", "group", 128, False)]
for shape, qt, st, gs, sym in cases:
    w = torch.randn(shape, generator=g, device="cuda") * 0.02
    f = lambda: D.rtn_quantize(w, qt, st, gs, sym, False, 1.0, False)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    alg = w.numel() * 5.0
    print(f"{shape} {qt:5s} {st:7s} gs={gs:4d}: {ms:.3f} ms  {alg/ms/1e6:.0f} GB/s algorithmic (5 B/elt) = {alg/ms/1e6/6542.4:.2f} of the HBM copy rate", flush=True)
    del w
print("ok")
