"""Kernel-level timeline (torch profiler / CUPTI) of one rtn_quantize call: which kernels, how long.
usage: prof_kernels.py K N qtype strategy group_size symmetric mse"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from onnx_quantize_b200 import device_api as D
k, n, qt, st, gs, sym, mse = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = torch.randn((k, n), generator=g, device="cuda") * 0.02
f = lambda: D.rtn_quantize(w, qt, st, gs, bool(sym), False, 1.0, bool(mse))
for _ in range(2):
    f()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f(); torch.cuda.synchronize()
print(f"--- {k}x{n} {qt} {st} gs={gs} sym={sym} mse={mse}")
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"  {e.device_time_total/1e3:9.3f} ms x{e.count:<3d} {e.key[:100]}")
