"""cfg1 (RTN int8 symmetric per-tensor, two 4096 x 4096 weights) as bench.py times it (back to back:
part of the 128 MiB stays in the 126 MB L2 from one iteration to the next) and with the L2 emptied
by READING a 512 MB buffer before every iteration (clean lines: nothing to write back)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import _device as dev, device_api as D
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
ws = [torch.randn((4096, 4096), generator=gen, device="cuda") * 0.02 for _ in range(2)]
plan = D.RtnBatchPlan(ws, "int8", "tensor", -1, True, False, 1.0, False)
flush = torch.zeros(128 << 20, dtype=torch.float32, device="cuda")
def t(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
with dev.inputs_resident():
    ms = t(plan.run)
elts = sum(w.numel() for w in ws)
print(f"warm (back to back): {ms*1e3:.1f} us per 2 weights, {5.0*elts/ms/1e6:.0f} GB/s algorithmic = {5.0*elts/ms/1e6/6542.4:.3f} of the HBM copy rate")
ts = []
for _ in range(6):
    flush.sum(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts)//2]
print(f"cold (L2 emptied by reads before): {ms*1e3:.1f} us per 2 weights, {5.0*elts/ms/1e6:.0f} GB/s algorithmic = {5.0*elts/ms/1e6/6542.4:.3f}")
