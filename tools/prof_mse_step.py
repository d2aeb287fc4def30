"""The headline step (RTN uint4 asym g128 + MSE search over the Llama-3-8B-shaped set, MatMulNBits layout)
timed as bench.py times it, without the rest of the bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
LAYER = [(4096, 4096), (4096, 1024), (4096, 1024), (4096, 4096), (4096, 14336), (4096, 14336), (14336, 4096)]
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator(device="cuda"); g.manual_seed(0)
ws = [torch.randn(s, generator=g, device="cuda") * 0.02 for _ in range(layers) for s in LAYER]
plan = D.RtnBatchPlan(ws, "uint4", "group", 128, False, False, 0.9, True, layout="matmul_nbits")
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    plan.run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
elts = sum(w.numel() for w in ws)
print(f"{ms:.2f} ms per step, {4*elts/ms/1e6:.1f} GB/s of f32 weight")
