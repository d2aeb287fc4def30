"""The whole Llama-3-8B-shaped set through the batched RTN entry point (cfg2a: two launches of rtn_group_nbits4_ring_kernel; round 1: one launch of the batch kernel it replaced), for
ncu: how much DRAM traffic does the launch really cause, and how busy is DRAM?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
LAYER = [(4096, 4096), (4096, 1024), (4096, 1024), (4096, 4096), (4096, 14336), (4096, 14336), (14336, 4096)]
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = torch.Generator(device="cuda"); g.manual_seed(0)
ws = [torch.randn(s, generator=g, device="cuda") * 0.02 for _ in range(layers) for s in LAYER]
plan = D.RtnBatchPlan(ws, "uint4", "group", 128, False, False, 0.9, False, layout="matmul_nbits")
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); plan.run(); b.record(); torch.cuda.synchronize()
elts = sum(w.numel() for w in ws)
print(f"{a.elapsed_time(b):.3f} ms, {elts*4.535/a.elapsed_time(b)/1e6:.0f} GB/s algorithmic")

if os.environ.get("B200Q_PROF_NAMES"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        plan.run(); torch.cuda.synchronize()
    for e in prof.key_averages():
        print("  kernel:", e.key[:90], f"{e.device_time_total/1e3:.3f} ms x{e.count}")
