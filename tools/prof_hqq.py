"""HQQ timing: one 4096 x 4096 weight, g128, 20 iterations (early stop off so all run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.core._algorithms.hqq import hqq_quantize_device
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = torch.randn((4096, 4096), generator=g, device="cuda") * 0.02
for es in (True, False):
    for _ in range(2):
        hqq_quantize_device(w, 128, early_stop=es)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = hqq_quantize_device(w, 128, early_stop=es, return_info=True)
    b.record(); torch.cuda.synchronize()
    print(f"early_stop={es}: {a.elapsed_time(b)/5:.3f} ms per 4096x4096 weight, best iter {int(out[3][0])}")
print("ok")
