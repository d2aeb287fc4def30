"""BF16x3 Hessian, one-CTA tiles vs CTA pairs (B200Q_HESSIAN_PAIRS=1), over K: device time and
TFLOP/s of bf16 MMAs issued (upper-triangle + diagonal tiles, 3 products)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.hessian import hessian_accumulate

shapes = [(32768, k) for k in (1024, 1152, 2048, 3072, 4096, 5120, 6912, 8192, 11008, 14336)]
for (t, k) in shapes:
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn((t, k), device="cuda", generator=g)
    h = torch.zeros((k, k), device="cuda")
    res = []
    for pairs in ("0", "1"):
        os.environ["B200Q_HESSIAN_PAIRS"] = pairs
        for _ in range(2):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision="bf16x3")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision="bf16x3")
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 4)
    f = 3 * 2.0 * t * (k * k / 2 + k * 128) / 1e9
    print(f"T={t} K={k:5d}: one-CTA {res[0]:7.3f} ms ({f/res[0]:5.0f} TF)   pairs {res[1]:7.3f} ms ({f/res[1]:5.0f} TF)   pairs/one = {res[1]/res[0]:.3f}", flush=True)
    del x, h
print("ok")
