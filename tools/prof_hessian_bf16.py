"""BF16x3 Hessian: sweep of the pre-split chunk length (B200Q_HESSIAN_BF16_CHUNK) — run plain for
timings; with --one for a single configuration under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.hessian import hessian_accumulate

one = "--one" in sys.argv
shapes = [(32768, 14336 if "--big" in sys.argv else 4096)] if one else [(65536, 4096), (32768, 14336), (131072, 1152)]
chunks = [0] if one else [0, 1024, 2048, 4096, 8192, 16384, 32768]
for (t, k) in shapes:
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn((t, k), device="cuda", generator=g)
    for c in chunks:
        if c:
            os.environ["B200Q_HESSIAN_BF16_CHUNK"] = str(c)
        else:
            os.environ.pop("B200Q_HESSIAN_BF16_CHUNK", None)
        h = torch.zeros((k, k), device="cuda")
        for _ in range(1 if one else 2):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision="bf16x3")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 1 if one else 4
        a.record()
        for _ in range(n):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision="bf16x3")
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"T={t} K={k} chunk={c or 'default'}: {ms:.3f} ms  {2.0*t*k*k/ms/1e9:.1f} TFLOP/s (square-equivalent)", flush=True)
    del x
print("ok")
