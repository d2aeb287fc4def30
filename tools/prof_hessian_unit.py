"""BF16x3 Hessian: tokens per MMA unit (B200Q_HESSIAN_BF16_UNIT) against time and error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.hessian import hessian_accumulate

for (t, k) in [(16384, 14336), (32768, 4096), (65536, 1152)]:
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn((t, k), device="cuda", generator=g)
    want = (x.double().T @ x.double()) / t
    for unit in (512, 1024, 2048, 4096):
        os.environ["B200Q_HESSIAN_BF16_UNIT"] = str(unit)
        h = torch.zeros((k, k), device="cuda")
        for _ in range(2):
            hessian_accumulate(x, h, 1.0 / t, 0.0, precision="bf16x3")
        err = ((h.double() - want).abs().max() / want.abs().max()).item()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision="bf16x3")
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 4
        print(f"T={t} K={k} unit={unit}: {ms:.3f} ms  {2.0*t*k*k/ms/1e9:.1f} TFLOP/s sq-eq  err {err:.2e}", flush=True)
    del x, want
print("ok")
