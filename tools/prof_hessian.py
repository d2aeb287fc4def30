"""Hessian accumulation throughput on one GPU: X (T, K) fp32 resident in HBM, H <- H + a X^T X.
Prints TFLOP/s (2*T*K^2 useful flops, i.e. the full square although only the upper triangle is
computed) per precision mode, CUDA-event timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.hessian import hessian_accumulate

shapes = [(32768, 4096), (16384, 14336), (65536, 1152)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
for (t, k) in shapes:
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    x = torch.randn((t, k), device="cuda", generator=g)
    for prec in ("tf32", "tf32x3", "bf16x3"):
        h = torch.zeros((k, k), device="cuda")
        for _ in range(2):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision=prec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        a.record()
        for _ in range(n):
            hessian_accumulate(x, h, 1.0 / t, 1.0, precision=prec)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"T={t} K={k} {prec}: {ms:.3f} ms  {2.0*t*k*k/ms/1e9:.1f} TFLOP/s (square-equivalent)", flush=True)
    if t * k <= 2**28:
        want = (x.double().T @ x.double()) / t
        h = torch.zeros((k, k), device="cuda")
        for prec in ("tf32x3", "bf16x3"):
            hessian_accumulate(x, h, 1.0 / t, 0.0, precision=prec)
            print("  ", prec, "max rel err", ((h.double() - want).abs().max() / want.abs().max()).item())
    del x
print("ok")
