"""3xTF32 Hessian at the bench's two sizes: time per call, achieved TFLOP/s, run-to-run reproducibility."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.hessian import hessian_accumulate
torch.manual_seed(0)
for (t, k) in ((131072, 4096), (65536, 14336)):
    x = torch.randn((t, k), device="cuda")
    h = torch.zeros((k, k), device="cuda")
    ref = None
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); hessian_accumulate(x, h, 2.0 / 128, 0.0, "tf32x3"); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    xs = x[:, :256].double()
    want = (2.0 / 128) * (xs.T @ xs)
    err = ((h[:256, :256].double() - want).abs().max() / want.abs().max()).item()
    print(f"chunk={os.environ.get('B200Q_HESSIAN_CHUNK','512')} T={t} K={k}: {ms:.2f} ms  {2*t*k*k/ms/1e9:.1f} TFLOP/s sq-eq, max rel err {err:.2e}")
