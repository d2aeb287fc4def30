"""Fixed GPTQ workload for ncu / timing: inverse-Hessian factor + block loop for a q_proj-shaped
(4096x4096) and a down_proj-shaped (14336x4096) weight, Hessians from 8192 random tokens."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate

torch.manual_seed(0)
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
for k, n in ((4096, 4096), (14336, 4096)):
    x = torch.randn((8192, k), device="cuda")
    h = torch.zeros((k, k), device="cuda")
    hessian_accumulate(x, h, 2.0 / 128, 0.0, prec)
    w = torch.randn((k, n), device="cuda") * 0.02
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f = G.hinv_cholesky_upper(h, 0.01, False, prec)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", prec)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"K={k} N={n} rep{rep}: hinv {1e3 * (t1 - t0):.2f} ms, block loop {1e3 * (t2 - t1):.2f} ms, ok={f.ok}")
print("ok")
