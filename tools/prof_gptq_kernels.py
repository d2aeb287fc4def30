"""Kernel-level timeline (CUPTI via torch.profiler) of one inverse-Hessian factor and one GPTQ block
loop at K = 4096 and K = 14336: which kernels the chain consists of and where its time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate

torch.manual_seed(0)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
for k, n in ((4096, 4096), (14336, 4096)):
    x = torch.randn((8192, k), device="cuda")
    h = torch.zeros((k, k), device="cuda")
    hessian_accumulate(x, h, 2.0 / 128, 0.0, prec)
    w = torch.randn((k, n), device="cuda") * 0.02
    f = G.hinv_cholesky_upper(h, 0.01, False, prec)
    G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", prec)
    torch.cuda.synchronize()
    for what in ("factor", "loop"):
        t0 = time.perf_counter()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            if what == "factor":
                f = G.hinv_cholesky_upper(h, 0.01, False, prec)
            else:
                G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", prec)
            torch.cuda.synchronize()
        tot = sum(e.device_time_total for e in prof.key_averages())
        print(f"--- K={k} N={n} {what}: {tot/1e3:.2f} ms of kernel time")
        for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:9]:
            if e.device_time_total > 0:
                print(f"  {e.device_time_total/1e3:8.3f} ms x{e.count:<4d} {e.device_time_total/max(e.count,1):8.1f} us  {e.key[:80]}")
    torch.cuda.synchronize()
    t0 = time.perf_counter(); f = G.hinv_cholesky_upper(h, 0.01, False, prec); torch.cuda.synchronize(); t1 = time.perf_counter()
    G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", prec); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"wall: factor {1e3*(t1-t0):.2f} ms, loop {1e3*(t2-t1):.2f} ms")
