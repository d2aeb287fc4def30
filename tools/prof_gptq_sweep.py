"""GPTQ configurations on one 4096 x 4096 weight (device API, Hessian given): anything pathologically slow?"""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate
torch.manual_seed(0)
k, n = 4096, 4096
x = torch.randn((8192, k), device="cuda")
h = torch.zeros((k, k), device="cuda")
hessian_accumulate(x, h, 2.0 / 128, 0.0)
w = torch.randn((k, n), device="cuda") * 0.02
rows = []
for (st, gs), mse, act, mode, qt, sym in itertools.product((("channel", -1), ("group", 128), ("group", 32), ("tensor", -1)), (False, True), (False, True),
                                                           ("reference", "propagate"), ("int4", "uint8"), (True, False)):
    if qt == "uint8" and sym: continue
    try:
        def f():
            fac = G.hinv_cholesky_upper(h, 0.01, act)
            return G.gptq_quantize(w, fac, qt, st, gs, sym, False, 1.0, mse, 128, mode)
        f(); torch.cuda.synchronize()
        t0 = time.perf_counter(); f(); torch.cuda.synchronize()
        rows.append(((time.perf_counter() - t0) * 1e3, st, gs, mse, act, mode, qt, sym))
    except Exception as e:  # noqa: BLE001
        rows.append((-1.0, st, gs, mse, act, mode, qt, sym, "ERR " + str(e)[:80]))
for r in sorted(rows, reverse=True)[:14]:
    print("%9.2f ms  %s" % (r[0], r[1:]))
print("fastest %.2f ms; configs: %d; errors: %d" % (min(r[0] for r in rows if r[0] > 0), len(rows), sum(r[0] < 0 for r in rows)))
