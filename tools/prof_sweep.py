"""Sweep of reference-facing RTN configurations on one 4096 x 4096 weight: anything pathologically slow?"""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = torch.randn((4096, 4096), generator=g, device="cuda") * 0.02
rows = []
for qt, (st, gs), mse, sym, lay in itertools.product(("int8", "uint4"), (("tensor", -1), ("channel", -1), ("group", -1), ("group", 32), ("group", 128), ("group", 512), ("group", 48)),
                                                      (False, True), (False, True), ("kn", "packed_flat", "matmul_nbits")):
    if st == "group" and gs > 0 and 4096 % gs: continue
    if lay == "packed_flat" and qt != "uint4": continue
    if lay == "matmul_nbits" and (st != "group" or sym or gs in (-1, 48)): continue
    try:
        f = lambda: D.rtn_quantize(w, qt, st, gs, sym, False, 1.0, mse, layout=lay)
        f(); torch.cuda.synchronize()
        t0 = time.perf_counter(); f(); torch.cuda.synchronize()
        rows.append(((time.perf_counter() - t0) * 1e3, qt, st, gs, mse, sym, lay))
    except Exception as e:  # noqa: BLE001
        rows.append((-1.0, qt, st, gs, mse, sym, lay + " ERR " + str(e)[:60]))
for r in sorted(rows, reverse=True)[:24]:
    print("%9.3f ms  %s" % (r[0], r[1:]))
print("configs:", len(rows))
