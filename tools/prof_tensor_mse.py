"""TENSOR / CHANNEL strategy with the MSE search: the generic exact kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
g = torch.Generator(device="cuda"); g.manual_seed(0)
for shape in ((512, 512), (1024, 1024), (4096, 4096)):
    w = torch.randn(shape, generator=g, device="cuda") * 0.02
    for st in ("tensor", "channel"):
        D.rtn_quantize(w, "int8", st, -1, True, False, 1.0, True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        D.rtn_quantize(w, "int8", st, -1, True, False, 1.0, True); torch.cuda.synchronize()
        print(f"{shape} {st} mse: {(time.perf_counter()-t0)*1e3:.2f} ms", flush=True)
print("ok")
