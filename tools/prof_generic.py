"""Timing of the generic RTN routes (CHANNEL, TENSOR, odd group sizes, 8-bit) next to the tuned ones."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import device_api as D

g = torch.Generator(device="cuda"); g.manual_seed(0)
ws = {s: torch.randn(s, generator=g, device="cuda") * 0.02 for s in ((4096, 4096), (4096, 14336), (14336, 4096))}
cases = [("int8", "channel", -1, True, False, "kn"), ("uint8", "channel", -1, False, False, "kn"),
         ("int4", "channel", -1, True, False, "kn"), ("int8", "channel", -1, True, True, "kn"),
         ("int8", "tensor", -1, True, False, "kn"), ("uint4", "group", 128, False, False, "kn"),
         ("uint4", "group", 128, False, False, "matmul_nbits"), ("int8", "group", 128, True, False, "kn"),
         ("uint4", "group", 256, False, False, "kn"), ("int4", "group", 128, True, False, "packed_flat")]
with dev.inputs_resident():
    for shape, w in ws.items():
        for qt, st, gs, sym, mse, lay in cases:
            f = lambda: D.rtn_quantize(w, qt, st, gs, sym, False, 1.0, mse, layout=lay)
            for _ in range(2):
                f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                f()
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"{shape} {qt:5s} {st:7s} gs={gs:4d} sym={int(sym)} mse={int(mse)} {lay:12s}: {ms:.3f} ms  {w.numel()*4/ms/1e6:.0f} GB/s of f32 weight", flush=True)
print("ok")
