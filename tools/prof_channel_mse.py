"""One CHANNEL + MSE quantization of a 4096 x 4096 weight (int8 symmetric) for ncu: the tier-1 kernel
(mse_channel_approx_kernel) and the warp-per-pair exact tier."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = torch.randn((4096, 4096), generator=g, device="cuda") * 0.02
for _ in range(3):
    D.rtn_quantize(w, "int8", "channel", -1, True, False, 1.0, True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); D.rtn_quantize(w, "int8", "channel", -1, True, False, 1.0, True); b.record(); torch.cuda.synchronize()
print(f"{a.elapsed_time(b):.3f} ms")
