"""One short pass for `ncu --set full`: the token-major BF16x3 dense kernel (51200 x 4096 x 4096, the
cfg3 layer), its two split kernels, and the slab kernels of the CHANNEL route (4096 x 14336 int8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from onnx_quantize_b200 import dense as DN
from onnx_quantize_b200 import device_api as D

dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(0)
x = torch.randn((51200, 4096), generator=g, device=dev)
w = torch.randn((4096, 4096), generator=g, device=dev) * 0.02
b = torch.randn((4096,), generator=g, device=dev)
wc = torch.randn((4096, 14336), generator=g, device=dev) * 0.02
reps = 1 if "--once" in sys.argv else 2
for _ in range(reps):
    wp = DN.Planes.of_weight(w)
    y = DN.dense_forward(x, wp, b, True)
    D.rtn_quantize(wc, "int8", "channel", -1, True)
torch.cuda.synchronize()
print("ok")
