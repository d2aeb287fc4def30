"""One short pass over the kernels added late in round 1, for `ncu --set full -k regex:...`:
minmax_partials (84 MB batch), quantize_flat (per-tensor int8, 64 MiB), the BF16x3 Hessian (split
pre-pass + two fused MMA chunks, K = 4096) and the HQQ iteration kernel (4096 x 4096, g128)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._algorithms.hqq import hqq_quantize_device
from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.hessian import hessian_accumulate

dev = torch.device("cuda")
g = torch.Generator(device=dev)
g.manual_seed(0)
act = torch.randn((10, 512, 4096), generator=g, device=dev)
w = torch.randn((4096, 4096), generator=g, device=dev) * 0.02
x = torch.randn((32768, 4096), generator=g, device=dev)
h = torch.zeros((4096, 4096), device=dev)
slots = torch.empty((1, D.minmax_partials_stride(), 2), dtype=torch.float32, device=dev)
counts = torch.zeros((1,), dtype=torch.int32, device=dev)
reps = 1 if "--once" in sys.argv else 2
for _ in range(reps):
    D.minmax_partials(act.reshape(-1), slots[0], counts[0:1])
    D.rtn_quantize(w, QuantType.QInt8, "tensor", -1, True, False, 1.0, False)
    hessian_accumulate(x, h, 1.0 / 32768, 0.0, precision="bf16x3")
    hqq_quantize_device(w, 128)
torch.cuda.synchronize()
print("ok")
