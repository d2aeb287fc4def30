"""Small run of the kernels written in round 2 for compute-sanitizer (memcheck): the persistent ring
kernel (multi-step CTAs, ragged last column tile, odd group count), the CHANNEL two-tier MSE kernels,
the ragged-shape column walkers, both BF16x3 Hessian kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.hessian import hessian_accumulate
g = torch.Generator(device="cuda"); g.manual_seed(0)
ws = [torch.randn(s, generator=g, device="cuda") * 0.02 for s in [(640, 2064), (384, 4096), (1152, 1040), (128, 48)] * 3]
D.rtn_quantize_batch(ws, "uint4", "group", 128, False, False, 0.9, False, layout="matmul_nbits")
D.rtn_quantize(torch.randn((2048, 2064), generator=g, device="cuda"), "uint4", "group", 128, False, False, 1.0, False, layout="matmul_nbits")
D.rtn_quantize(torch.randn((300, 70), generator=g, device="cuda"), "int8", "channel", -1, True, False, 1.0, True)
D.rtn_quantize(torch.randn((257, 67), generator=g, device="cuda"), "uint4", "group", -1, False, False, 0.9, False)
D.rtn_quantize(torch.randn((96, 50), generator=g, device="cuda"), "int4", "group", 24, True, False, 1.0, False)
for pairs in ("0", "1"):
    os.environ["B200Q_HESSIAN_PAIRS"] = pairs
    x = torch.randn((1500, 352), generator=g, device="cuda")
    h = torch.zeros((352, 352), device="cuda")
    hessian_accumulate(x, h, 0.5, 0.0, "bf16x3")
torch.cuda.synchronize()
print("ok")
