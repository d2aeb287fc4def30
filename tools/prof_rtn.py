"""Small fixed workload for ncu: the plain and the two-tier MSE group kernels on one gate_proj-shaped
weight (4096 x 14336) and one q_proj-shaped weight, three launches each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D

torch.manual_seed(0)
for shape in ((4096, 14336), (4096, 4096)):
    w = torch.randn(shape, device="cuda") * 0.02
    for mse in (False, True):
        for _ in range(3):
            D.rtn_quantize(w, "uint4", "group", 128, False, False, 0.9, mse, layout="matmul_nbits")
    torch.cuda.synchronize()
# error of the approximate power by decade
g = torch.Generator(device="cuda"); g.manual_seed(0)
for lo, hi in ((-12, -9), (-9, -6), (-6, -3), (-3, 0), (0, 4)):
    e = torch.rand(4_000_000, device="cuda", generator=g, dtype=torch.float64) * (hi - lo) + lo
    x = (10.0 ** e).to(torch.float32)
    a = D.debug_pow_approx(x).to(torch.float64)
    ex = x.to(torch.float64) ** 2.4000000953674316
    rel = ((a - ex).abs() / ex)
    print(f"pow_approx 1e{lo}..1e{hi}: max rel err {rel.max().item():.3e} mean {rel.mean().item():.3e}")
print("ok")
