"""Small fixed workload for ncu: the streaming (no search) and the two-tier MSE group kernels on one
gate_proj-shaped weight (4096 x 14336) and one q_proj-shaped weight, three launches each."""
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
print("ok")
