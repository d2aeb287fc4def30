"""Small fixed workload for ncu: the batched streaming kernel (no search) and the two-tier MSE kernel
on one gate_proj-shaped (4096 x 14336) and one q_proj-shaped weight, two passes each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from onnx_quantize_b200 import device_api as D

torch.manual_seed(0)
ws = [torch.randn(shape, device="cuda") * 0.02 for shape in ((4096, 14336), (4096, 4096))]
for mse in (False, True):
    plan = D.RtnBatchPlan(ws, "uint4", "group", 128, False, False, 0.9, mse, layout="matmul_nbits")
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
print("ok")
