import os, sys
sys.path.insert(0, os.getcwd())
import torch
from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate
torch.manual_seed(0)
k, n = 2048, 4096
x = torch.randn((4096, k), device="cuda")
h = torch.zeros((k, k), device="cuda")
hessian_accumulate(x, h, 2.0 / 128, 0.0, "bf16x3")
w = torch.randn((k, n), device="cuda") * 0.02
f = G.hinv_cholesky_upper(h, 0.01, False, "bf16x3")
G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", "bf16x3")
torch.cuda.synchronize()
print("ok", f.ok)
