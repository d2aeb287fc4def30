"""Hessian kernel against a closed-form pattern (H[i][j] = (T/8)(i+1)(j+1)) and random data at small
shapes: the first thing to run when a tensor-map / swizzle / descriptor change breaks the contraction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200.hessian import hessian_accumulate
torch.manual_seed(0)
for (t, k) in ((256, 128), (64, 256), (4096, 512)):
    for mode in ("pattern", "randn"):
        if mode == "pattern":
            # X[t, i] = 1 if t % 8 == 0 else 0, scaled by (i+1): H[i,j] = (T/8)*(i+1)*(j+1)
            x = torch.zeros((t, k), device="cuda")
            x[::8, :] = torch.arange(1, k + 1, device="cuda", dtype=torch.float32)[None, :]
        else:
            x = torch.randn((t, k), device="cuda")
        for prec in ("tf32", "tf32x3"):
            h = torch.zeros((k, k), device="cuda")
            hessian_accumulate(x, h, 1.0, 0.0, precision=prec)
            torch.cuda.synchronize()
            want = (x.double().T @ x.double())
            err = ((h.double() - want).abs().max() / want.abs().max()).item()
            print(f"T={t} K={k} {mode} {prec}: rel err {err:.3e}  nnz={int((h != 0).sum())}/{k*k}")
            if err > 1e-2:
                print(" got  ", h[:3, :6].tolist())
                print(" want ", want[:3, :6].float().tolist())
                print(" got diag", h.diag()[:8].tolist(), " want diag", want.diag()[:8].float().tolist())
                print(" row 0 cols 30..36", h[0, 30:36].tolist(), want[0, 30:36].float().tolist())
                print(" row 40 cols 40..44", h[40, 40:44].tolist(), want[40, 40:44].float().tolist())
