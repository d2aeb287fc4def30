"""Single-weight launches of the HBM-bound route (uint4 g128, MatMulNBits layout): device time per
shape with an L2 flush between iterations.  B200Q_STREAM_RING=0 selects the one-tile-per-CTA kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200 import device_api as D
g = torch.Generator(device="cuda"); g.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for shape in ((4096, 1024), (4096, 4096), (4096, 14336), (14336, 4096), (1152, 6912), (2048, 2056)):
    w = torch.randn(shape, generator=g, device="cuda") * 0.02
    out = D.rtn_quantize(w, "uint4", "group", 128, False, False, 0.9, False, layout="matmul_nbits")
    ts = []
    for _ in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); D.rtn_quantize(w, "uint4", "group", 128, False, False, 0.9, False, layout="matmul_nbits", out=out); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2]
    print(f"{shape}: {t*1e3:.1f} us, {w.numel()*4.535/t/1e6:.0f} GB/s algorithmic")
