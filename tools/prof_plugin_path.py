"""The reference-facing call with pageable NumPy arrays: `_rtn_quantize(array, ...)` — where does the
time go (H2D from pageable memory, kernels, D2H + NumPy conversion)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import onnx_quantize_b200 as q
from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize

rng = np.random.default_rng(0)
for shape in ((4096, 4096), (4096, 14336)):
    w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    for mse in (False, True):
        for _ in range(2):
            _rtn_quantize(w, q.QuantType.QUInt4, q.QuantizationStrategy.GROUP, 128, False, False, 1.0, mse, np.dtype(np.float32), q.QuantType.QUInt4.np_dtype)
        t0 = time.perf_counter()
        for _ in range(5):
            out = _rtn_quantize(w, q.QuantType.QUInt4, q.QuantizationStrategy.GROUP, 128, False, False, 1.0, mse, np.dtype(np.float32), q.QuantType.QUInt4.np_dtype)
        dt = (time.perf_counter() - t0) / 5
        print(f"{shape} mse={mse}: {dt*1e3:.1f} ms per call = {w.nbytes/dt/1e9:.1f} GB/s of f32 weight", flush=True)
    t0 = time.perf_counter()
    for _ in range(5):
        t = torch.from_numpy(w).cuda(); torch.cuda.synchronize()
    print(f"   pageable H2D alone: {(time.perf_counter()-t0)/5*1e3:.1f} ms")
    codes = torch.empty(shape, dtype=torch.uint8, device="cuda")
    t0 = time.perf_counter()
    for _ in range(5):
        c = codes.cpu().numpy()
    print(f"   D2H of the codes to pageable: {(time.perf_counter()-t0)/5*1e3:.1f} ms")
print("ok")
