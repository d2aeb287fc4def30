"""The reference-facing call with pageable NumPy arrays: `_rtn_quantize(array, ...)` — where does the
time go (H2D from pageable memory, kernels, D2H + NumPy conversion)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import onnx_quantize_b200 as q
from onnx_quantize_b200 import _device as dev, device_api as D
from onnx_quantize_b200.core._algorithms.rtn import _rtn_quantize
from onnx_quantize_b200.core._algorithms.utils import _codes_to_numpy

print("stage threads", dev._STAGE_THREADS, flush=True)
rng = np.random.default_rng(0)
QT = q.QuantType.QUInt4
for shape in ((4096, 1024), (4096, 4096), (4096, 14336), (14336, 4096)):
    w = (rng.standard_normal(shape) * 0.02).astype(np.float32)
    for mse in (False, True):
        for _ in range(2):
            _rtn_quantize(w, QT, q.QuantizationStrategy.GROUP, 128, False, False, 1.0, mse, np.dtype(np.float32), QT.np_dtype)
        t0 = time.perf_counter()
        for _ in range(5):
            out = _rtn_quantize(w, QT, q.QuantizationStrategy.GROUP, 128, False, False, 1.0, mse, np.dtype(np.float32), QT.np_dtype)
        dt = (time.perf_counter() - t0) / 5
        print(f"{shape} mse={mse}: {dt*1e3:.1f} ms per call = {w.nbytes/dt/1e9:.1f} GB/s of f32 weight", flush=True)
    # the legs of the plain route, each with a synchronize on both sides
    def leg(fn, n=5):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            r = fn(); torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, r
    t_up, wd = leg(lambda: dev.to_device_f32(w))
    t_k, (codes, s, z) = leg(lambda: D.rtn_quantize(wd, QT, "group", 128, False, False, 1.0, True))
    t_dn, _ = leg(lambda: _codes_to_numpy(codes, QT))
    t_p, _ = leg(lambda: (s.cpu().numpy(), z.cpu().numpy()))
    print(f"   legs: staged upload {t_up:.2f} ms ({w.nbytes/t_up/1e6:.1f} GB/s), MSE kernel {t_k:.2f} ms, codes to a fresh NumPy array "
          f"{t_dn:.2f} ms ({codes.numel()/t_dn/1e6:.1f} GB/s), parameters {t_p:.2f} ms", flush=True)
print("ok")
