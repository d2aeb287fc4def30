"""cfg1 (per-tensor int8) and cfg3 (activation min/max) steps as bench.py runs them, plus the same
kernels launched back to back through the raw C ABI with pre-computed arguments — separates the
kernel time from the Python launch path.  Run plain for timings, under ncu for the launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from onnx_quantize_b200 import _device as dev
from onnx_quantize_b200 import _lib
from onnx_quantize_b200 import device_api as D
from onnx_quantize_b200.core._dtypes import QuantType


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


lib = _lib.load()
if "--no-pdl" not in sys.argv:
    lib.b200q_assume_inputs_resident(1)      # weights and batches are resident before anything is launched
device = torch.device("cuda")
gen = torch.Generator(device=device)
gen.manual_seed(0)
ws = [torch.randn((4096, 4096), generator=gen, device=device) * 0.02 for _ in range(2)]
acts = [torch.randn((10, 512, 4096), generator=gen, device=device) for _ in range(10)]
slots = torch.empty((10, D.minmax_partials_stride(), 2), dtype=torch.float32, device=device)
counts = torch.zeros((10,), dtype=torch.int32, device=device)
state = torch.zeros((2,), dtype=torch.float32, device=device)
valid = torch.zeros((1,), dtype=torch.int32, device=device)
rng2 = torch.zeros((2,), dtype=torch.float32, device=device)
sc = torch.empty((1,), dtype=torch.float32, device=device)
zp = torch.empty((1,), dtype=torch.uint8, device=device)
short = "--short" in sys.argv
it = 2 if short else 20


plan1 = D.RtnBatchPlan(ws, QuantType.QInt8, "tensor", -1, True, False, 1.0, False)


def cfg1():
    plan1.run()


batches = [(x.reshape(-1), slots[i], counts[i:i + 1]) for i, x in enumerate(acts)]


def cfg3():
    valid.zero_()
    for x, slot, cnt in batches:
        D.minmax_partials(x, slot, cnt)
    D.minmax_fold_merge(state, valid, slots, counts, 10, 0.0, None, rng2)
    D.qparams(rng2[0:1], rng2[1:2], QuantType.QUInt8)


print("cfg1 bench-style ms/step", time_ms(cfg1, it))
print("cfg3 bench-style ms/step", time_ms(cfg3, it))

# raw: pre-computed pointers, nothing but the C calls in the loop
st = dev.stream_ptr()
codes = [torch.empty((4096, 4096), dtype=torch.uint8, device=device) for _ in ws]
wsb = lib.b200q_rtn_workspace_bytes(4096, 4096, 0, -1, 0)
wsp = torch.empty((wsb,), dtype=torch.uint8, device=device)
args1 = [(w.data_ptr(), c.data_ptr()) for w, c in zip(ws, codes)]


def cfg1_raw():
    for wp, cp in args1:
        rc = lib.b200q_rtn_quantize(wp, 4096, 4096, 2, 0, -1, 1, 0, 1.0, 0, 0, cp, sc.data_ptr(), zp.data_ptr(),
                                    None, wsp.data_ptr(), wsb, st)
        assert rc == 0, rc


args3 = [(x.data_ptr(), x.numel(), slots[i].data_ptr(), counts[i:i + 1].data_ptr()) for i, x in enumerate(acts)]


def cfg3_raw():
    for xp, n, sp, cp in args3:
        lib.b200q_minmax_partials(xp, n, sp, cp, st)
    lib.b200q_minmax_fold_merge(state.data_ptr(), valid.data_ptr(), slots.data_ptr(), counts.data_ptr(), 10, 0.0,
                                None, rng2.data_ptr(), st)
    lib.b200q_qparams(rng2.data_ptr(), rng2.data_ptr() + 4, 1, 3, 0, 0, sc.data_ptr(), zp.data_ptr(), st)


try:
    print("cfg1 raw ms/step", time_ms(cfg1_raw, it))
except Exception as e:  # noqa: BLE001
    print("cfg1 raw failed", e)
print("cfg3 raw ms/step", time_ms(cfg3_raw, it))

if not short:
    # the same raw sequences replayed from a CUDA graph
    for name, fn in (("cfg1", cfg1_raw), ("cfg3", cfg3_raw)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            st = dev.stream_ptr()
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                st = dev.stream_ptr()
                fn()
        st = dev.stream_ptr()
        print(name, "graph ms/step", time_ms(g.replay, it))
print("ok")
