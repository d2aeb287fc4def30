"""Where does the end-to-end (host buffers in, host results out) time go?  Per-step wall times of
quantize_weights_bulk on the Llama-3-8B-shaped set, with and without the D2H leg, against the plain
pinned H2D rate of the box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onnx_quantize_b200._device import bind_host_to_gpu_numa_node
from onnx_quantize_b200.core._dtypes import QuantType
from onnx_quantize_b200.pipeline import RtnSpec, quantize_weights_bulk

LAYER = [(4096, 4096), (4096, 1024), (4096, 1024), (4096, 4096), (4096, 14336), (4096, 14336), (14336, 4096)]
bind_host_to_gpu_numa_node(0)
gens = []
for j, (k, n) in enumerate(LAYER):
    g = torch.Generator(); g.manual_seed(j)
    gens.append((torch.randn((k, n), generator=g) * 0.02).pin_memory())
host_set = [gens[j] for _ in range(32) for j in range(len(LAYER))]
nbytes = sum(w.numel() * 4 for w in host_set)
for mse in (True, False):
    spec = RtnSpec(QuantType.QUInt4, "group", 128, False, False, 0.9, mse, "matmul_nbits")
    for keep in (False, True):
        quantize_weights_bulk(host_set[:7], spec, keep_on_device=keep)
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            r = quantize_weights_bulk(host_set, spec, keep_on_device=keep)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
            del r
        print(f"mse={mse} keep_on_device={keep}: " + " ".join(f"{t*1e3:.0f}ms" for t in ts) + f"  best {nbytes/min(ts)/1e9:.1f} GB/s", flush=True)
dst = [torch.empty_like(g_, device="cuda") for g_ in gens]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(8):
    for g_, d_ in zip(gens, dst):
        d_.copy_(g_, non_blocking=True)
torch.cuda.synchronize()
print(f"plain pinned H2D: {8*sum(g_.numel()*4 for g_ in gens)/(time.perf_counter()-t0)/1e9:.1f} GB/s")
print("ok")
