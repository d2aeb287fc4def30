"""GPTQ solve chains (inverse factor + block loops) issued launch by launch vs replayed as CUDA graphs,
side by side on 8 streams as the pipeline runs them: are the entry points capturable, and what does a
model's solve phase cost without the host's launch work?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate
from onnx_quantize_b200.parallel.streams import StreamPool

torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = sys.argv[1] if len(sys.argv) > 1 else "gemma3_1b"
n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 8
layers, groups = bench.GPTQ_MODELS[model]
def chain(h, ws):
    f = G.hinv_cholesky_upper(h, 0.01, False, "bf16x3")
    return f, [G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", "bf16x3") for w in ws]
shapes = []
for name, k, wshapes in groups:
    x = torch.randn((8192, k), device=dev)
    h = torch.zeros((k, k), device=dev)
    hessian_accumulate(x, h, 2.0 / 128, 0.0, "bf16x3")
    shapes.append((name, h, [torch.randn(s, device=dev) * 0.02 for s in wshapes]))
    del x
pool = StreamPool(n_streams, dev)
units = [shapes[g] for _ in range(layers) for g in range(len(groups))]
costs = [float(h.shape[0]) ** 3 for _, h, _ in units]
def phase(jobs):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pool.run(jobs, costs); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
eager_jobs = [(lambda h=h, ws=ws: chain(h, ws)) for _, h, ws in units]
phase(eager_jobs)
print(f"{model}: {len(units)} units on {n_streams} streams, launch by launch: {phase(eager_jobs):.1f} ms", flush=True)
graphs = {}
def graph_job(name, h, ws):
    def job():
        s = torch.cuda.current_stream(dev)
        key = (s.cuda_stream, name)
        g = graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            g.capture_begin(capture_error_mode="thread_local")
            try:
                keep = chain(h, ws)
            finally:
                g.capture_end()
            graphs[key] = (g, keep)
            g = graphs[key]
        g[0].replay()
    return job
graph_jobs = [graph_job(name, h, ws) for name, h, ws in units]
t0 = time.perf_counter(); first = phase(graph_jobs); t1 = time.perf_counter()
print(f"   first pass (captures {len(graphs)} graphs on the way): {first:.1f} ms device, {1e3*(t1-t0):.1f} ms wall", flush=True)
print(f"   graph replays: {phase(graph_jobs):.1f} ms, again {phase(graph_jobs):.1f} ms", flush=True)
