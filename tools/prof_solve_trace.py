"""Kernel timeline (CUPTI through torch.profiler, rank 0) of the exchange + solve phase of the GPTQ
pipeline under torchrun: when do the NCCL kernels run, when do the solve chains start and end, how
much kernel time do they hold.  usage: torchrun --nproc-per-node N tools/prof_solve_trace.py [model]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench
from onnx_quantize_b200.parallel.gptq_pipeline import GptqPipeline, GptqSpec, GptqUnit

model = sys.argv[1] if len(sys.argv) > 1 else "gemma3_1b"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(device)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
layers, groups = bench.GPTQ_MODELS[model]
tokens = bench.GPTQ_SAMPLES * bench.GPTQ_SEQ // world
gen = torch.Generator(device=device); gen.manual_seed(1 + rank)
xs = {k: torch.randn((tokens, k), generator=gen, device=device) for k in sorted({g[1] for g in groups})}
ws = {}
for _, _, shapes in groups:
    for shp in shapes:
        ws.setdefault(shp, torch.randn(shp, generator=gen, device=device) * 0.02)
units = [GptqUnit(f"l{l}.{g[0]}", g[1], [ws[s] for s in g[2]], xs[g[1]], bench.GPTQ_SAMPLES)
         for l in range(layers) for g in groups]
spec = GptqSpec("int4", "group", 128, True, False, 1.0, False, 128, 0.01, False, "propagate", "bf16x3")
pipe = GptqPipeline(8, device)
pipe.run(units, spec); torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run = pipe.run(units, spec)
    torch.cuda.synchronize()
if world > 1: dist.barrier()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
def cls(name):
    if "nccl" in name.lower(): return "nccl"
    if "hessian" in name or "split" in name or "mirror_upper" in name: return "hessian"
    return "solve"
by = {"nccl": [], "hessian": [], "solve": []}
for e in ev:
    by[cls(e.name)].append((e.time_range.start, e.time_range.end, e.name))
t0 = max(b for _, b, _ in by["hessian"])
def rel(t): return (t - t0) / 1e3
if rank == 0:
    print(f"{model} on {world} rank(s), rank 0; times in ms after the last Hessian kernel ended")
    print(f"device-timed: hessian {run.start.elapsed_time(run.hessians_done):.1f} ms, solve phase {run.hessians_done.elapsed_time(run.end):.1f} ms")
    n = sorted(by["nccl"])
    if n:
        print(f"NCCL: {len(n)} kernels, first starts {rel(n[0][0]):.2f}, last ends {rel(max(b for _, b, _ in n)):.2f}, busy {sum(b - a for a, b, _ in n)/1e3:.2f} ms")
        print("   first ten [start, end]:", [(round(rel(a), 2), round(rel(b), 2)) for a, b, _ in n[:10]])
    s = sorted(by["solve"])
    print(f"solve kernels: {len(s)}, first starts {rel(s[0][0]):.2f}, last ends {rel(max(b for _, b, _ in s)):.2f}, kernel time {sum(b - a for a, b, _ in s)/1e3:.1f} ms")
    # kernel time per 5 ms window
    import collections
    win = collections.Counter()
    for a, b, _ in s:
        win[int(rel(a) // 5)] += (b - a) / 1e3
    print("   solve kernel-ms per 5 ms window:", [round(win[i], 1) for i in range(0, max(win) + 1)])
    agg = collections.Counter(); cnt = collections.Counter()
    for a, b, nme in s:
        agg[nme[:60]] += (b - a) / 1e3; cnt[nme[:60]] += 1
    for nme, t in agg.most_common(8):
        print(f"   {t:8.2f} ms x{cnt[nme]:<5d} {nme}")
if world > 1:
    dist.destroy_process_group()
