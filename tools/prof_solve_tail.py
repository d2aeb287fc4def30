"""How long does the solve phase of ONE rank's share at 8 GPUs take on its own (no reduces, no other
ranks)?  Runs the bench's GPTQ variant on one GPU over 1/8 of the layers: if the solve phase is as long
as at 8 GPUs, the tail is the latency of the solve chains themselves; if it is shorter, the difference
is waiting for (and competing with) the Hessian exchange."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

args = argparse.Namespace(gptq_layers=None, gptq_precision="bf16x3", gptq_streams=8)
device = torch.device("cuda", 0)
torch.cuda.set_device(device)
for model, layers in (("gemma3_1b", 3), ("gemma3_1b", 4), ("llama3_8b", 4)):
    for streams in (8, 16):
        args.gptq_streams = streams
        r = bench.run_gptq_variant(args, torch, None, device, 1, 0, model, layers=layers)
        print(f"{model} {layers} layers, {streams} solve streams: hessian {r['hessian_s']*1e3:.1f} ms, solve {r['solve_s']*1e3:.1f} ms", flush=True)
