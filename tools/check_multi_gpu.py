"""Correctness of the N > 1 paths over NCCL (run under torchrun on 2+ GPUs of one box):
  1. parallel.shard.quantize_weights_sharded: the triples gathered on rank 0 (packed tensors, point to
     point) equal the oracle's for every weight, in dtype / shape / value;
  2. parallel.gptq_pipeline.GptqPipeline: tokens split over the ranks, Hessians reduced to the owners,
     solves on the owners — every unit's codes against the single-process oracle fed with ALL tokens
     (north_star gate: <= 0.1 % of the codes differ, each by +-1).
Prints one line per check on rank 0 and exits non-zero on a mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import onnx_quantize_b200 as q
from onnx_quantize_b200.parallel.gptq_pipeline import GptqPipeline, GptqSpec, GptqUnit
from onnx_quantize_b200.parallel.shard import quantize_weights_sharded
from onnx_quantize_b200.pipeline import RtnSpec
from oracle import np_oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
device = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=device)
ok = True

# ---- 1. sharded RTN ------------------------------------------------------------------------------
rng = np.random.default_rng(0)                       # the same model on every rank
shapes = [(512, 256), (256, 1024), (1024, 128), (384, 640), (128, 128), (768, 96), (256, 256)]
weights = {f"w{i}": (rng.standard_normal(s) * 0.02).astype(np.float32) for i, s in enumerate(shapes)}
for qt, strategy, gs, sym, mse in (("uint4", "group", 128, False, True), ("int8", "channel", -1, True, False),
                                   ("int4", "tensor", -1, True, False)):
    spec = RtnSpec(q.QuantType.from_string(qt), strategy, gs, sym, False, 0.9, mse, "kn")
    merged = quantize_weights_sharded(weights, spec, publish=False)
    if rank == 0:
        bad = 0
        for name, w in weights.items():
            want = O.rtn_quantize(w, qt, strategy, gs, sym, False, 0.9, mse)
            for a, b in zip(merged[name], want):
                if a.dtype != b.dtype or a.shape != b.shape or not np.array_equal(
                        np.asarray(a).astype(np.float32).view(np.uint32), np.asarray(b).astype(np.float32).view(np.uint32)):
                    bad += 1
        print(f"sharded RTN {qt} {strategy} mse={mse} over {world} ranks: {len(weights)} weights, {bad} mismatching arrays", flush=True)
        ok = ok and bad == 0
    else:
        assert merged is None

# ---- 2. GPTQ pipeline ----------------------------------------------------------------------------
rng = np.random.default_rng(1)
units_np = []
for u, (k, ns) in enumerate(((512, (256, 128)), (256, (384,)), (384, (128, 128, 64)), (512, (128, 64)),
                             (256, (256,)), (512, (64,)), (512, (256, 64)))):
    x = (rng.standard_normal((16, 64, k)) * rng.uniform(0.5, 2.0, k)).astype(np.float32)     # 16 samples
    units_np.append((f"u{u}", k, x, [(rng.standard_normal((k, n)) * 0.05).astype(np.float32) for n in ns]))
per = 16 // world
units = [GptqUnit(name, k, [torch.from_numpy(w).to(device) for w in ws],
                  torch.from_numpy(x[rank * per:(rank + 1) * per].reshape(-1, k)).to(device), 16)
         for name, k, x, ws in units_np]
run = GptqPipeline(4, device).run(units, GptqSpec("int4", "group", 128, True, mode="propagate", precision="bf16x3"))
if rank == 0:
    from onnx_quantize_b200.parallel.gptq_pipeline import exchange_plan
    costs = [u.solve_cost() for u in units]
    order = sorted(range(len(units)), key=lambda i: (-costs[i], i))
    plan = exchange_plan([u.k for u in units], order, [run.owner[u.name] for u in units], world)
    print(f"exchange steps: {[[units[i].name for i in ex] for ex in plan]} "
          f"({sum(len(ex) > 1 for ex in plan)} reduce-scatter, {sum(len(ex) == 1 for ex in plan)} reduce)", flush=True)
torch.cuda.synchronize()
for name, k, x, ws in units_np:
    owner = run.owner[name]
    if rank == owner:
        flips, worst, total = 0, 0, 0
        for w, (codes, s, z) in zip(ws, run.results[name]):
            want = O.gptq_quantize(w, x, "int4", "group", 128, True, mode="propagate")
            c = codes.cpu().numpy().view(np.int8).astype(np.int32)
            c = np.where(c > 7, c - 16, c)
            d = np.abs(c - np.asarray(want[0]).astype(np.int32))
            flips += int((d != 0).sum()); worst = max(worst, int(d.max())); total += d.size
        good = flips <= 1e-3 * total and worst <= 1 and run.factors[name].ok
        print(f"GPTQ unit {name} (K={k}) solved on rank {owner} of {world}: {flips} of {total} codes differ, max |diff| {worst}"
              f" -> {'ok' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
flag = torch.tensor([0 if ok else 1], device=device)
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
