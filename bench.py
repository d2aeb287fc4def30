#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload ("step") : one pass of RTN UINT4 asymmetric group_size=128 **with the MSE search**
(utils.py:140-239; mse=True discards clip_ratio exactly as the reference does) over a
Llama-3-8B-shaped set of linear weights — 32 layers x {q,o 4096x4096; k,v 4096x1024; gate,up
4096x14336; down 14336x4096} = 224 matrices, 6.979 G elements, 27.9 GB of float32 — written in
the MatMulNBits layout (packed nibbles + scales + packed zero points).  Synthetic randn*0.02
weights generated on the device.  N > 1: every rank quantizes its own model-sized set
(independent matrices, no data-path collective) -> weak scaling.

Metric: GB/s of float32 weight consumed (4*K*N bytes / device time).  `value` = inputs resident in
HBM; `e2e` = same work through the host-buffer API (pinned host weights -> H2D -> kernels -> D2H of
codes/scales/zero points inside the timed region).  `roofline` is for the dominant kernel
(rtn_group_mse4_kernel<128>): algorithmic bytes 4.535 B/element vs the measured HBM copy
bandwidth; the same pass without the MSE search (config 2a, clip_ratio 0.9 — the HBM-bound
kernel) is reported under `variants`, next to cfg1 / cfg3, the weight-sharding axis, GPTQ and the
figures measured through the plugin seam (`e2e_plugin`).

`--impl reference` times the CPU oracle port of the same path (NumPy, all host cores via a
process pool over column slices) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA3_8B_LAYER = [("q", 4096, 4096), ("k", 4096, 1024), ("v", 4096, 1024), ("o", 4096, 4096),
                   ("gate", 4096, 14336), ("up", 4096, 14336), ("down", 14336, 4096)]
N_LAYERS = 32
ALGO_BYTES_PER_ELT = 4 + 0.5 + 4 / 128 + 0.5 / 128
WORKLOAD = "cfg2: RTN uint4 asym g128 + MSE, Llama-3-8B-shaped linear set (224 matrices, 6.979 G elts), MatMulNBits layout"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--layers", type=int, default=N_LAYERS, help="model depth (debug only)")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-gptq", action="store_true")
    p.add_argument("--gptq-layers", type=int, default=0,
                   help="layers in the GPTQ variants (0 = every layer of the model: 32 Llama-3-8B / 26 Gemma-3-1B)")
    p.add_argument("--gptq-precision", default="bf16x3", choices=["tf32", "tf32x3", "bf16x3"])
    p.add_argument("--gptq-streams", type=int, default=8,
                   help="CUDA streams the independent GPTQ solves of a rank are spread over")
    return p.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
_CPU_WEIGHTS: list = []          # filled BEFORE the worker processes are forked: inherited, never pickled


def _cpu_slice(task):
    j, c0, c1, mse = task
    from oracle import np_oracle as O
    w = np.ascontiguousarray(_CPU_WEIGHTS[j][:, c0:c1])
    q, s, z = O.rtn_quantize(w, "uint4", "group", 128, False, False, 0.9, mse)
    b, bs, bz = O.matmul_nbits_layout(q, s, z, 128, 4)
    return b.shape


def cpu_reference_step(mse: bool, pool, cols_per_task: int = 256) -> float:
    """One bounded-sample step on the host: every matrix of the sample quantized with the oracle,
    split over column slices (groups never span columns, so the result is identical), largest
    slices first."""
    tasks = [(j, c, min(c + cols_per_task, w.shape[1]), mse) for j, w in enumerate(_CPU_WEIGHTS)
             for c in range(0, w.shape[1], cols_per_task)]
    tasks.sort(key=lambda t: -_CPU_WEIGHTS[t[0]].shape[0] * (t[2] - t[1]))
    t0 = time.perf_counter()
    if pool is None:
        for t in tasks:
            _cpu_slice(t)
    else:
        list(pool.map(_cpu_slice, tasks, chunksize=1))
    return time.perf_counter() - t0


def run_cpu_arm(args, sample=None):
    """The reference's path on the host cores: the oracle port over one Llama-3-8B-shaped layer
    (7 matrices, 218 M elements = 1/32 of the workload) per step, all cores."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    shapes = sample or [(k, n) for _, k, n in LLAMA3_8B_LAYER]
    _CPU_WEIGHTS.clear()
    _CPU_WEIGHTS.extend((rng.standard_normal(shp, dtype=np.float32) * np.float32(0.02)) for shp in shapes)
    nbytes = sum(w.nbytes for w in _CPU_WEIGHTS)
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork")) as pool:
        list(pool.map(_cpu_slice, [(0, 0, 2, False)] * cores))          # start the workers (imports) untimed
        for _ in range(1 if args.warmup > 0 else 0):
            cpu_reference_step(True, pool)
        times = [cpu_reference_step(True, pool) for _ in range(max(args.steps, 1))]
    t = statistics.mean(times)
    gbs = nbytes / t / 1e9
    sample_txt = f"one Llama-3-8B-shaped layer per step ({len(shapes)} matrices, {nbytes / 4 / 1e6:.0f} M elements = " \
                 f"1/{N_LAYERS} of the workload), 256-column slices over {cores} forked processes"
    _CPU_WEIGHTS.clear()
    return gbs, t, cores, sample_txt


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4)
                          if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_weights(torch, device, layers, rank):
    ws = []
    for layer in range(layers):
        for j, (_, k, n) in enumerate(LLAMA3_8B_LAYER):
            g = torch.Generator(device=device)
            g.manual_seed(rank * 100003 + layer * 7 + j)
            ws.append(torch.randn((k, n), generator=g, device=device, dtype=torch.float32) * 0.02)
    return ws


# ------------------------------------------------------------------------------------------------
# GPTQ variant (BASELINE config 5): INT4 g128 on Llama-3-8B-shaped layers, 128 x 2048 calibration
# tokens, calibration tokens split over the ranks (NCCL sum-reduce of every Hessian to the rank that
# owns its solve), solves sharded over the ranks.  Strong scaling: the sample is fixed.
# ------------------------------------------------------------------------------------------------
GPTQ_MODELS = {
    # name: (layers in the model, [(Hessian name, K, [(K, N) weights sharing that input])])
    "llama3_8b": (32, [("qkv", 4096, [(4096, 4096), (4096, 1024), (4096, 1024)]), ("o", 4096, [(4096, 4096)]),
                       ("gate_up", 4096, [(4096, 14336), (4096, 14336)]), ("down", 14336, [(14336, 4096)])]),
    "gemma3_1b": (26, [("qkv", 1152, [(1152, 1024), (1152, 256), (1152, 256)]), ("o", 1024, [(1024, 1152)]),
                       ("gate_up", 1152, [(1152, 6912), (1152, 6912)]), ("down", 6912, [(6912, 1152)])]),
}
GPTQ_SAMPLES, GPTQ_SEQ = 128, 2048


def gptq_cpu_reference(groups, model_layers):
    """The reference's CPU path for the same work, timed on a bounded sample and extrapolated
    (SURVEY.md §8d): `_accumulate_hessian` on 2 x 2048 tokens per distinct K (sgemm is linear in
    tokens), `_gptq` on one small weight per distinct K (the Python row loop is linear in K*N)."""
    from oracle import np_oracle as O
    rng = np.random.default_rng(0)
    tokens = GPTQ_SAMPLES * GPTQ_SEQ
    hess_s, solve_s, measured = 0.0, 0.0, 0.0
    # the Python row loop of `_gptq` (gptq.py:164-201): cost per row = a + b * N, measured once
    k0 = min(g[1] for g in groups)
    h0 = np.eye(k0, dtype=np.float32)
    loop = []
    for n_small in (64, 192):
        w = (rng.standard_normal((k0, n_small)) * 0.02).astype(np.float32)
        t0 = time.perf_counter()
        O.gptq(w, h0, "int4", "group", 128, True, mode="reference")
        loop.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    O.hinv_cholesky_upper(h0, 0.01)
    fact0 = time.perf_counter() - t0
    measured += sum(loop) + fact0
    per_row_col = max(loop[1] - loop[0], 0.0) / (128 * k0)
    per_row = max(loop[0] - fact0 - per_row_col * 64 * k0, 0.0) / k0
    per_k = {}
    for _, k, shapes in groups:
        if k not in per_k:
            x = rng.standard_normal((2, GPTQ_SEQ, k)).astype(np.float32)
            t0 = time.perf_counter()
            h, _ = O.accumulate_hessian(x, np.zeros((k, k), np.float32), 0)
            th = time.perf_counter() - t0
            t0 = time.perf_counter()
            O.hinv_cholesky_upper(h + np.eye(k, dtype=np.float32) * np.float32(0.1 * np.trace(h) / k), 0.01)
            tf = time.perf_counter() - t0          # cholesky + inv + cholesky: once per WEIGHT in the reference
            per_k[k] = (th * tokens / (2 * GPTQ_SEQ), tf)
            measured += th + tf
        th_full, tf = per_k[k]
        # the reference accumulates one Hessian PER NODE (calibrate.py:296-307), not per shared input
        hess_s += th_full * len(shapes)
        solve_s += sum(tf + k * (per_row + per_row_col * n) for _, n in shapes)
    return {"kind": "port", "cores": os.cpu_count(), "s_per_layer_extrapolated": hess_s + solve_s,
            "s_per_model_extrapolated": (hess_s + solve_s) * model_layers,
            "hessian_s_per_layer": hess_s, "solve_s_per_layer": solve_s, "measured_seconds": measured,
            "sample": "np.matmul Hessian on 2x2048 tokens per distinct K scaled to 128x2048; LAPACK "
                      "cholesky+inv+cholesky once per distinct K; the Python row loop of _gptq timed on 64 and "
                      "192 columns at the smallest K and scaled as K*(a + b*N); one Hessian and one "
                      "factorization per node, as the reference does"}


def run_gptq_variant(args, torch, dist, device, world, rank, model="llama3_8b", layers=None, cpu_ref=False):
    model_layers, GPTQ_GROUPS = GPTQ_MODELS[model]
    from onnx_quantize_b200.parallel.gptq_pipeline import GptqPipeline, GptqSpec, GptqUnit

    layers = layers or args.gptq_layers or model_layers
    tokens = GPTQ_SAMPLES * GPTQ_SEQ
    t_local = tokens // world
    gen = torch.Generator(device=device)
    gen.manual_seed(4242 + rank)
    xs = {k: torch.randn((t_local, k), generator=gen, device=device, dtype=torch.float32)
          for k in sorted({g[1] for g in GPTQ_GROUPS})}
    ws = {}
    for _, _, shapes in GPTQ_GROUPS:
        for shp in shapes:
            if shp not in ws:
                ws[shp] = torch.randn(shp, generator=gen, device=device, dtype=torch.float32) * 0.02
    units = [(l, gi) for l in range(layers) for gi in range(len(GPTQ_GROUPS))]
    # the layers are identical in shape: one token tensor per K and one weight tensor per shape stand
    # for all of them (every unit still gets its own Hessian, factor and results)
    punits = [GptqUnit(f"l{l}.{GPTQ_GROUPS[gi][0]}", GPTQ_GROUPS[gi][1], [ws[shp] for shp in GPTQ_GROUPS[gi][2]],
                       xs[GPTQ_GROUPS[gi][1]], GPTQ_SAMPLES) for l, gi in units]
    spec = GptqSpec("int4", "group", 128, True, False, 1.0, False, 128, 0.01, False, "propagate", args.gptq_precision)
    pipe = GptqPipeline(args.gptq_streams, device)

    run = pipe.run(punits, spec)                           # warm-up (workspaces, attributes, NCCL channels)
    del run
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    run = pipe.run(punits, spec)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1, e2, e3 = run.start, run.hessians_done, run.reduces_done, run.end
    t = torch.tensor([e0.elapsed_time(e3), e0.elapsed_time(e1), e1.elapsed_time(e2), e1.elapsed_time(e3)],
                     device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, hess_ms, red_ms, solve_ms = (float(v) for v in t)
    n_failed = sum(0 if f.ok else 1 for f in run.factors.values())
    del run
    flops = sum(2.0 * tokens * GPTQ_GROUPS[gi][1] ** 2 for _, gi in units)     # de-duplicated by input

    def executed(k):   # the kernel only computes the 128x256 tiles that touch the upper triangle
        n_ib, n_jb = -(-k // 128), -(-k // 256)
        tiles = sum(max(n_jb - (ib * 128) // 256, 0) for ib in range(n_ib))
        return 2.0 * tokens * tiles * 128 * 256 * (1 if args.gptq_precision == "tf32" else 3)

    mma_flops = sum(executed(GPTQ_GROUPS[gi][1]) for _, gi in units)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    bf16 = float(peaks.get("bf16_tflops_sustained", 1393.9))
    tf = mma_flops / world / (hess_ms * 1e-3) / 1e12       # per GPU: every rank contracts tokens/world
    is_bf16 = args.gptq_precision == "bf16x3"
    mma_peak = bf16 if is_bf16 else bf16 / 2
    out = {
        "workload": f"GPTQ int4 sym g128 (block 128, percdamp 0.01, mode=propagate) on {layers} of {model_layers} "
                    f"{model}-shaped layers ({len(units)} Hessians, {7 * layers} weights), "
                    f"{GPTQ_SAMPLES}x{GPTQ_SEQ} calibration tokens split over {world} rank(s)",
        "precision": args.gptq_precision, "solve_streams": args.gptq_streams, "scaling": "strong", "n_gpus": world,
        "s_per_step": total_ms * 1e-3, "s_per_model_extrapolated": total_ms * 1e-3 * model_layers / layers,
        "extrapolation": f"x{model_layers / layers:g}: the {model_layers} layers are identical in shape",
        "hessian_s": hess_ms * 1e-3, "solve_s": solve_ms * 1e-3,
        "hessian_reduce_s_under_solves": red_ms * 1e-3,
        "phases": "hessian_s = every unit's Hessian on this rank's tokens; then the partial Hessians reach the owners "
                  "(in-place NCCL reduce-scatters over stacked buffers, one unit per owner each; plain reduces for leftovers) on "
                  "a communication stream UNDERNEATH the solves (each solve waits for its own unit's reduce only): "
                  "solve_s spans both, hessian_reduce_s_under_solves is when the last exchange step finished",
        "factorizations_failed": n_failed,
        "hessian_tflops_algorithmic": flops / (hess_ms * 1e-3) / 1e12,
        "roofline": {"bound": "tensor", "kernel": "hessian_bf16x3_kernel" if is_bf16 else "hessian_kernel",
                     "achieved": tf, "unit": "TFLOP/s", "peak": mma_peak, "frac": tf / mma_peak,
                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kind::f16, bf16 inputs); achieved = bf16 "
                                     "MMA flops actually issued per GPU (upper-triangle tiles only, the 3 products of "
                                     "the BF16x3 split) / Hessian time, split pre-pass included") if is_bf16 else
                                    ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (kind::tf32 runs at half the "
                                     "bf16 rate); achieved = tf32 MMA flops actually issued per GPU (upper-triangle "
                                     "tiles only, x3 products in 3xTF32 mode) / Hessian time"),
                     "flops": flops},
    }
    del xs, ws, punits
    pipe.release()
    from onnx_quantize_b200 import device_api as D
    D.dev.release_workspaces()
    torch.cuda.empty_cache()
    if cpu_ref and rank == 0:
        out["cpu_reference"] = gptq_cpu_reference(GPTQ_GROUPS, model_layers)
    return out


# ------------------------------------------------------------------------------------------------
# cfg1 / cfg3 variants: the other two HBM-bound kernels of the path
# ------------------------------------------------------------------------------------------------
def run_small_variants(torch, device, peak):
    from onnx_quantize_b200 import _device as dev
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

    out = {}
    # cfg1: RTN int8 symmetric per-tensor, two 4096x4096 weights (64 MiB each: L2-resident second pass)
    gen = torch.Generator(device=device)
    gen.manual_seed(0)
    ws = [torch.randn((4096, 4096), generator=gen, device=device) * 0.02 for _ in range(2)]
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=device)   # 512 MB > L2 (126 MB)

    def time_cold_ms(fn, iters=6):
        """Median device time of fn with the L2 emptied before every iteration by READING the flush buffer
        (clean lines: the timed kernels do not pay for somebody else's write-backs)."""
        ts = []
        for _ in range(iters):
            flush.sum()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    plan1 = D.RtnBatchPlan(ws, QuantType.QInt8, "tensor", -1, True, False, 1.0, False)   # the model's two weights

    def cfg1():
        plan1.run()

    with dev.inputs_resident():          # the weights were resident before anything was launched
        ms = time_ms(cfg1)
        ms_cold = time_cold_ms(cfg1)
    elts = sum(w.numel() for w in ws)
    out["cfg1_int8_sym_tensor"] = {
        "workload": "cfg1: RTN int8 symmetric per-tensor, 2 x (4096x4096) f32 (codes one byte per element)",
        "ms_per_step": ms, "value": 4 * elts / (ms * 1e-3) / 1e9, "unit": "GB/s",
        "cache": "ms_per_step: iterations back to back — the two 64 MiB weights (128 MiB, L2 is 126 MB) partly "
                 "survive in L2 from one iteration to the next; cold: L2 emptied by reading 512 MB before every iteration",
        "cold": {"ms_per_step": ms_cold, "value": 4 * elts / (ms_cold * 1e-3) / 1e9, "unit": "GB/s",
                 "roofline_frac": 5.0 * elts / (ms_cold * 1e-3) / 1e9 / peak},
        "roofline": {"bound": "hbm", "kernel": "minmax_partials_kernel + quantize_flat_kernel",
                     "achieved": 5.0 * elts / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": 5.0 * elts / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_element": 5.0,
                     "note": "two passes over W (global min/max must precede any code); the second pass "
                             "re-reads the 64 MiB weight from L2; programmatic dependent launches overlap "
                             "the passes of consecutive weights"}}
    # cfg3: static-calibration min/max of 100 x 512 x 4096 activations in 10 batches, then A3
    acts = [torch.randn((10, 512, 4096), generator=gen, device=device) for _ in range(10)]
    slots = torch.empty((10, D.minmax_partials_stride(), 2), dtype=torch.float32, device=device)
    counts = torch.zeros((10,), dtype=torch.int32, device=device)
    state = torch.zeros((2,), dtype=torch.float32, device=device)
    valid = torch.zeros((1,), dtype=torch.int32, device=device)
    rng2 = torch.zeros((2,), dtype=torch.float32, device=device)

    batches = [(x.reshape(-1), slots[i], counts[i:i + 1]) for i, x in enumerate(acts)]

    def cfg3():   # what MinMaxCalibrator.collect x 10 + compute_range launch: 10 reductions, 1 fold+merge, A3
        valid.zero_()
        for x, slot, cnt in batches:
            D.minmax_partials(x, slot, cnt)
        D.minmax_fold_merge(state, valid, slots, counts, 10, 0.0, None, rng2)   # range incl. zero (minmax.py:84-87)
        D.qparams(rng2[0:1], rng2[1:2], QuantType.QUInt8)

    with dev.inputs_resident():          # so were the calibration batches
        ms = time_ms(cfg3, iters=5)
    elts = sum(x.numel() for x in acts)
    out["cfg3_activation_minmax"] = {
        "workload": "cfg3: MinMax calibration of one tensor, 100x512x4096 f32 activations in 10 batches + uint8 scale/zp",
        "ms_per_step": ms, "value": 4 * elts / (ms * 1e-3) / 1e9, "unit": "GB/s",
        "roofline": {"bound": "hbm", "kernel": "minmax_partials_kernel", "achieved": 4.0 * elts / (ms * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": 4.0 * elts / (ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_element": 4.0}}
    # N3: AWQ scale + clip search (30 candidates) for one q_proj-shaped weight, statistics given
    from onnx_quantize_b200 import awq as A
    k = 4096
    st = A.AwqStatistics(k)
    xa = torch.randn((8192, k), generator=gen, device=device) * (torch.rand((k,), generator=gen, device=device) * 2 + 0.2)
    ms_stats = time_ms(lambda: st.add(xa), iters=3, warm=1)
    wq = ws[0]
    ms = time_ms(lambda: A.awq_search(wq, st, QuantType.QUInt4, "group", 128, False, False, clip_search=True),
                 iters=2, warm=1)
    flops = 30 * 2.0 * k * k * wq.shape[1] * 3          # 30 Gram products, 3 bf16 products each (BF16x3)
    out["awq_uint4_g128_q_proj"] = {
        "workload": "AWQ scale grid (20) + clip grid (10) for one 4096x4096 weight, uint4 g128, through the Gram "
                    "matrix (pre_passes/awq.py:121-184, :207-254); statistics accumulation timed separately",
        "ms_per_step": ms, "stats_ms_per_8192_tokens": ms_stats,
        "roofline": {"bound": "tensor", "kernel": "dense_bf16x3_kernel", "achieved": flops / (ms * 1e-3) / 1e12,
                     "unit": "TFLOP/s", "peak": 1393.9, "frac": flops / (ms * 1e-3) / 1e12 / 1393.9,
                     "note": "bf16 MMA flops of the 30 (K,K)x(K,N) products / whole search time (the search also "
                             "runs 30 RTN parameter passes, residual, plane-split and dot kernels — about half "
                             "of the time); peak = measured sustained bf16"}}
    del flush
    return out


# ------------------------------------------------------------------------------------------------
# cfg3 end to end: static W8A8 of a synthetic Gemm MLP, calibration forward on the device, batches
# sharded over the ranks, activation ranges combined with one NCCL min/max all-reduce
# ------------------------------------------------------------------------------------------------
def run_cfg3_mlp(torch, dist, device, world, rank):
    import onnx_quantize_b200 as q
    from onnx_quantize_b200 import device_api as D
    from onnx_quantize_b200.calibrate_mlp import DenseLayer, calibrate_mlp

    gen = torch.Generator(device=device)
    gen.manual_seed(7)                                   # identical model and data on every rank
    k = 4096
    layers = [DenseLayer(f"fc{i}", torch.randn((k, k), generator=gen, device=device) * 0.02,
                         torch.randn((k,), generator=gen, device=device) * 0.1, "relu" if i == 0 else None)
              for i in range(2)]
    data = torch.randn((100, 512, k), generator=gen, device=device)
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int8", symmetric=True, strategy="channel"),
                    input_activations=q.QActivationArgs(dtype="uint8", is_static=True),
                    output_activations=q.QActivationArgs(dtype="uint8", is_static=True),
                    calibration_params=q.CalibrationParams(num_samples=100, batch_size=10))

    def step():
        acts = calibrate_mlp(layers, data, cfg)
        outs = []
        for l in layers:                                  # weights: RTN int8 per-channel; bias: int32 (A9)
            _, ws, _ = wq = D.rtn_quantize(l.weight, q.QuantType.QInt8, "channel", -1, True)
            outs.append((wq, D.quantize_bias(l.bias, float(acts[l.name]["input_scale"]), ws)))
        return acts, outs

    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    step()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    del data, layers
    torch.cuda.empty_cache()
    return {"workload": "cfg3: static W8A8 of a 2-layer 4096-wide Gemm MLP, 100x512x4096 f32 calibration samples in 10 "
                        f"batches sharded over {world} rank(s): on-device forward (tcgen05 BF16x3), min/max of 4 "
                        "activation tensors, NCCL min/max all-reduce, uint8 activation parameters, int8 per-channel "
                        "weights, int32 biases", "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
            "calibration_bytes": 100 * 512 * k * 4, "value": 100 * 512 * k * 4 / (ms * 1e-3) / 1e9, "unit": "GB/s",
            "roofline": {"bound": "tensor", "kernel": "dense_bf16x3_kernel", "unit": "TFLOP/s",
                         "achieved": 2 * 2.0 * 51200 * k * k * 3 / world / (ms * 1e-3) / 1e12, "peak": 1393.9,
                         "frac": 2 * 2.0 * 51200 * k * k * 3 / world / (ms * 1e-3) / 1e12 / 1393.9,
                         "note": "bf16 MMA flops (3 products of the BF16x3 split) of the two forward products per "
                                 "GPU / whole step (splits, min/max passes, weight and bias quantization included); "
                                 "peak = measured sustained bf16"}}


# ------------------------------------------------------------------------------------------------
# The reference-facing seam, measured as the reference drives it (qrules/_common.py:133): one
# `qconfig.weights.algorithm.quantize_weights(w, qconfig)` call per initializer with the PAGEABLE
# NumPy array `w.const_value.numpy()` delivers, NumPy results back, single-threaded caller.
# ------------------------------------------------------------------------------------------------
class _Const:
    def __init__(self, a):
        self._a = a

    def numpy(self):
        return self._a


class _Value:                      # ir.Value: .name, .const_value.numpy(), .producer()
    def __init__(self, name, a=None, node=None):
        self.name, self.const_value, self._node = name, _Const(a), node

    def producer(self):
        return self._node


class _Node:                       # ir.Node: .meta
    def __init__(self, meta):
        self.meta = meta


def _pageable_layer(rank):
    rng = np.random.default_rng(500 + rank)
    return [rng.standard_normal((k, n), dtype=np.float32) * np.float32(0.02) for _, k, n in LLAMA3_8B_LAYER]


def run_plugin_e2e(args, torch):
    """cfg2 through the plugin seam, two ways: (a) plugin calls alone, one weight at a time; (b) the
    pre-pass the package offers for exactly this (parallel.shard.quantize_weights_sharded publishes
    reference-shaped triples, the plugin calls that follow are look-ups) — timed INCLUDING the 224
    plugin calls.  Inputs: 7 distinct pageable float32 arrays (one Llama-3-8B-shaped layer) visited
    once per layer of the model; every call returns fresh NumPy arrays, all 224 results are kept
    alive until the step ends, as the rewriter keeps its initializers."""
    import onnx_quantize_b200 as q
    from onnx_quantize_b200.parallel import prequantized
    from onnx_quantize_b200.parallel.shard import quantize_weights_sharded
    from onnx_quantize_b200.pipeline import RtnSpec

    layer = _pageable_layer(0)
    names = [f"model.layers.{l}.{nm}.weight" for l in range(args.layers) for nm, _, _ in LLAMA3_8B_LAYER]
    arrays = [layer[j] for _ in range(args.layers) for j in range(len(layer))]
    in_bytes = sum(a.nbytes for a in arrays)
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="uint4", group_size=128, symmetric=False, mse=True, clip_ratio=0.9))
    plugin = cfg.weights.algorithm

    def direct():
        return [plugin.quantize_weights(_Value(nm, a), cfg) for nm, a in zip(names, arrays)]

    def with_prepass():
        with prequantized.scope():
            quantize_weights_sharded(dict(zip(names, arrays)), RtnSpec.from_weight_args(cfg.weights))
            return [plugin.quantize_weights(_Value(nm, a), cfg) for nm, a in zip(names, arrays)]

    out = {}
    for key, fn, api in (("plugin_calls", direct, "RTNConfig.quantize_weights(w, qconfig) per initializer, nothing else"),
                         ("prepass_then_plugin_calls", with_prepass,
                          "parallel.shard.quantize_weights_sharded (bulk pipeline, results published) followed by the same "
                          "224 RTNConfig.quantize_weights calls (look-ups keyed by name + request digest + weight fingerprint)")):
        res = fn()                                                 # warm-up (staging, pinned pool, workspaces)
        d2h = sum(sum(np.asarray(x).nbytes for x in r) for r in res)
        del res
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        del res
        out[key] = {"value": in_bytes / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": d2h, "api": api}
    out["note"] = ("pageable NumPy in, NumPy (reference dtypes/shapes, (K,N) codes one byte per element) out; the host "
                   "side is bound by memcpy into pinned staging (35-60 GB/s on these hosts) and by the first-touch page "
                   "faults of the 7.0 GB of fresh result arrays (17-37 GB/s), see DESIGN.md §6")
    return out


def run_gptq_plugin_e2e(args, torch, device):
    """One Llama-3-8B-shaped layer through `GPTQConfig.quantize_weights` exactly as the rewriter calls
    it: host float32 weights, host calibration activations in node.meta["input"] — 128 x 2048 tokens per
    distinct layer input, the SAME array object for q/k/v and for gate/up (calibrate.py:296-307) —
    host codes/scales/zero points back.  Activations are generated on the device and copied to
    pageable host memory before the timed region."""
    import onnx_quantize_b200 as q
    from onnx_quantize_b200 import _lib
    from onnx_quantize_b200.core._algorithms.gptq import calibration_cache

    lib = _lib.load()
    gen = torch.Generator(device=device)
    gen.manual_seed(99)

    def host_tokens(k):
        x = torch.empty((GPTQ_SAMPLES, GPTQ_SEQ, k), dtype=torch.float32)
        for s0 in range(0, GPTQ_SAMPLES, 16):
            x[s0:s0 + 16] = torch.randn((16, GPTQ_SEQ, k), generator=gen, device=device).cpu()
        return x.numpy()

    inputs = {"qkv": host_tokens(4096), "o": host_tokens(4096), "gate_up": host_tokens(4096), "down": host_tokens(14336)}
    use = {"q": "qkv", "k": "qkv", "v": "qkv", "o": "o", "gate": "gate_up", "up": "gate_up", "down": "down"}
    layer = _pageable_layer(1)
    cfg = q.QConfig(weights=q.QWeightArgs(dtype="int4", group_size=128, symmetric=True,
                                          algorithm=q.GPTQConfig(mode="propagate", precision=args.gptq_precision)))
    plugin = cfg.weights.algorithm

    def step():
        calibration_cache.clear()
        return [plugin.quantize_weights(_Value(nm, w), cfg, out=_Value("y", node=_Node({"input": inputs[use[nm]]})))
                for (nm, _, _), w in zip(LLAMA3_8B_LAYER, layer)]

    step()
    torch.cuda.synchronize()
    launches0 = lib.b200q_launch_count()
    m0 = calibration_cache.misses
    t0 = time.perf_counter()
    res = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    h2d = sum(x.nbytes for x in inputs.values()) + sum(w.nbytes for w in layer)
    d2h = sum(sum(np.asarray(x).nbytes for x in r) for r in res)
    calibration_cache.clear()
    return {"workload": "GPTQ int4 sym g128 of ONE Llama-3-8B-shaped layer through GPTQConfig.quantize_weights: 7 plugin "
                        "calls, host weights + host activations (4 distinct inputs of 128x2048 tokens, q/k/v and gate/up "
                        "share theirs) in, host results out",
            "s_per_layer": dt, "s_per_model_extrapolated": dt * 32, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "h2d_gbs": h2d / dt / 1e9, "hessians_computed": calibration_cache.misses - m0,
            "gpu_launches": int(lib.b200q_launch_count() - launches0),
            "api": "GPTQConfig.quantize_weights(w, qconfig, out=out) with node.meta['input'] (pageable NumPy)",
            "note": "28.3 GB of calibration activations cross PCIe once per distinct input (the reference keeps them in "
                    "host RAM); the step is bound by that upload"}


def run_cfg2_strong(args, torch, dist, device, world, rank, weights):
    """The weight-sharding axis (SURVEY.md §8e first bullet): ONE Llama-3-8B-shaped set quantized by all
    ranks together.  (a) device-resident shards, one batched launch per rank, CUDA events, max over
    ranks; (b) end to end through parallel.shard.quantize_weights_sharded: every rank uploads its
    share from (pageable) host memory, packed results travel to rank 0 as tensors over NCCL and are
    copied to the host once — wall clock between barriers, max over ranks."""
    from onnx_quantize_b200 import device_api as D
    from onnx_quantize_b200.core._dtypes import QuantType
    from onnx_quantize_b200.parallel.shard import assign_units, quantize_weights_sharded
    from onnx_quantize_b200.pipeline import RtnSpec

    layer = _pageable_layer(0)                               # identical on every rank
    names = [f"model.layers.{l:02d}.{nm}.weight" for l in range(args.layers) for nm, _, _ in LLAMA3_8B_LAYER]
    named = {nm: layer[i % len(layer)] for i, nm in enumerate(names)}   # 7 distinct arrays, visited once per layer
    costs = [w.numel() for w in weights]                     # the rank's resident set stands for the one model
    mine = assign_units(costs, world)[rank]
    shard = [weights[i] for i in mine]
    plan = D.RtnBatchPlan(shard, QuantType.QUInt4, "group", 128, False, False, 0.9, True, layout="matmul_nbits")

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        plan.run()
    sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        plan.run()
    b.record()
    sync_all()
    t = torch.tensor([a.elapsed_time(b) / 3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t[0])
    del plan, shard
    spec = RtnSpec(QuantType.QUInt4, "group", 128, False, False, 0.9, True, "kn")
    quantize_weights_sharded(named, spec, publish=False)     # warm-up
    sync_all()
    t0 = time.perf_counter()
    res = quantize_weights_sharded(named, spec, publish=False)
    sync_all()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
    d2h = sum(sum(np.asarray(x).nbytes for x in r) for r in res.values()) if res else 0
    del res
    torch.cuda.empty_cache()
    in_bytes = 4 * sum(costs)
    return {"workload": "cfg2 (RTN uint4 asym g128 + MSE) on ONE Llama-3-8B-shaped set sharded over the ranks by weight "
                        "matrix (largest first), no data-path collective", "scaling": "strong", "n_gpus": world,
            "device": {"ms_per_step": dev_ms, "value": in_bytes / (dev_ms * 1e-3) / 1e9, "unit": "GB/s",
                       "timing": "CUDA events around one batched launch over the rank's shard, max over ranks"},
            "e2e": {"ms_per_step": dt * 1e3, "value": in_bytes / dt / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_on_rank0": d2h,
                    "api": "parallel.shard.quantize_weights_sharded(named pageable weights, RtnSpec(layout='kn')): "
                           "results gathered on rank 0 as packed tensors over NCCL, one D2H there",
                    "timing": "wall clock between barriers, max over ranks"}}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from onnx_quantize_b200 import _lib
    from onnx_quantize_b200 import device_api as D
    from onnx_quantize_b200._device import bind_host_to_gpu_numa_node
    from onnx_quantize_b200.core._dtypes import QuantType
    from onnx_quantize_b200.pipeline import RtnSpec, quantize_weights_bulk, result_bytes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    numa_cpus = bind_host_to_gpu_numa_node(local_rank)      # before any pinned allocation: H2D from the local node

    weights = make_weights(torch, device, args.layers, rank)
    elts = sum(w.numel() for w in weights)
    in_bytes = 4 * elts
    qt = QuantType.QUInt4

    plans = {mse: D.RtnBatchPlan(weights, qt, "group", 128, False, False, 0.9, mse, layout="matmul_nbits")
             for mse in (False, True)}

    def step(mse: bool):
        return plans[mse].run()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mse: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(mse)
        sync_all()
        launches0 = lib.b200q_launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0 = time.perf_counter()
        for a, b in evs:
            a.record()
            step(mse)
            b.record()
        sync_all()
        wall = time.perf_counter() - t0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        total_ms = max(dev_ms, 0.0)
        if world > 1:
            t = torch.tensor([total_ms, wall * 1e3], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, wall = float(t[0]), float(t[1]) / 1e3
        return total_ms / steps, wall / steps, lib.b200q_launch_count() - launches0

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_mse, wall_mse, launches = timed(True, args.steps, max(args.warmup, 3))
    if sampler:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    ms_plain, _, _ = timed(False, args.steps, max(args.warmup, 3))

    # ---- e2e: host buffers through the bulk pipeline (H2D + kernels + D2H in the timed region)
    e2e = None
    if not args.no_e2e:
        gens = []
        for j, (_, k, n) in enumerate(LLAMA3_8B_LAYER):      # one layer of distinct pinned weights,
            g = torch.Generator()                             # revisited for every layer of the model
            g.manual_seed(1000 + rank * 10 + j)
            gens.append((torch.randn((k, n), generator=g, dtype=torch.float32) * 0.02).pin_memory())
        host_set = [gens[j] for _ in range(args.layers) for j in range(len(LLAMA3_8B_LAYER))]
        spec = RtnSpec(qt, "group", 128, False, False, 0.9, True, "matmul_nbits")
        res = quantize_weights_bulk(host_set[:len(gens)], spec)        # warm-up (pinned pool, workspace)
        d2h_bytes = result_bytes(res) * args.layers
        for _ in range(1):
            quantize_weights_bulk(host_set, spec)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            quantize_weights_bulk(host_set, spec)
        sync_all()
        e2e_s = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t[0])
        # what the link itself delivers on this box: plain pinned -> device copies of the same buffers
        dst = [torch.empty_like(g_, device=device) for g_ in gens]
        for g_, d_ in zip(gens, dst):
            d_.copy_(g_, non_blocking=True)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(4):
            for g_, d_ in zip(gens, dst):
                d_.copy_(g_, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = 4 * sum(g_.numel() * 4 for g_ in gens) / (time.perf_counter() - t0) / 1e9
        del dst
        e2e = {"value": world * in_bytes / e2e_s / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h_bytes,
               "ms_per_step": e2e_s * 1e3,
               "pcie_h2d_gbs_measured_per_gpu": h2d_gbs,
               "frac_of_pcie_h2d": (in_bytes / e2e_s / 1e9) / h2d_gbs,
               "api": "onnx_quantize_b200.pipeline.quantize_weights_bulk (PINNED host weights in, pinned-backed host results "
                      "out; the bulk pipeline the multi-GPU pre-pass runs on each rank) — NOT the per-initializer plugin "
                      "call: that is e2e_plugin",
               "host_inputs": "7 distinct pinned matrices (one Llama-3-8B-shaped layer), revisited once per layer of the "
                              "model: every step moves all 27.9 GB across PCIe"}

    strong = run_cfg2_strong(args, torch, dist, device, world, rank, weights)
    plugin = run_plugin_e2e(args, torch) if (world == 1 and not args.no_e2e) else None
    small = run_small_variants(torch, device, measured_peaks()[0]) if rank == 0 else None
    cfg3 = run_cfg3_mlp(torch, dist, device, world, rank)
    if small is not None:
        small["cfg3_static_w8a8_mlp"] = cfg3
    gptq = None
    if not args.no_gptq:
        del plans, weights
        D.dev.release_workspaces()
        torch.cuda.empty_cache()
        want_cpu = not args.no_cpu_baseline and world == 1
        gptq = {}
        if world == 1 and not args.no_e2e:
            gptq["gptq_int4_g128_llama3_8b_layer_e2e_plugin"] = run_gptq_plugin_e2e(args, torch, device)
        gptq = {**gptq,
                "gptq_int4_g128_llama3_8b": run_gptq_variant(args, torch, dist, device, world, rank, "llama3_8b",
                                                             cpu_ref=want_cpu),
                "gptq_int4_g128_gemma3_1b": run_gptq_variant(args, torch, dist, device, world, rank, "gemma3_1b")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    algo_bytes = ALGO_BYTES_PER_ELT * elts
    achieved = algo_bytes / (ms_mse * 1e-3) / 1e9
    achieved_plain = algo_bytes / (ms_plain * 1e-3) / 1e9
    line = {
        "metric": "RTN+MSE weight GB/s (fp32 weight bytes consumed / device time)",
        "value": world * in_bytes / (ms_mse * 1e-3) / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_mse,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "b200",
        "config": {"workload": WORKLOAD if args.layers == N_LAYERS else WORKLOAD + f" [{args.layers} layers]"},
        "config_detail": {"elements_per_gpu": elts, "input_bytes_per_gpu": in_bytes,
                          "parallelism": f"{world} rank(s), one model-sized set each, no collective (replicas: the "
                                         "strong-scaling figures are variants.cfg2_strong_weight_sharding and the GPTQ variants)",
                          "cache": "inputs (27.9 GB) larger than L2; no flush needed",
                          "host_cores_bound_to_gpu_numa_node": len(numa_cpus) if numa_cpus else None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": 243.74e6,
                     "traffic_note": "dram__bytes_read+write of ONE 4096x14336 launch of the kernel (ncu --set full, "
                                     "profiles/r1_prof_rtn_final_details.txt; the kernel is unchanged since): 235.2 MB read = the f32 weight once, "
                                     "8.5 MB written (the rest of the 29 MB result is still in L2 at kernel end); "
                                     "algorithmic bytes of that launch: 266.3 MB",
                     "peak_source": peak_src,
                     "kernel": "rtn_group_mse4_kernel<128>",
                     "compute_pipes_ncu": {"xu_mufu_busy_pct": 73, "issue_active_pct": 52,
                                           "source": "profiles/r1_prof_rtn_final_details.txt (4096x14336 launch, 0.82 ms)"},
                     # the bound that actually applies: 2 MUFU operations (lg2, ex2) per candidate-element,
                     # 20 candidates, 16 MUFU lanes per SM and clock
                     "mufu_roofline": {"ops_per_step": 2.0 * 20 * elts,
                                       "achieved_gops": 2.0 * 20 * elts / (ms_mse * 1e-3) / 1e9,
                                       "peak_gops": 148 * 16 * 1.965,
                                       "frac": 2.0 * 20 * elts / (ms_mse * 1e-3) / 1e9 / (148 * 16 * 1.965),
                                       "peak_source": "148 SMs x 16 MUFU lanes x 1.965 GHz (nominal)"},
                     "note": "the MSE search is bound by the MUFU pipe and instruction issue (20 candidates x "
                             "(quantize, dequantize, |d|^2.4) per element: ncu XU 73 %, issue 52 %), not by HBM; "
                             "the HBM-bound kernel of the same path is variants.cfg2a_no_mse"},
        "variants": {"cfg2a_no_mse_clip0.9": {
            "ms_per_step": ms_plain, "value": world * in_bytes / (ms_plain * 1e-3) / 1e9, "unit": "GB/s",
            "roofline": {"bound": "hbm", "kernel": "rtn_group_nbits4_ring_kernel<128>", "achieved": achieved_plain,
                         "peak": peak, "unit": "GB/s", "frac": achieved_plain / peak,
                         "launches_per_step": 2,
                         "traffic": 17.922e9,
                         "traffic_note": "the step is two launches of the persistent ring kernel (128 + 96 of the 224 jobs; "
                                         "the tensor maps live in the kernel parameters).  ncu --set full on the FIRST launch "
                                         "(profiles/r2_ring_stream_raw.csv): dram__bytes_read 15.806 GB + dram__bytes_write "
                                         "2.116 GB = 17.92 GB against 17.90 GB of algorithmic bytes for its 128 matrices "
                                         "(3.947 G elements x 4.535 B): no re-reads; DRAM 78.5 % of ncu's peak"}}},
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall_mse * 1e3,
        "clocks": sampler.summary() if sampler else None,
    }
    line["variants"]["cfg2_strong_weight_sharding"] = strong
    if plugin:
        line["e2e_plugin"] = plugin
    if small:
        line["variants"].update(small)
    if gptq:
        line["variants"].update(gptq)
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        gbs, t, cores, sample = run_cpu_arm(argparse.Namespace(steps=1, warmup=0))
        line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                                "sample": sample, "seconds": t}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        gbs, t, cores, sample = run_cpu_arm(args)
        print(json.dumps({
            "metric": "RTN+MSE weight GB/s (fp32 weight bytes consumed / device time)",
            "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD},
            "config_detail": {"sample": sample},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }), flush=True)
        return
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
