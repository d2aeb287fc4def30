"""Measure (not assert) GPTQ parity at the shapes bench.py times: Hessian error vs float64, factor
residuals, and the full propagate chain against the oracle at 4096x4096 for every Hessian
precision mode and two kinds of calibration data.  Writes gpurun_out/explore_gptq_parity.json;
the thresholds of tests/test_gptq_bench_shapes_gpu.py come from these numbers."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate
from oracle import np_oracle as O

out = {}
dev = torch.device("cuda:0")
what = set(sys.argv[1:]) or {"hessian", "hinv", "chain", "mse"}


def hess_err(t, k, precision, seed=0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randn((t, k), device=dev, generator=g) * (torch.rand((k,), device=dev, generator=g) * 2.7 + 0.3)
    h = torch.zeros((k, k), device=dev)
    hessian_accumulate(x, h, alpha=2.0 / 128, beta=0.0, precision=precision)
    want = torch.zeros((k, k), device=dev, dtype=torch.float64)
    for c in range(0, t, 8192):
        xc = x[c:c + 8192].double()
        want += xc.T @ xc
    want *= 2.0 / 128
    err = ((h.double() - want).abs().max() / want.abs().max()).item()
    # relative to each entry's own scale sqrt(h_ii h_jj): what matters for the factor
    d = want.diagonal().sqrt()
    err_rel = ((h.double() - want).abs() / (d[:, None] * d[None, :])).max().item()
    return err, err_rel


if "hessian" in what:
    for t, k in ((65536, 4096), (16384, 14336), (262144, 4096)):
        for p in ("bf16x3", "tf32x3"):
            t0 = time.time()
            e = hess_err(t, k, p)
            out[f"hessian|{t}|{k}|{p}"] = e
            print("hessian", t, k, p, e, f"{time.time() - t0:.1f}s", flush=True)

if "hinv" in what:
    for k in (4096, 14336):
        g = torch.Generator(device=dev)
        g.manual_seed(k)
        t = 32768
        x = torch.randn((t, k), device=dev, generator=g) * (torch.rand((k,), device=dev, generator=g) * 2.7 + 0.3)
        x[:, 1:] += 0.5 * x[:, :-1]
        h64 = torch.zeros((k, k), device=dev, dtype=torch.float64)
        for c in range(0, t, 4096):
            xc = x[c:c + 4096].double()
            h64 += xc.T @ xc
        h64 *= 2.0 / 128
        h = h64.float()
        del x
        for p in ("bf16x3", "tf32x3", "fp32"):
            t0 = time.time()
            f = G.hinv_cholesky_upper(h, 0.01, False, p)
            ok = f.ok
            dt = time.time() - t0
            u = f.u.double()
            hd = h.double().clone()
            hd.diagonal().add_(float(np.float32(0.01)) * hd.diagonal().mean())
            resid = (u.T @ u @ hd - torch.eye(k, device=dev, dtype=torch.float64)).abs().max().item()
            rec = {"ok": ok, "resid_inf": resid, "seconds": dt}
            if k == 4096:
                u_ref, ok_ref = O.hinv_cholesky_upper(h.cpu().numpy(), 0.01)
                rec["vs_oracle"] = float(np.abs(f.u.cpu().numpy() - u_ref).max() / np.abs(u_ref).max())
                # the oracle's own residual (float32 LAPACK, three calls)
                ur = torch.from_numpy(u_ref).to(dev).double()
                rec["oracle_resid_inf"] = (ur.T @ ur @ hd - torch.eye(k, device=dev, dtype=torch.float64)).abs().max().item()
            out[f"hinv|{k}|{p}"] = rec
            print("hinv", k, p, rec, flush=True)
            del u, hd, f
        del h, h64
        torch.cuda.empty_cache()

if "chain" in what:
    k = n = 4096
    for data in ("iid", "correlated"):
        rng = np.random.default_rng(5)
        t = 16384
        x = rng.standard_normal((t, k), dtype=np.float32)
        if data == "correlated":
            x *= rng.uniform(0.3, 3.0, k).astype(np.float32)
            x[:, 1:] += 0.5 * x[:, :-1]
        w = (rng.standard_normal((k, n), dtype=np.float32) * np.float32(0.02))
        h_np, _ = O.accumulate_hessian(x.reshape(8, -1, k), np.zeros((k, k), np.float32), 0)
        t0 = time.time()
        want = O.gptq(w, h_np, "int4", "group", 128, True, False, 1.0, 128, 0.01, False, False, None, "propagate",
                      return_aux=True)
        print("oracle gptq", data, f"{time.time() - t0:.1f}s", flush=True)
        want_codes = np.asarray(want[0]).astype(np.int32)
        e_want = O.layer_output_rel_mse(x, w, want[3]["deq"])
        xd, wd = torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev)
        for hp, sp in (("bf16x3", "bf16x3"), ("tf32x3", "tf32x3"), ("fp32", "tf32x3"), ("fp32", "fp32"), ("numpy", "tf32x3"),
                       ("numpy", "fp32")):
            if hp == "numpy":
                h = torch.from_numpy(h_np).to(dev)
            else:
                h = torch.zeros((k, k), device=dev)
                hessian_accumulate(xd.reshape(8, -1, k), h, alpha=2.0 / 8, beta=0.0, precision=hp)
            f = G.hinv_cholesky_upper(h, 0.01, False, sp)
            codes, s, z, deq = G.gptq_quantize(wd, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate",
                                               sp, return_deq=True)
            c = codes.cpu().numpy().view(np.int8).astype(np.int32)
            c = np.where(c > 7, c - 16, c)
            diff = np.abs(c - want_codes)
            e = O.layer_output_rel_mse(x, w, deq.cpu().numpy())
            rec = {"flip_frac": float((diff != 0).mean()), "max_diff": int(diff.max()), "n_gt1": int((diff > 1).sum()),
                   "rel_mse": e, "rel_mse_oracle": e_want, "rel_mse_dev": abs(e - e_want) / e_want,
                   "h_err_vs_numpy": float(np.abs(h.cpu().numpy() - h_np).max() / np.abs(h_np).max())}
            out[f"chain|{data}|H={hp}|solve={sp}"] = rec
            print("chain", data, hp, sp, rec, flush=True)

if "mse" in what:
    from onnx_quantize_b200 import device_api as D
    rng = np.random.default_rng(9)
    w = (rng.standard_normal((4096, 4096), dtype=np.float32) * np.float32(0.02))
    t0 = time.time()
    want = O.rtn_quantize(w, "uint4", "group", 128, False, False, 1.0, True)
    print("oracle mse", f"{time.time() - t0:.1f}s", flush=True)
    codes, s, z = D.rtn_quantize(torch.from_numpy(w).to(dev), "uint4", "group", 128, False, False, 1.0, True)
    s_np, z_np = s.cpu().numpy().reshape(-1), z.cpu().numpy().reshape(-1)
    ws, wz = np.asarray(want[1]).reshape(-1), np.asarray(want[2]).astype(np.uint8).reshape(-1)
    bad = (s_np.view(np.uint32) != ws.view(np.uint32)) | (z_np != wz)
    out["mse|4096x4096|g128"] = {"groups": int(bad.size), "groups_differing": int(bad.sum()),
                                 "codes_differing": int((codes.cpu().numpy() != np.asarray(want[0]).astype(np.uint8)).sum())}
    print("mse", out["mse|4096x4096|g128"], flush=True)

if "pd" in what:
    # where do the device and LAPACK disagree on positive definiteness?  rank-deficient H, shrinking damping
    for k in (1024,):
        rng = np.random.default_rng(1)
        x = rng.standard_normal((k // 2, k)).astype(np.float32)
        h32 = ((2.0 / x.shape[0]) * (x.T.astype(np.float64) @ x.astype(np.float64))).astype(np.float32)
        for pd in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 0.0):
            _, ok_ref = O.hinv_cholesky_upper(h32, pd)
            oks = {p_: G.hinv_cholesky_upper(torch.from_numpy(h32).to(dev), pd, False, p_).ok for p_ in ("bf16x3", "tf32x3", "fp32")}
            out[f"pd|{k}|{pd}"] = {"lapack": bool(ok_ref), **oks}
            print("pd", k, pd, ok_ref, oks, flush=True)

if "unit" in what:
    # flips of the correlated chain against the token length of a truncating accumulation unit
    k = n = 4096
    rng = np.random.default_rng(5)
    x = rng.standard_normal((16384, k), dtype=np.float32)
    x *= rng.uniform(0.3, 3.0, k).astype(np.float32)
    x[:, 1:] += 0.5 * x[:, :-1]
    w = (rng.standard_normal((k, n), dtype=np.float32) * np.float32(0.02))
    h_np, _ = O.accumulate_hessian(x.reshape(8, -1, k), np.zeros((k, k), np.float32), 0)
    want = O.gptq(w, h_np, "int4", "group", 128, True, False, 1.0, 128, 0.01, False, False, None, "propagate")
    want_codes = np.asarray(want[0]).astype(np.int32)
    xd, wd = torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev)
    h64 = (xd.double().T @ xd.double()) * (2.0 / 8)
    for unit in (256, 512, 1024, 2048, 4096):
        os.environ["B200Q_HESSIAN_BF16_UNIT"] = str(unit)
        h = torch.zeros((k, k), device=dev)
        t0 = time.time()
        hessian_accumulate(xd.reshape(8, -1, k), h, alpha=2.0 / 8, beta=0.0, precision="bf16x3")
        torch.cuda.synchronize()
        dt = time.time() - t0
        err = ((h.double() - h64).abs().max() / h64.abs().max()).item()
        f = G.hinv_cholesky_upper(h, 0.01, False, "bf16x3")
        codes = G.gptq_quantize(wd, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", "bf16x3")[0]
        c = codes.cpu().numpy().view(np.int8).astype(np.int32)
        c = np.where(c > 7, c - 16, c)
        diff = np.abs(c - want_codes)
        anyd = diff != 0
        first = np.argmax(anyd, axis=0)
        cols = np.nonzero(anyd.any(axis=0))[0]
        rec = {"h_err": err, "flip_frac": float(anyd.mean()), "n_gt1": int((diff > 1).sum()),
               "first_diff_gt1": int((diff[first[cols], cols] > 1).sum()), "hessian_ms": dt * 1e3}
        out[f"unit|{unit}"] = rec
        print("unit", unit, rec, flush=True)
    os.environ.pop("B200Q_HESSIAN_BF16_UNIT", None)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/explore_gptq_parity.json", "w") as f:
    json.dump(out, f, indent=1)
print("ok")
