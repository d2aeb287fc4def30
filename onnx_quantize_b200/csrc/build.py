"""Build libb200quant.so in-tree with nvcc for sm_100a (B200).  No other target is supported."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB = os.path.join(OUT_DIR, "libb200quant.so")

# translation unit -> extra flags.  The bit-exact integer/float32 paths forbid FMA contraction.
SOURCES = {
    "api.cu": [],
    "rtn.cu": ["-fmad=false"],
    "minmax.cu": ["-fmad=false"],
}
OPTIONAL = {
    "hessian.cu": [],
    "gemm_tn.cu": [],
    "gemm_tn_tc.cu": [],
    "linalg.cu": [],
    "gptq.cu": ["-fmad=false"],
    "awq.cu": ["-fmad=false"],
    "hqq.cu": ["-fmad=false"],
    "mlp.cu": [],
    "dense_bf16.cu": [],
}

COMMON = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "b200q.h"))
    units = dict(SOURCES)
    for k, v in OPTIONAL.items():
        if os.path.exists(os.path.join(HERE, k)):
            units[k] = v

    def compile_one(item):
        src, extra = item
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        path = os.path.join(HERE, src)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc, *COMMON, *extra, "-c", path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(units))) as ex:
        objs = list(ex.map(compile_one, units.items()))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
