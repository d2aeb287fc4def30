"""parallel/gptq_pipeline.py on one GPU: units (one Hessian, several weights) solved side by side on
the stream pool must give what the oracle's `_gptq` gives for every weight (gptq.py:119-243), and what
the plain device calls give.  The multi-rank exchange is exercised by tools/check_multi_gpu.py under
torchrun (profiles/r2_check_multi_gpu_n2.log) and its planning by tests/test_parallel_cpu.py."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from tests.helpers import stable_seed

pytestmark = pytest.mark.gpu


def _units(device, seed):
    from onnx_quantize_b200.parallel.gptq_pipeline import GptqUnit

    rng = np.random.default_rng(seed)
    raw = []
    for u, (k, ns) in enumerate(((512, (256, 128)), (256, (384,)), (384, (128, 64)), (512, (64,)))):
        x = (rng.standard_normal((16, 64, k)) * rng.uniform(0.5, 2.0, k)).astype(np.float32)
        raw.append((f"u{u}", k, x, [(rng.standard_normal((k, n)) * 0.05).astype(np.float32) for n in ns]))
    units = [GptqUnit(name, k, [torch.from_numpy(w).to(device) for w in ws],
                      torch.from_numpy(x.reshape(-1, k)).to(device), 16) for name, k, x, ws in raw]
    return raw, units


def _signed4(codes):
    c = codes.cpu().numpy().view(np.int8).astype(np.int32)
    return np.where(c > 7, c - 16, c)


@pytest.mark.parametrize("streams", [1, 4])
def test_pipeline_matches_oracle_and_plain_calls(streams):
    from onnx_quantize_b200 import gptq_device as G
    from onnx_quantize_b200.hessian import hessian_accumulate
    from onnx_quantize_b200.parallel.gptq_pipeline import GptqPipeline, GptqSpec

    device = torch.device("cuda", 0)
    raw, units = _units(device, stable_seed("gptq_pipeline", streams))
    spec = GptqSpec("int4", "group", 128, True, mode="propagate", precision="bf16x3")
    pipe = GptqPipeline(streams, device)
    run = pipe.run(units, spec)
    torch.cuda.synchronize()
    assert set(run.results) == {name for name, *_ in raw} and all(r == 0 for r in run.owner.values())
    assert run.start.elapsed_time(run.end) > 0 and run.hessians_done is not None
    total = flips = 0
    for (name, k, x, ws), unit in zip(raw, units):
        assert run.factors[name].ok
        h = torch.zeros((k, k), device=device)
        hessian_accumulate(unit.tokens, h, 2.0 / 16, 0.0, "bf16x3")
        f = G.hinv_cholesky_upper(h, 0.01, False, "bf16x3")
        for w_np, w, (codes, scale, zp) in zip(ws, unit.weights, run.results[name]):
            plain = G.gptq_quantize(w, f, "int4", "group", 128, True, False, 1.0, False, 128, "propagate", "bf16x3")
            d_plain = np.abs(_signed4(codes) - _signed4(plain[0]))
            assert d_plain.max() <= 1 and (d_plain != 0).mean() <= 1e-3      # split reductions are unordered
            want = O.gptq_quantize(w_np, x, "int4", "group", 128, True, mode="propagate")
            d = np.abs(_signed4(codes) - np.asarray(want[0]).astype(np.int32))
            assert d.max() <= 1
            flips += int((d != 0).sum())
            total += d.size
            assert scale.shape == plain[1].shape and zp.shape == plain[2].shape
    assert flips <= 1e-3 * total
    # a second run over the same units reuses the pipeline's buffers and streams
    again = pipe.run(units, spec)
    torch.cuda.synchronize()
    for name in run.results:
        for (c0, *_), (c1, *_) in zip(run.results[name], again.results[name]):
            d = np.abs(_signed4(c0) - _signed4(c1))
            assert d.max() <= 1 and (d != 0).mean() <= 1e-3
    pipe.release()
