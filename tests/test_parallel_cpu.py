"""Multi-rank host logic on CPU: world_size-2 gloo process groups (the N>1 path of SURVEY.md §8e).
The collectives are backend-agnostic; on the GPU box the same functions run over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onnx_quantize_b200.parallel import calibration as C
from onnx_quantize_b200.parallel import shard as S
from oracle import np_oracle as O

LLAMA_LAYER = [(4096, 4096), (4096, 1024), (4096, 1024), (4096, 4096), (4096, 14336), (4096, 14336),
               (14336, 4096)]


def test_assign_units_balances_the_llama_set():
    costs = [k * n for _ in range(32) for (k, n) in LLAMA_LAYER]
    for ranks in (1, 2, 4, 8):
        plan = S.assign_units(costs, ranks)
        assert sorted(i for p in plan for i in p) == list(range(len(costs)))
        assert S.imbalance(costs, plan) < 0.02
    assert S.assign_units(costs, 8) == S.assign_units(costs, 8)       # deterministic
    assert S.assign_units([], 3) == [[], [], []]
    assert S.assign_units([5.0], 4)[0] == [0]


def test_exchange_plan_groups_one_unit_per_owner():
    from onnx_quantize_b200.parallel.gptq_pipeline import exchange_plan

    ks, costs = [], []
    for _ in range(32):
        for k, n in LLAMA_LAYER:
            ks.append(k)
            costs.append(k ** 3 * 2.0 / 3 + k * k * n)
    order = sorted(range(len(ks)), key=lambda i: (-costs[i], i))
    for ranks in (1, 2, 4, 8):
        owners = [None] * len(ks)
        for r, idxs in enumerate(S.assign_units(costs, ranks)):
            for i in idxs:
                owners[i] = r
        plan = exchange_plan(ks, order, owners, ranks)
        assert sorted(i for ex in plan for i in ex) == list(range(len(ks)))     # every unit exactly once
        for ex in plan:
            assert len(ex) in (1, ranks)
            if len(ex) > 1:                                  # one size, entry r owned by rank r
                assert len({ks[i] for i in ex}) == 1 and [owners[i] for i in ex] == list(range(ranks))
        firsts = [order.index(ex[0]) for ex in plan]
        assert firsts == sorted(firsts)                      # steps start in longest-solve-first order
        if ranks > 1:
            assert sum(len(ex) > 1 for ex in plan) >= len(ks) // ranks - 4
    # sizes that never line up fall back to single reduces; an owner without a unit of a size blocks its group
    assert exchange_plan([256, 512, 384], [1, 2, 0], [0, 1, 0], 2) == [[1], [2], [0]]
    assert exchange_plan([512, 512, 512], [0, 1, 2], [0, 0, 1], 2) == [[0, 2], [1]]


def test_shard_batches_round_robin():
    assert C.shard_batches(10, 0, 4) == [0, 4, 8] and C.shard_batches(10, 3, 4) == [3, 7]
    assert sorted(sum((C.shard_batches(10, r, 4) for r in range(4)), [])) == list(range(10))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)     # same stream on every rank
        n_batches, n_tensors, k = 7, 3, 24
        acts = rng.standard_normal((n_batches, n_tensors, 5, 11)).astype(np.float32)
        xs = rng.standard_normal((n_batches, 4, 6, k)).astype(np.float32)
        mine = C.shard_batches(n_batches)

        # ---- activation ranges, momentum 0: one MIN all-reduce --------------------------------
        local = torch.tensor([[acts[mine, t].min(), acts[mine, t].max()] for t in range(n_tensors)])
        got = C.allreduce_minmax(local.clone())
        want = torch.tensor([[acts[:, t].min(), acts[:, t].max()] for t in range(n_tensors)])
        assert torch.equal(got, want)

        # ---- momentum > 0: gather per-batch pairs, replay in global order ----------------------
        pairs = torch.tensor([[[acts[b, t].min(), acts[b, t].max()] for t in range(n_tensors)] for b in mine])
        allp = C.gather_batch_pairs(pairs, n_batches)
        assert allp.shape == (n_batches, n_tensors, 2)
        for m in (0.0, 0.9):
            state = C.replay_ema(allp, m).numpy()
            for t in range(n_tensors):
                ref = O.MinMax(momentum=m)
                for b in range(n_batches):
                    ref.collect("x", acts[b, t])
                assert np.float32(ref.stats["x"][0]) == state[t, 0] and np.float32(ref.stats["x"][1]) == state[t, 1]

        # ---- Hessian: weighted SUM all-reduce equals the single-process accumulation ----------
        h = np.zeros((k, k), np.float32)
        n = 0
        for b in mine:
            h, n = O.accumulate_hessian(xs[b], h, n)
        ht, n_total = C.allreduce_hessian(torch.from_numpy(h.copy()), n)
        h_ref = np.zeros((k, k), np.float32)
        n_ref = 0
        for b in range(n_batches):
            h_ref, n_ref = O.accumulate_hessian(xs[b], h_ref, n_ref)
        assert n_total == n_ref
        assert np.abs(ht.numpy() - h_ref).max() / np.abs(h_ref).max() < 1e-6
        # reduce-to-owner variant
        hd, _ = C.allreduce_hessian(torch.from_numpy(h.copy()), n, dst=1)
        if rank == 1:
            assert np.abs(hd.numpy() - h_ref).max() / np.abs(h_ref).max() < 1e-6

        # ---- gather of sharded results: packed tensors, point to point, no pickling ------------
        names = [f"w{i}" for i in range(5)]
        shapes = [(6, 4), (16, 8), (2, 2), (8, 12), (10, 4)]           # (K, N) per unit; codes (K,N) u8, scale/zp (N,)
        costs = [k * n for k, n in shapes]
        plan = S.assign_units(costs, world)
        dtypes = (torch.uint8, torch.float32, torch.uint8)
        layouts, totals = [], []
        for r in range(world):
            lay, tot = S.packed_layout([[((shapes[i]), dtypes[0]), ((shapes[i][1],), dtypes[1]),
                                         ((shapes[i][1],), dtypes[2])] for i in plan[r]])
            layouts.append(lay)
            totals.append(tot)
        local = [(torch.full(shapes[i], i, dtype=torch.uint8), torch.arange(shapes[i][1], dtype=torch.float32) + i,
                  torch.full((shapes[i][1],), 100 + i, dtype=torch.uint8)) for i in plan[rank]]
        flat = S.pack_results(local, layouts[rank], totals[rank], torch.device("cpu"))
        assert all(off % 256 == 0 for unit in layouts[rank] for off, _, _ in unit)
        bufs = S.gather_packed(flat, totals, dst=0)
        if rank == 0:
            seen = {}
            for r, buf in enumerate(bufs):
                for i, unit in zip(plan[r], S.unpack_results(buf, layouts[r])):
                    seen[names[i]] = unit
            assert sorted(seen) == names
            for i, n in enumerate(names):
                c, sc, z = seen[n]
                assert c.shape == shapes[i] and bool((c == i).all()) and bool((z == 100 + i).all())
                assert torch.equal(sc, torch.arange(shapes[i][1], dtype=torch.float32) + i)
        else:
            assert bufs is None
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_collectives():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
