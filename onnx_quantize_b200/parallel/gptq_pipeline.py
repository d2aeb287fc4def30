"""GPTQ of a whole set of layers on the ranks of one box (BASELINE config 5).

The path shards two ways at once:

* calibration tokens are split over the ranks — every rank contracts ITS tokens into a partial
  Hessian of every unit (a unit = one distinct layer input: q/k/v share one, gate/up share one);
* the solves (inverse factor + block loop of every weight of a unit) are assigned to owner ranks,
  largest first (``shard.assign_units``).

The one real exchange step is the sum of a unit's partial Hessians on its owner, over NVLink.  Units
of equal size with one owner on EVERY rank are exchanged together: their Hessian buffers are the
slices of one stacked tensor in owner order and a single in-place NCCL ``reduce_scatter`` leaves every
owner with its sum — all links busy at once instead of one ``reduce`` tree per unit (a ``reduce`` to a
single destination moved 0.2-0.3 TB/s on the 8-GPU box).  Leftover units use ``reduce``.  Phase 1 contracts every unit's Hessian (the tensor-core kernels are persistent and own all
148 SMs: anything issued next to them would wait for a chunk boundary and then stretch the tail of
the next chunk).  Phase 2 issues the reduces on a communication stream, longest solve first, and the
owner's solve streams wait for their OWN unit's reduce event only — the solves are chains of small
latency-bound launches that leave most SMs idle, so the remaining reduces run underneath them
(round 1 ran reduce and solve back to back: 23 ms of exposed reduces per 8 layers at 8 GPUs).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import torch
import torch.distributed as dist

from onnx_quantize_b200 import gptq_device as G
from onnx_quantize_b200.hessian import hessian_accumulate
from onnx_quantize_b200.parallel.shard import assign_units, world
from onnx_quantize_b200.parallel.streams import StreamPool


@dataclass
class GptqUnit:
    """One Hessian and the weights that share it."""

    name: str
    k: int
    weights: Sequence[torch.Tensor]                 # (K, N) float32 CUDA tensors (on the owner at least)
    tokens: torch.Tensor | Callable[[], torch.Tensor]   # this rank's share of the calibration tokens (T_local, K)
    n_samples_total: int                            # the reference's sample count n of (2/n)·XᵀX, over ALL ranks

    def solve_cost(self) -> float:
        return self.k ** 3 * 2.0 / 3 + sum(self.k * self.k * int(w.shape[1]) for w in self.weights)


@dataclass
class GptqSpec:
    quant_type: object = "int4"
    strategy: str = "group"
    group_size: int = 128
    is_symmetric: bool = True
    reduce_range: bool = False
    clip_ratio: float = 1.0
    mse: bool = False
    block_size: int = 128
    percdamp: float = 0.01
    actorder: bool = False
    mode: str = "propagate"
    precision: str = "bf16x3"


@dataclass
class GptqRun:
    """What ``gptq_quantize_units`` hands back: results on the owner, plus the events a caller can time."""

    results: dict = field(default_factory=dict)     # unit name -> [(codes, scale, zp), ...] (owner only)
    factors: dict = field(default_factory=dict)     # unit name -> HinvFactor (owner only; status not yet read)
    owner: dict = field(default_factory=dict)       # unit name -> rank
    start: torch.cuda.Event | None = None
    hessians_done: torch.cuda.Event | None = None   # main stream, after the last Hessian kernel
    reduces_done: torch.cuda.Event | None = None    # communication stream, after the last reduce
    end: torch.cuda.Event | None = None             # main stream, after every solve of this rank


def exchange_plan(ks: Sequence[int], order: Sequence[int], owners: Sequence[int], n_ranks: int) -> list[list[int]]:
    """The exchange steps of phase 2a, in ``order`` (longest solve first): lists of unit indices.  A list
    of ``n_ranks`` indices is one reduce-scatter — units of one size K, entry r owned by rank r; a list of
    one index is a plain reduce to its owner.  Every rank derives the same plan from the same inputs."""
    if n_ranks <= 1:
        return [[i] for i in order]
    queues: dict[int, list[list[int]]] = {}
    for i in order:
        queues.setdefault(ks[i], [[] for _ in range(n_ranks)])[owners[i]].append(i)
    head: dict[int, list[int]] = {}
    for k, per_owner in queues.items():
        for j in range(min(len(q) for q in per_owner)):
            group = [per_owner[r][j] for r in range(n_ranks)]
            for i in group:
                head[i] = group
    plan, seen = [], set()
    for i in order:
        if i in seen:
            continue
        group = head.get(i, [i])
        plan.append(group)
        seen.update(group)
    return plan


class GptqPipeline:
    """Reusable streams and Hessian buffers for repeated runs over the same set of units."""

    def __init__(self, n_solve_streams: int = 8, device=None, group=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.pool = StreamPool(n_solve_streams, self.device)
        self.comm = torch.cuda.Stream(device=self.device)
        self.kick = torch.cuda.Stream(device=self.device)
        self.group = group
        self._h: dict[str, torch.Tensor] = {}
        self._stacks: dict[tuple, torch.Tensor] = {}

    def _hessian_buffer(self, unit: GptqUnit) -> torch.Tensor:
        h = self._h.get(unit.name)
        if h is None or h.shape[0] != unit.k:
            h = self._h[unit.name] = torch.empty((unit.k, unit.k), dtype=torch.float32, device=self.device)
        return h

    def _stack_buffers(self, group: Sequence[GptqUnit]) -> torch.Tensor:
        """One (n, K, K) tensor whose slices are the Hessian buffers of ``group`` (owner order)."""
        key = tuple(u.name for u in group)
        st = self._stacks.get(key)
        if st is None:
            k = group[0].k
            st = self._stacks[key] = torch.empty((len(group), k, k), dtype=torch.float32, device=self.device)
        for r, u in enumerate(group):
            self._h[u.name] = st[r]
        return st

    def release(self) -> None:
        self._h.clear()
        self._stacks.clear()

    def _scatter_capable(self) -> bool:
        if not dist.is_available() or not dist.is_initialized():
            return False
        return str(dist.get_backend(self.group)) == "nccl"

    def run(self, units: Sequence[GptqUnit], spec: GptqSpec) -> GptqRun:
        rank, n_ranks = world()
        if self.group is not None and n_ranks > 1 and dist.get_world_size(self.group) != n_ranks:
            # owners are global ranks (`dist.reduce(dst=...)`) and slice r of a stacked buffer belongs to rank r
            raise ValueError("GptqPipeline needs a process group that spans all ranks")
        costs = [u.solve_cost() for u in units]
        plan = assign_units(costs, n_ranks)
        out = GptqRun()
        for r, idxs in enumerate(plan):
            for i in idxs:
                out.owner[units[i].name] = r
        order = sorted(range(len(units)), key=lambda i: (-costs[i], i))
        exchanges = exchange_plan([u.k for u in units], order, [out.owner[u.name] for u in units],
                                  n_ranks if n_ranks > 1 and self._scatter_capable() else 1)
        stacks = {}
        for ex in exchanges:
            if len(ex) > 1:
                stacks[ex[0]] = self._stack_buffers([units[i] for i in ex])
        main = torch.cuda.current_stream(self.device)
        ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
        out.start = ev()
        out.start.record(main)
        for i in order:                                    # phase 1 — G1 on every rank, its share of the tokens
            u = units[i]
            x = u.tokens() if callable(u.tokens) else u.tokens
            hessian_accumulate(x, self._hessian_buffer(u), alpha=2.0 / u.n_samples_total, beta=0.0,
                               precision=spec.precision)
        out.hessians_done = ev()
        out.hessians_done.record(main)
        landed: dict[str, torch.cuda.Event] = {}
        self.comm.wait_event(out.hessians_done)
        if n_ranks > 1:                                    # phase 2a — the exchange step, longest solve first
            with torch.cuda.stream(self.comm):
                for ex in exchanges:
                    if len(ex) > 1:                        # one unit per owner: in-place reduce-scatter
                        st = stacks[ex[0]]
                        dist.reduce_scatter_tensor(st[rank], st.view(-1, st.shape[-1]), op=dist.ReduceOp.SUM,
                                                   group=self.group)
                    else:
                        u = units[ex[0]]
                        dist.reduce(self._h[u.name], dst=out.owner[u.name], op=dist.ReduceOp.SUM, group=self.group)
                    done = torch.cuda.Event()
                    done.record(self.comm)
                    for i in ex:
                        landed[units[i].name] = done
        out.reduces_done = ev()
        out.reduces_done.record(self.comm)

        def solve(u: GptqUnit):                            # phase 2b — G2-G4 of one unit on one of the pool's streams
            def job():
                if u.name in landed:
                    torch.cuda.current_stream(self.device).wait_event(landed[u.name])
                f = G.hinv_cholesky_upper(self._h[u.name], spec.percdamp, spec.actorder, spec.precision)
                res = [G.gptq_quantize(w, f, spec.quant_type, spec.strategy, spec.group_size, spec.is_symmetric,
                                       spec.reduce_range, spec.clip_ratio, spec.mse, spec.block_size, spec.mode,
                                       spec.precision) for w in u.weights]
                return f, res
            return job

        mine = [i for i in order if out.owner[units[i].name] == rank]
        # StreamPool.run makes its streams wait for the CURRENT stream: a kick-off stream that is only
        # behind the Hessians, not behind the reduces queued on the communication stream
        self.kick.wait_event(out.hessians_done)
        with torch.cuda.stream(self.kick):
            solved = self.pool.run([solve(units[i]) for i in mine], [costs[i] for i in mine])
        main.wait_stream(self.kick)
        main.wait_stream(self.comm)
        for i, (f, res) in zip(mine, solved):
            out.factors[units[i].name] = f
            out.results[units[i].name] = res
        out.end = ev()
        out.end.record(main)
        return out
