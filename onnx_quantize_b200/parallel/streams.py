"""Independent GPTQ solves side by side on one GPU.

The inverse-Hessian factor and the block loop of one layer are chains of hundreds of small,
dependent launches (one-CTA diagonal-block factorisations, 128-CTA block kernels, skinny GEMMs):
on their own they leave most of the 148 SMs idle (ncu launch list in profiles/: K = 4096 factor
5.5 ms of which 2.5 ms is a single CTA).  Different Hessian groups and different layers do not
depend on each other, so they are issued on separate CUDA streams — the library is thread-safe per
stream and its workspaces are per (device, stream) — and the hardware interleaves the chains.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Sequence

import torch


class StreamPool:
    """``n`` side streams of the current device, reused across calls."""

    def __init__(self, n: int = 4, device=None):
        if n < 1:
            raise ValueError("n must be >= 1")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n)]

    def run(self, jobs: Sequence[Callable[[], object]], costs: Sequence[float] | None = None) -> list:
        """Run every job (a callable that enqueues GPU work on the CURRENT stream) on one of the
        pool's streams — longest first, each on the least-loaded stream — and make the caller's
        stream wait for all of them.  Returns the jobs' results in the order given.  Work already
        enqueued on the caller's stream is complete before any job starts."""
        main = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(main)
        n = len(self.streams)
        costs = [1.0] * len(jobs) if costs is None else [float(c) for c in costs]
        order = sorted(range(len(jobs)), key=lambda i: (-costs[i], i))
        load = [0.0] * n
        queues: list[list[int]] = [[] for _ in range(n)]
        for i in order:
            s = min(range(n), key=lambda j: (load[j], j))
            queues[s].append(i)
            load[s] += costs[i]
        results: list = [None] * len(jobs)
        dev_index = self.device.index

        def issue(s: int):
            # One host thread per stream: a factorisation is ~1000 launches, more than a stream's
            # launch queue holds, so a single issuing thread would block on the first chain and
            # leave the other streams empty.  The C calls release the GIL (ctypes); the current
            # stream and the library's hints / error text are thread-local.
            torch.cuda.set_device(dev_index)
            self.streams[s].wait_event(start)
            with torch.cuda.stream(self.streams[s]):
                for i in queues[s]:
                    results[i] = jobs[i]()
            done = torch.cuda.Event()
            done.record(self.streams[s])
            return done

        used = [s for s in range(n) if queues[s]]
        if len(used) <= 1:
            events = [issue(s) for s in used]
        else:
            with ThreadPoolExecutor(max_workers=len(used)) as ex:
                events = list(ex.map(issue, used))
        for done in events:
            main.wait_event(done)
        return results
