"""Whole-episode first-fit runs with upload / compute / download overlapped.

`Engine.load_trace_host` + `step_first_fit` + `actions_host` on one context serialise H2D -> kernels -> D2H.
`PipelinedEpisodes` splits the env batch into `slices` contiguous env ranges, each with its own context
(include/qrmsa_b200.h: one context per env slice, no shared state) and CUDA stream: while slice i computes, slice
i+1 uploads its requests and slice i-1 downloads its decisions.  Results are identical to the single-context run
(envs are independent); tests/test_gpu_full_size.py::test_pipelined_equals_single_context checks that.
Slice size: the step kernel runs one warp per env on one 1024-thread CTA per SM, so a slice of at most
148 x 32 = 4,736 envs is one wave and the slices' kernels interleave on the SMs; larger slices leave a partial
second wave behind (measured, 65,536 envs: 4 / 8 / 12 / 16 / 32 slices -> 8.0 / 8.0 / 7.7 / 8.4 / 8.6e8 env-steps/s).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .engine import Engine
from .sharding import shard_range
from .tables import StaticTables


class PipelinedEpisodes:
    def __init__(self, tables: StaticTables, n_envs: int, max_requests: int, slices: int = 4, device: int = 0,
                 n_groups_per_slice: int = 1):
        import torch

        self.n_envs, self.max_requests, self.slices = int(n_envs), int(max_requests), int(slices)
        self.device = device
        self.ranges = [shard_range(self.n_envs, i, self.slices) for i in range(self.slices)]
        self.engines = [Engine(tables, b - a, max_requests, device=device) for a, b in self.ranges]
        with torch.cuda.device(device):
            self.streams = [torch.cuda.Stream() for _ in range(self.slices)]

    def run(self, trace: Sequence, out_actions=None, n_requests: int = None, launch_steps: int = 512,
            per_slice: bool = False) -> np.ndarray:
        """trace: five PINNED torch tensors [n_requests, n_envs] (uint8 x3, float32 x2); out_actions: pinned int32
        [n_requests-1, n_envs], or None when only the counters are wanted (load sweeps).  Everything is enqueued
        asynchronously; returns the counters after a sync -- summed over the slices, or one matrix per slice
        (`per_slice=True`: [slices][n_groups][N_COUNTERS], for per-load-point sums)."""
        import torch

        n_req = int(trace[0].shape[0] if n_requests is None else n_requests)
        assert all(t.is_pinned() for t in trace) and (out_actions is None or out_actions.is_pinned())
        esz = [t.element_size() for t in trace]
        cur = torch.cuda.current_stream(self.device)
        for (a, b), eng, st in zip(self.ranges, self.engines, self.streams):
            st.wait_stream(cur)
            eng.reset(stream=st)
            eng.load_trace_host_strided([t.data_ptr() + a * s for t, s in zip(trace, esz)], n_req, self.n_envs, stream=st)
            done = 0
            while done < n_req - 1:
                n = min(launch_steps, n_req - 1 - done)
                eng.step_first_fit(n, stream=st)
                done += n
            if out_actions is not None:
                eng.actions_host_strided(0, n_req - 1, out_actions.data_ptr() + a * 4, self.n_envs, stream=st)
        for st in self.streams:
            cur.wait_stream(st)
        each = [eng.counters(stream=st) for eng, st in zip(self.engines, self.streams)]
        torch.cuda.synchronize(self.device)
        return np.stack(each) if per_slice else sum(each[1:], each[0])

    def close(self):
        for e in self.engines:
            e.close()
