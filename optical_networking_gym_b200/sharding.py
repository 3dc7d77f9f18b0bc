"""Multi-GPU plumbing: environments shard trivially (contiguous env ranges, one process per GPU, static
tables replicated); the ONLY cross-GPU traffic of the path is one all-reduce (SUM) of the int64 counter
matrix [n_groups][N_COUNTERS] at episode end.  `torch.distributed` is used as plumbing: NCCL on
GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, last) env range of `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_total, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def env_seed(base_seed: int, global_env_index: int) -> int:
    """Env i of the whole job replays random.Random(base_seed + i), whatever the sharding."""
    return int(base_seed) + int(global_env_index)


def allreduce_counters(counters, group=None):
    """Sum the counter matrix over ranks.  Accepts a numpy int64 array (gloo / CPU tests) or a torch tensor
    (CUDA int64 for NCCL); returns the same kind.  No-op when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters
    if isinstance(counters, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(counters, np.int64).copy())
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def blocking_from_counters(c: np.ndarray) -> dict:
    """Blocking probabilities of one counter row (reference info keys, qrmsa.pyx:1013-1031)."""
    dec = max(int(c[0]), 1)
    req = max(int(c[3]), 1)
    return {"service_blocking_rate": (int(c[0]) - int(c[1])) / dec,
            "bit_rate_blocking_rate": (int(c[3]) - int(c[4])) / req}
