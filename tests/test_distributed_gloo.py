"""N>1 host logic on CPU: world_size-2 `gloo` run of the env sharding + the episode-end counter all-reduce
(the path's only collective).  Each rank plays its shard of envs with the ORACLE (no GPU here) and
the reduced counters must equal a single-process run over all envs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import ROOT, load_tables

N_ENVS, N_STEPS, BASE_SEED, LOAD = 6, 120, 77, 250.0


def _play(tb, first, last):
    from oracle import oracle as orc
    from optical_networking_gym_b200.tracegen import TraceGenerator

    c = np.zeros((1, 32), np.int64)
    for e in range(first, last):
        tr = TraceGenerator(1, tb.n_nodes, tb.n_rates, LOAD, base_seed=BASE_SEED + e, n_threads=1).next(N_STEPS + 1)
        o = orc.OracleEnv(tb, N_STEPS + 1)
        o.reset(*[a[:, 0] for a in tr])
        o.run_first_fit(N_STEPS, log_qot=False)
        k = o.counters()
        c[0, 0] += N_STEPS
        c[0, 1] += k["ep_accepted"]
        c[0, 2] += k["bl_reject"]
        c[0, 3] += int(round(k["bit_rate_requested"] * 1000))
        c[0, 4] += int(round(k["bit_rate_provisioned"] * 1000))
    return c


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from optical_networking_gym_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tb = load_tables("nsfnet", 320)
    first, last = sharding.shard_range(N_ENVS, rank, world)
    local = _play(tb, first, last)
    total = sharding.allreduce_counters(local)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), total)
    np.save(os.path.join(out_dir, f"l{rank}.npy"), local)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range():
    from optical_networking_gym_b200 import sharding

    for n, w in [(65536, 8), (10, 3), (1, 2), (0, 4), (7, 7)]:
        spans = [sharding.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 4, 4)


def test_two_rank_counters_equal_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    l0, l1 = np.load(tmp_path / "l0.npy"), np.load(tmp_path / "l1.npy")
    assert np.array_equal(r0, r1) and np.array_equal(r0, l0 + l1)
    single = _play(load_tables("nsfnet", 320), 0, N_ENVS)
    assert np.array_equal(r0, single)
    from optical_networking_gym_b200 import sharding

    b = sharding.blocking_from_counters(r0[0])
    assert 0.0 <= b["service_blocking_rate"] < 1.0 and 0.0 <= b["bit_rate_blocking_rate"] < 1.0
