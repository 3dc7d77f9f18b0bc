#!/usr/bin/env python
"""BASELINE config 3: JOCN_Benchmark_2024-style load sweep (blocking probability vs offered load).

    python examples/load_sweep.py [--envs-per-load 4096] [--episode-length 1000] [--loads 100 200 300 400 500]
    torchrun --nproc-per-node N examples/load_sweep.py ...          # envs sharded over N GPUs

The reference runs one simulation per (load, episode) in a multiprocessing.Pool
(examples/JOCN_Benchmark_2024/graph_load.py:340-363, loads of :18-19, README: episodes of 1000 requests) and
writes one CSV row per episode.  Here every (load, replica) is one env of a batch: envs of one load point are
contiguous (one counter group per load), env i replays random.Random(seed + i), and the only cross-GPU traffic
is the all-reduce of the per-load counter matrix at episode end.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    from optical_networking_gym_b200 import sharding
    from optical_networking_gym_b200.env import BatchedQRMSAEnv
    from optical_networking_gym_b200.tables import StaticTables

    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-load", type=int, default=4096)
    ap.add_argument("--episode-length", type=int, default=1000)
    ap.add_argument("--loads", type=float, nargs="+", default=[100, 200, 300, 400, 500])
    ap.add_argument("--topology", default="nobel-eu")
    ap.add_argument("--slots", type=int, default=320)
    ap.add_argument("--seed", type=int, default=50)
    ap.add_argument("--requests", choices=["replay", "device"], default="replay",
                    help="replay: CPython-exact host streams (reference parity); device: Philox streams drawn on the GPU")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", f"tables_{args.topology}_{args.slots}.npz"))
    L = len(args.loads)
    # this rank's replicas of every load point; global env id = load_index * envs_per_load + replica
    r0, r1 = sharding.shard_range(args.envs_per_load, rank, world)
    per = r1 - r0
    loads = np.repeat(np.asarray(args.loads, np.float64), per)
    t0 = time.time()
    env = BatchedQRMSAEnv(tb, L * per, num_spectrum_resources=args.slots, episode_length=args.episode_length, load=loads,
                          bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, bandwidth=args.slots * 12.5e9,
                          seed=args.seed + (rank * 10_000_019 if args.requests == "replay" else 0), n_groups=L, device=local,
                          reset=False, request_source=args.requests, env_offset=rank * L * per)
    env.reset()
    sharding.allreduce_counters(np.zeros((L, 32), np.int64))   # NCCL communicator set-up is not part of the episode
    t_setup = time.time() - t0
    torch.cuda.synchronize()
    t0 = time.time()
    while not env.terminated:
        env.step_first_fit(512)
    counters = sharding.allreduce_counters(env.engine.counters())        # [L][32], summed over GPUs
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        n_total = args.envs_per_load * L * (args.episode_length - 1)
        print(f"{args.topology}/{args.slots}: {args.envs_per_load * L} envs x {args.episode_length - 1} steps on {world} GPU(s): "
              f"{dt:.2f}s = {n_total / dt:,.0f} env-steps/s (setup incl. {args.requests} request generation {t_setup:.1f}s)")
        print("load,episodes,service_blocking_rate,ci95,bit_rate_blocking_rate,near_threshold_decisions")
        for i, load in enumerate(args.loads):
            c = counters[i]
            p = (c[0] - c[1]) / c[0]
            ci = 1.96 * (p * (1 - p) / c[0]) ** 0.5
            print(f"{load:g},{args.envs_per_load},{p:.5f},{ci:.5f},{(c[3] - c[4]) / c[3]:.5f},{int(c[11])}")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
