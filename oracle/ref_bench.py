"""Time the compiled reference (oracle/_ref) on host cores.   TEST / BENCH INFRASTRUCTURE ONLY.

The reference's own scaling mechanism is one independent single-env simulation per process
(`multiprocessing.Pool.starmap(run_environment, ...)`, examples/JOCN_Benchmark_2024/graph_load.py:361-363);
this module does the same: each worker process owns one reference QRMSAEnv and runs the benchmark
loop `action,_,_ = heuristic(env); env.step(action)` (graph_load.py:161-163) on command.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time


def _worker(conn, topo_name, n_slots, load, seed, episode_length):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_harness as rh

    topo = rh.make_topology(topo_name)
    env = rh.make_env(topo, seed, n_slots=n_slots, load=load, episode_length=episode_length)
    heuristic = rh.first_fit_heuristic()
    conn.send("ready")
    while True:
        n = conn.recv()
        if n is None:
            break
        t0 = time.perf_counter()
        for _ in range(n):
            a, _, _ = heuristic(env)
            env.step(a)
        conn.send(time.perf_counter() - t0)
    conn.close()


class ReferencePool:
    """n_procs worker processes, each one reference env on its own request stream (seed base+i)."""

    def __init__(self, n_procs: int, topo_name: str = "nobel-eu", n_slots: int = 320, load: float = 300.0,
                 base_seed: int = 50, episode_length: int = 10_000_000):
        ctx = mp.get_context("spawn")
        self.n_procs = n_procs
        self.conns, self.procs = [], []
        for i in range(n_procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, topo_name, n_slots, load, base_seed + i, episode_length),
                            daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            assert c.recv() == "ready"

    def run(self, n_steps: int) -> float:
        """Every worker advances its env by n_steps requests; returns the wall time of the slowest."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(int(n_steps))
        for c in self.conns:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()


def c1_single_core(out_npz: str, n_steps: int = 10_000, topo_name: str = "nsfnet", n_slots: int = 320, load: float = 300.0,
                   seed: int = 50) -> None:
    """BASELINE config 1: ONE reference env, NSFNET / k=5 / 320 slots / first-fit, `n_steps` step() calls on one
    core.  The loop is the reference's own (`a,_,_ = heuristic(env); env.step(a)`) plus the
    recording of each request and decision (a few microseconds per 7 ms step), saved for the device parity replay."""
    import sys

    import numpy as np

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_harness as rh

    topo = rh.make_topology(topo_name)
    env = rh.make_env(topo, seed, n_slots=n_slots, load=load, episode_length=n_steps + 1)
    heuristic = rh.first_fit_heuristic()
    node_index = {n: i for i, n in enumerate(topo.graph["node_indices"])}
    rates = list(env.bit_rates)

    def fields(svc):
        return (node_index[svc.source], node_index[svc.destination], rates.index(int(svc.bit_rate)),
                np.float32(svc.arrival_time), np.float32(svc.holding_time))

    trace = [fields(env.current_service)]
    actions = np.zeros(n_steps, np.int64)
    t0 = time.perf_counter()
    for t in range(n_steps):
        a, _, _ = heuristic(env)
        env.step(a)
        actions[t] = a
        trace.append(fields(env.current_service))
    secs = time.perf_counter() - t0
    tr = np.array(trace, dtype=[("src", "u1"), ("dst", "u1"), ("rate", "u1"), ("arrival", "f4"), ("holding", "f4")])
    np.savez(out_npz, src=tr["src"], dst=tr["dst"], rate=tr["rate"], arrival=tr["arrival"], holding=tr["holding"],
             action=actions, seconds=secs, final_slots=np.array(env.topology.graph["available_slots"], dtype=np.uint8))


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    import platform

    return platform.processor() or "unknown"


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)
