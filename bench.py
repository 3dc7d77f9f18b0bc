#!/usr/bin/env python
"""bench.py -- QRMSA env-steps/s on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 65,536 envs per GPU on nobel-eu, 320 slots,
k=5, 6 modulations, first-fit heuristic, load 300 Erlang, launch power 1 dBm, bit rates
(10,40,100,400,1000), env i replaying the request stream random.Random(50 + i).

A bench "step" = one pass of the fused hot path over one batch of synthetic input = ONE kernel
launch of the step kernel (followed by the small decision-log counting kernel) that advances every env by
`chunk` requests (default 256).  Before the timed region every
env is brought to steady state by an untimed prefill of 1000 requests from the empty network
(SURVEY §8d C2).  `value` = env-steps/s with the trace resident in HBM; `e2e` = the same metric
through the host-buffer C-ABI calls for a whole episode (reset + H2D of the trace from pinned memory
+ schedule build + every step + D2H of every decision + counters).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QRMSA env-steps/sec"
UNIT = "env-steps/s"
TOPOLOGY, N_SLOTS, LOAD, BASE_SEED = "nobel-eu", 320, 300.0, 50
PREFILL = 1000
MAX_REQUESTS = 16384


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--chunk", type=int, default=256, help="requests per env per bench step (one launch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-slices", type=int, default=16,
                    help="env slices of the pipelined episode; at most one wave of warps (148 SMs x 32 envs) per "
                         "slice, so that slices interleave instead of leaving a tail: 65,536 envs -> 16")
    ap.add_argument("--ref-chunk", type=int, default=64, help="--impl reference: requests per env per step")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--topology", default=TOPOLOGY, help="other BASELINE configs (parity cases), e.g. germany50")
    ap.add_argument("--slots", type=int, default=N_SLOTS)
    ap.add_argument("--load", type=float, default=LOAD)
    return ap.parse_args()


def workload_name(n_envs):
    return (f"{TOPOLOGY}/{N_SLOTS}-slot/k=5/6-mod first-fit, load {LOAD:g}, launch 1 dBm, "
            f"{n_envs} envs per GPU, steady state after {PREFILL}-request prefill")


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled every few ms by an NVML thread while the timed region runs
    (nvidia-smi -lms takes longer to start than the region lasts); falls back to one nvidia-smi query."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.power, self.reasons, self.mx = [], [], set(), None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        try:
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mask = int(get(self.h))
            for bit, name in self.BITS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.004)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        elif self.nv is None:
            try:   # no NVML binding: one nvidia-smi query right after the region
                q = "clocks.sm,clocks.max.sm,power.draw"
                f = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.sm, self.mx, self.power = [float(f[0])], float(f[1]), [float(f[2])]
            except Exception:
                pass
        if self.sm:
            sm = sorted(self.sm)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.mx, reasons=sorted(self.reasons), samples=len(sm),
                       power_w_max=max(self.power) if self.power else None)
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own Cython path on host cores
# ------------------------------------------------------------------------------------------------
def reference_kind():
    from oracle import ref_harness

    return "reference" if ref_harness.available() else "port"


def time_reference(n_procs, warm_steps, timed_chunks, chunk):
    """Returns (env-steps/s aggregate, seconds, per-chunk seconds list)."""
    from oracle import ref_bench

    pool = ref_bench.ReferencePool(n_procs, TOPOLOGY, N_SLOTS, LOAD, BASE_SEED)
    try:
        if warm_steps:
            pool.run(warm_steps)
        per = [pool.run(chunk) for _ in range(timed_chunks)]
    finally:
        pool.close()
    total = sum(per)
    return n_procs * chunk * timed_chunks / total, total, per


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import ref_bench

    kind = reference_kind()  # "reference" = oracle/_ref (compiled Cython), else the oracle port
    procs = args.ref_procs or min(ref_bench.usable_cores(), 128)
    t0 = time.time()
    if kind == "reference":
        pool = ref_bench.ReferencePool(procs, TOPOLOGY, N_SLOTS, LOAD, BASE_SEED)
        try:
            pool.run(PREFILL)                       # untimed: reach steady state
            for _ in range(args.warmup):
                pool.run(args.ref_chunk)
            per = [pool.run(args.ref_chunk) for _ in range(args.steps)]
        finally:
            pool.close()
    else:
        per = time_port(procs, PREFILL, args.warmup, args.steps, args.ref_chunk)
    total = sum(per)
    value = procs * args.ref_chunk * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.envs), "reference_sample":
                   f"{procs} independent reference envs (one per process), {args.ref_chunk} requests each per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{procs} procs x {args.ref_chunk} requests x {args.steps} steps after a "
                                   f"{PREFILL}-request prefill; wall {time.time() - t0:.1f}s incl. setup"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _port_worker(conn, seed):
    import numpy as np

    from oracle import oracle as orc
    from optical_networking_gym_b200.tables import StaticTables
    from optical_networking_gym_b200.tracegen import TraceGenerator

    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", f"tables_{TOPOLOGY}_{N_SLOTS}.npz"))
    n = MAX_REQUESTS
    tr = TraceGenerator(1, tb.n_nodes, tb.n_rates, LOAD, base_seed=seed, n_threads=1).next(n)
    env = orc.OracleEnv(tb, n)
    env.reset(*[np.ascontiguousarray(a[:, 0]) for a in tr])
    conn.send("ready")
    while True:
        k = conn.recv()
        if k is None:
            break
        t0 = time.perf_counter()
        env.run_first_fit(k, log_qot=False)
        conn.send(time.perf_counter() - t0)


def time_port(procs, prefill, warmup, steps, chunk):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    conns, ps = [], []
    for i in range(procs):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_port_worker, args=(b, BASE_SEED + i), daemon=True)
        p.start(); conns.append(a); ps.append(p)
    for c in conns:
        c.recv()

    def run(k):
        t0 = time.perf_counter()
        for c in conns:
            c.send(k)
        for c in conns:
            c.recv()
        return time.perf_counter() - t0

    run(prefill)
    for _ in range(warmup):
        run(chunk)
    per = [run(chunk) for _ in range(steps)]
    for c in conns:
        c.send(None)
    for p in ps:
        p.join(timeout=5)
    return per


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_step(c, n_slots):
    """SURVEY §8(d) model, evaluated with the run's own measured Lr, Nq, a, h (counter deltas)."""
    dec = max(c["decided"], 1)
    acc = max(c["accepted"], 1)
    W, R, C, V = 4 * ((n_slots + 31) // 32), 11, 4, 16
    Lr = c["links_read"] / dec
    Nq = c["records_read"] / dec
    a = c["accepted"] / dec
    h = c["hops_accepted"] / acc
    rel = c["releases"] / dec
    b = R + Lr * W + Nq * C + a * (h * W + h * C + V + 8) + rel * (V + 8 + 2 * h * W + 2 * h * C)
    return b, dict(Lr=Lr, Nq=Nq, accept_ratio=a, hops=h, releases_per_step=rel,
                   gn_evals_per_step=c["gn_evals"] / dec, gn_terms_per_step=c["gn_terms"] / dec,
                   gn_pruned_per_step=c.get("gn_pruned", 0) / dec)


def run_b200(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.tables import StaticTables
    from optical_networking_gym_b200.tracegen import TraceGenerator

    K, Wm = args.steps, max(args.warmup, 0)
    chunk = max(1, min(args.chunk, (MAX_REQUESTS - PREFILL - 1) // max(K + Wm, 1)))
    n_req = PREFILL + (K + Wm) * chunk + 1
    n_envs = args.envs
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", f"tables_{TOPOLOGY}_{N_SLOTS}.npz"))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(vec):
        t = torch.tensor(vec, dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the path's only collective: counters at episode end
        return t.cpu().numpy()

    # ---- synthetic input: request streams in pinned host memory
    t_gen = time.time()
    shape = (n_req, n_envs)
    pinned = [torch.empty(shape, dtype=dt, pin_memory=True) for dt in
              (torch.uint8, torch.uint8, torch.uint8, torch.float32, torch.float32)]
    trace = [p.numpy() for p in pinned]
    gen = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, LOAD, base_seed=BASE_SEED + rank * n_envs)
    gen.next(n_req, out=trace)
    gen.close()
    t_gen = time.time() - t_gen

    eng = Engine(tb, n_envs, n_req, device=local_rank)
    stream = torch.cuda.current_stream()

    # ---- phase A: trace resident in HBM, steady state
    eng.reset()
    eng.load_trace_host(*trace)
    eng.step_first_fit(PREFILL)
    for _ in range(Wm):
        eng.step_first_fit(chunk)
    barrier()
    c0 = eng.counters().sum(0)
    sampler = ClockSampler(local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(K):
        eng.step_first_fit(chunk)
        ev[i + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    elapsed_ms = max_over_ranks(ev[0].elapsed_time(ev[K]))
    launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    c1 = eng.counters().sum(0)
    delta = sum_over_ranks((c1 - c0).tolist())
    cd = {n: int(delta[i]) for i, n in enumerate(_lib.COUNTER_NAMES)}
    cd["gn_pruned"] = int(delta[24])
    env_steps = cd["decided"]
    assert env_steps == world * n_envs * chunk * K, (env_steps, world, n_envs, chunk, K)
    assert cd["errors"] == 0
    value = env_steps / (elapsed_ms * 1e-3)
    bytes_step, params = algorithmic_bytes_per_step(cd, N_SLOTS)

    peaks, peak_src = {}, "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    avg_launch_s = (sum(launch_ms) / len(launch_ms)) * 1e-3
    achieved = bytes_step * (n_envs * chunk) / avg_launch_s / 1e9   # per GPU, this rank's kernel
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj0 = json.load(open(tp))   # ncu --set full capture of this kernel (profiles/README.md)
            traffic = tj0.get("dram_bytes_per_launch")
            if tj0.get("dram_bytes_per_env_step"):   # per launch of THIS run's size
                traffic = float(tj0["dram_bytes_per_env_step"]) * n_envs * chunk
        except Exception:
            traffic = None
    # what actually binds the kernel (profiles/README.md): warp-instruction issue.  Instructions per env-step come from
    # the committed ncu capture of this kernel, the rate and the clock from this run.
    issue = None
    try:
        tj = json.load(open(tp))
        ipe = float(tj["warp_instructions_per_env_step"])
        sm_clock = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        issue = {"bound": "issue", "achieved": ipe * (n_envs * chunk) / avg_launch_s, "peak": n_sm * 4 * sm_clock,
                 "unit": "warp-instructions/s", "frac": ipe * (n_envs * chunk) / avg_launch_s / (n_sm * 4 * sm_clock),
                 "warp_instructions_per_env_step": ipe, "source": tj.get("source")}
    except Exception:
        issue = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_step_policy<320,6,5,first_fit> (+ k_count_decisions, <1 % of the launch)", "peak_source": peak_src,
                "algorithmic_bytes_per_env_step": bytes_step, "env_steps_per_launch": n_envs * chunk,
                "avg_launch_ms": avg_launch_s * 1e3, "workload_params": params, "issue_roofline": issue}

    # ---- phase B: end to end through the host-buffer C-ABI calls, one whole episode.  The public call is
    # PipelinedEpisodes.run: env slices on separate contexts/streams so that upload, kernels and download overlap.
    e2e = None
    if not args.no_e2e:
        from optical_networking_gym_b200.pipeline import PipelinedEpisodes

        out_actions = torch.empty((n_req - 1, n_envs), dtype=torch.int32, pin_memory=True)
        pipe = PipelinedEpisodes(tb, n_envs, n_req, slices=args.e2e_slices, device=local_rank)
        pipe.run(pinned, out_actions, launch_steps=512)          # untimed warm-up episode (first-touch, staging buffers)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        cnt = pipe.run(pinned, out_actions, launch_steps=512).sum(0)   # reset + H2D + schedule + steps + D2H + counters
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) * 1e-3
        t_e2e = max_over_ranks(max(wall, dev_s))
        assert int(cnt[0]) == n_envs * (n_req - 1)
        pipe.close()
        e2e = {"value": world * n_envs * (n_req - 1) / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": 11 * n_envs * chunk, "d2h_bytes_per_step": 4 * n_envs * chunk,
               "episode_requests": n_req, "seconds": t_e2e, "slices": args.e2e_slices,
               "note": "whole episode from reset (empty network): trace upload from pinned host memory, schedule "
                       "build, every step, download of every decision and the counters, through the host-buffer "
                       "C-ABI calls on env slices overlapped over CUDA streams; bytes are per bench step of "
                       "`chunk` requests per env"}

    # ---- phase C (informational): the same episode with the requests drawn on the device (qrmsa_generate_trace,
    # Philox streams) instead of uploaded: no host generation, no H2D; decisions still downloaded
    e2e_dev = None
    if not args.no_e2e:
        out_words = torch.empty((n_req - 1, n_envs), dtype=torch.int32, pin_memory=True)
        def episode():
            eng.reset()
            eng.generate_trace(n_req, LOAD, seed=BASE_SEED, env_offset=rank * n_envs)
            done = 0
            while done < n_req - 1:
                c = min(512, n_req - 1 - done)
                eng.step_first_fit(c)
                done += c
            eng.actions_host_strided(0, n_req - 1, out_words.data_ptr(), n_envs)
            return eng.counters().sum(0)
        episode()
        barrier()
        t0 = time.perf_counter()
        cnt = episode()
        barrier()
        t_dev = max_over_ranks(time.perf_counter() - t0)
        assert int(cnt[0]) == n_envs * (n_req - 1)
        e2e_dev = {"value": world * n_envs * (n_req - 1) / t_dev, "unit": UNIT, "seconds": t_dev,
                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n_envs * chunk,
                   "note": "requests generated on the device (Philox), decisions downloaded; not the contract's e2e"}

    # ---- CPU baseline: the reference's Cython path on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import ref_bench

            kind = reference_kind()
            procs = args.ref_procs or min(ref_bench.usable_cores(), 128)
            if kind == "reference":
                v, secs, _ = time_reference(procs, PREFILL, 4, 250)
                sample = f"{procs} procs x 1000 requests of the same workload after a {PREFILL}-request prefill"
            else:
                per = time_port(procs, PREFILL, 1, 4, 5000)
                v, secs = procs * 5000 * 4 / sum(per), sum(per)
                sample = f"{procs} procs x 20000 requests (C restatement) after a {PREFILL}-request prefill"
            cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample, "seconds": secs}
        except Exception as ex:  # the baseline is reported, never a gate
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n_envs), "chunk_requests_per_env_per_step": chunk,
                       "episode_requests": n_req, "l2_policy": "inputs larger than L2 (per-GPU env state + trace "
                       f"touched per step >= {n_envs * (tb.n_links * 64 + 16 * chunk) / 1e6:.0f} MB)",
                       "trace": f"CPython-random-exact streams, seeds {BASE_SEED}+i, generated on host in {t_gen:.1f}s"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_device_requests": e2e_dev,
            "gpu_launches": 2 * K, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    global TOPOLOGY, N_SLOTS, LOAD
    args = parse_args()
    TOPOLOGY, N_SLOTS, LOAD = args.topology, args.slots, args.load
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
