#!/usr/bin/env python
"""bench.py -- QRMSA env-steps/s on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--configs C1,C3,C4,C5|none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): 65,536 envs per GPU on nobel-eu, 320 slots,
k=5, 6 modulations, first-fit heuristic, load 300 Erlang, launch power 1 dBm, bit rates
(10,40,100,400,1000), env i replaying the request stream random.Random(50 + i).

A bench "step" = one pass of the fused hot path over one batch of synthetic input = ONE kernel
launch of the step kernel (followed by the small decision-log counting kernel) that advances every env by
`chunk` requests (default 256).  Before the timed region every
env is brought to steady state by an untimed prefill of 1000 requests from the empty network
(SURVEY §8d C2).  `value` = env-steps/s with the trace resident in HBM; `e2e` = the same metric
through the host-buffer C-ABI calls for a whole episode (reset + H2D of the trace from pinned memory
+ schedule build + every step + D2H of every decision + counters [+ their all-reduce at N > 1]).

The same JSON line carries a `configs` block with the other BASELINE.json configurations at their stated scale, each
with its own `value`, `roofline` (own Lr / Nq / a / h), `e2e` and a `parity_sample` -- envs of the TIMED state replayed
through the oracle after the timed region:
  C1  single env, NSFNET, 10,000 step()s: the compiled reference on ONE host core, the same trace on the device
  C3  JOCN-style load sweep, nobel-eu, loads 100..500 x 200,000 envs = 1 M envs sharded over the N GPUs (strong)
  C4  germany50 / 640 slots / load 800, 524,288 envs per GPU (4 M envs at N = 8)
  C5  PPO rollout, NSFNET, 16,384 envs per GPU x 1,024 steps, observation + action mask on the device, torch policy
"""
from __future__ import annotations

import argparse
import hashlib
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
BIT_RATES = (10, 40, 100, 400, 1000)


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
    ap.add_argument("--configs", default="C1,C3,C4,C5", help="BASELINE configs measured besides the headline, or 'none'")
    ap.add_argument("--c3-envs-per-load", type=int, default=200_000, help="C3: envs per load point over ALL GPUs")
    ap.add_argument("--c4-envs", type=int, default=524_288, help="C4: envs per GPU")
    ap.add_argument("--c5-envs", type=int, default=16_384, help="C5: envs per GPU")
    ap.add_argument("--c5-steps", type=int, default=1024)
    ap.add_argument("--parity-envs", type=int, default=16, help="envs per config replayed through the oracle")
    ap.add_argument("--lib", default="", help="kernel experiments: load this prebuilt libqrmsa_b200.so instead")
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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "cpu_model": ref_bench.cpu_model(),
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
# shared pieces of the B200 arm
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


def kernel_source_hash():
    """sha256 over the CUDA sources the shipped library is built from: ties profiles/traffic.json to a kernel state."""
    h = hashlib.sha256()
    for f in ("qrmsa_kernels.cuh", "qrmsa_b200.cu"):
        h.update(open(os.path.join(ROOT, "optical_networking_gym_b200", "csrc", f), "rb").read())
    return h.hexdigest()


class Cx:
    """Per-process plumbing: ranks, barrier, reductions over ranks, HBM peak."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.local, self.world = rank, local_rank, world
        self.dev = torch.device("cuda", local_rank)
        self.peak, self.peak_src, self.peaks = 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)", {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                self.peaks = json.load(f)
            self.peak, self.peak_src = float(self.peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, vec):
        t = self.torch.tensor(vec, dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def all_zero(self, x):
        """True iff x == 0 on every rank."""
        return int(self.sum_over_ranks([int(x)])[0]) == 0

    def stream(self):
        return self.torch.cuda.current_stream(self.dev)

    def events(self, n):
        return [self.torch.cuda.Event(enable_timing=True) for _ in range(n)]

    def pinned_trace(self, n_req, n_envs):
        t = self.torch
        return [t.empty((n_req, n_envs), dtype=dt, pin_memory=True) for dt in (t.uint8, t.uint8, t.uint8, t.float32, t.float32)]


def counter_dict(delta):
    from optical_networking_gym_b200 import _lib

    cd = {n: int(delta[i]) for i, n in enumerate(_lib.COUNTER_NAMES)}
    cd["gn_pruned"] = int(delta[24])
    return cd


def make_roofline(cx, cd, n_slots, env_steps_per_launch, avg_launch_s, kernel):
    bytes_step, params = algorithmic_bytes_per_step(cd, n_slots)
    achieved = bytes_step * env_steps_per_launch / avg_launch_s / 1e9   # per GPU, this rank's kernel
    return {"bound": "hbm", "achieved": achieved, "peak": cx.peak, "unit": "GB/s", "frac": achieved / cx.peak,
            "traffic": None, "kernel": kernel, "peak_source": cx.peak_src, "algorithmic_bytes_per_env_step": bytes_step,
            "env_steps_per_launch": env_steps_per_launch, "avg_launch_ms": avg_launch_s * 1e3, "workload_params": params}


def merge_parity(cx, res):
    """Sum a parity_sample dict over the ranks (every rank replays its own sampled envs)."""
    keys = [k for k in res if k not in ("steps", "checker")]
    tot = cx.sum_over_ranks([int(res[k]) for k in keys])
    out = dict(res)
    out.update({k: int(v) for k, v in zip(keys, tot)})
    return out


# ------------------------------------------------------------------------------------------------
# headline: BASELINE config 2
# ------------------------------------------------------------------------------------------------
def run_headline(cx, args):
    import numpy as np

    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.tables import StaticTables
    from optical_networking_gym_b200.tracegen import TraceGenerator
    from oracle import checker

    torch, rank, world = cx.torch, cx.rank, cx.world
    K, Wm = args.steps, max(args.warmup, 0)
    chunk = max(1, min(args.chunk, (MAX_REQUESTS - PREFILL - 1) // max(K + Wm, 1)))
    n_req = PREFILL + (K + Wm) * chunk + 1
    n_envs = args.envs
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", f"tables_{TOPOLOGY}_{N_SLOTS}.npz"))

    # ---- synthetic input: request streams in pinned host memory
    t_gen = time.time()
    pinned = cx.pinned_trace(n_req, n_envs)
    trace = [p.numpy() for p in pinned]
    gen = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, LOAD, base_seed=BASE_SEED + rank * n_envs)
    gen.next(n_req, out=trace)
    gen.close()
    t_gen = time.time() - t_gen

    eng = Engine(tb, n_envs, n_req, device=cx.local)
    stream = cx.stream()

    # ---- phase A: trace resident in HBM, steady state
    eng.reset()
    eng.load_trace_host(*trace)
    eng.step_first_fit(PREFILL)
    for _ in range(Wm):
        eng.step_first_fit(chunk)
    cx.barrier()
    c0 = eng.counters().sum(0)
    sampler = ClockSampler(cx.local)
    ev = cx.events(K + 1)
    cx.barrier()
    ev[0].record(stream)
    for i in range(K):
        eng.step_first_fit(chunk)
        ev[i + 1].record(stream)
    cx.barrier()
    clocks = sampler.stop()
    elapsed_ms = cx.max_over_ranks(ev[0].elapsed_time(ev[K]))
    launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
    c1 = eng.counters().sum(0)
    cd = counter_dict(cx.sum_over_ranks((c1 - c0).tolist()))
    env_steps = cd["decided"]
    assert env_steps == world * n_envs * chunk * K, (env_steps, world, n_envs, chunk, K)
    assert cd["errors"] == 0
    value = env_steps / (elapsed_ms * 1e-3)
    avg_launch_s = (sum(launch_ms) / len(launch_ms)) * 1e-3
    kname = f"k_step_policy<{N_SLOTS},6,5,first_fit> (+ k_count_decisions, <1 % of the launch)"
    roofline = make_roofline(cx, cd, N_SLOTS, n_envs * chunk, avg_launch_s, kname)
    # DRAM traffic and instruction counts come from the committed ncu capture -- only if it was taken on THIS kernel
    # (profiles/traffic.json carries the hash of the CUDA sources it profiled); otherwise they are not claimed
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    issue, src_hash = None, kernel_source_hash()
    try:
        tj = json.load(open(tp))
        if tj.get("kernel_source_sha256") == src_hash and (TOPOLOGY, N_SLOTS) == ("nobel-eu", 320):
            roofline["traffic"] = float(tj["dram_bytes_per_env_step"]) * n_envs * chunk
            roofline["traffic_source"] = tj.get("source")
            ipe = float(tj["warp_instructions_per_env_step"])
            sm_clock = float(clocks.get("sm_mhz") or cx.peaks.get("sm_max_mhz", 1965.0)) * 1e6
            n_sm = torch.cuda.get_device_properties(cx.local).multi_processor_count
            issue = {"bound": "issue", "achieved": ipe * (n_envs * chunk) / avg_launch_s, "peak": n_sm * 4 * sm_clock,
                     "unit": "warp-instructions/s", "frac": ipe * (n_envs * chunk) / avg_launch_s / (n_sm * 4 * sm_clock),
                     "warp_instructions_per_env_step": ipe, "source": tj.get("source")}
        else:
            roofline["traffic_note"] = "profiles/traffic.json was captured on other kernel sources: not claimed"
    except Exception:
        pass
    roofline["issue_roofline"] = issue
    roofline["kernel_source_sha256"] = src_hash

    # ---- parity sample of the timed state: envs replayed through the oracle from reset
    parity = checker.replay_first_fit(tb, eng, checker.spread_sample(n_envs, args.parity_envs), PREFILL + (K + Wm) * chunk)
    parity = merge_parity(cx, dict(parity, checker="oracle/qrmsa_oracle.c"))

    # ---- phase B: end to end through the host-buffer C-ABI calls, one whole episode.  The public call is
    # PipelinedEpisodes.run: env slices on separate contexts/streams so that upload, kernels and download overlap.
    e2e = None
    if not args.no_e2e:
        from optical_networking_gym_b200.pipeline import PipelinedEpisodes

        out_actions = torch.empty((n_req - 1, n_envs), dtype=torch.int32, pin_memory=True)
        pipe = PipelinedEpisodes(tb, n_envs, n_req, slices=args.e2e_slices, device=cx.local)

        def episode():
            cnt = pipe.run(pinned, out_actions, launch_steps=512)   # reset + H2D + schedule + steps + D2H + counters
            return cx.sum_over_ranks(cnt.sum(0).tolist())           # the episode-end all-reduce of the counters

        episode()          # untimed warm-up episode (first-touch, staging buffers, NCCL communicator)
        cx.barrier()
        e0, e1 = cx.events(2)
        t0 = time.perf_counter()
        e0.record(stream)
        cnt = episode()
        e1.record(stream)
        cx.barrier()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) * 1e-3
        t_e2e = cx.max_over_ranks(max(wall, dev_s))
        assert int(cnt[0]) == world * n_envs * (n_req - 1)
        pipe.close()
        h2d = 11 * n_envs * n_req
        e2e = {"value": world * n_envs * (n_req - 1) / t_e2e, "unit": UNIT,
               "h2d_bytes_per_step": 11 * n_envs * chunk, "d2h_bytes_per_step": 4 * n_envs * chunk,
               "episode_requests": n_req, "seconds": t_e2e, "slices": args.e2e_slices,
               "h2d_gbs_per_rank": h2d / t_e2e / 1e9,
               "note": "whole episode from reset (empty network): trace upload from pinned host memory, schedule "
                       "build, every step, download of every decision and the counters, their all-reduce over the "
                       "ranks, through the host-buffer C-ABI calls on env slices overlapped over CUDA streams; bytes "
                       "are per bench step of `chunk` requests per env; h2d_gbs_per_rank = trace bytes / episode time "
                       f"(the host side -- generating the replay streams -- took {t_gen:.1f}s and is outside)"}
        # the host link alone: the same trace bytes copied from pinned memory with nothing else running
        dst = torch.empty(64 << 20, dtype=torch.uint8, device=cx.dev)
        srcp = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
        cx.barrier()
        a, b = cx.events(2)
        a.record(stream)
        for _ in range(8):
            dst.copy_(srcp, non_blocking=True)
        b.record(stream)
        cx.barrier()
        e2e["h2d_link_gbs_per_rank_all_ranks_copying"] = 8 * (64 << 20) / (cx.max_over_ranks(a.elapsed_time(b)) * 1e-3) / 1e9
        del dst, srcp, out_actions

    # ---- phase C (informational): the same episode with the requests drawn on the device (qrmsa_generate_trace,
    # Philox streams) instead of uploaded: no host generation, no H2D; decisions still downloaded
    e2e_dev = None
    if not args.no_e2e:
        out_words = torch.empty((n_req - 1, n_envs), dtype=torch.int32, pin_memory=True)

        def episode_dev():
            eng.reset()
            eng.generate_trace(n_req, LOAD, seed=BASE_SEED, env_offset=rank * n_envs)
            done = 0
            while done < n_req - 1:
                c = min(512, n_req - 1 - done)
                eng.step_first_fit(c)
                done += c
            eng.actions_host_strided(0, n_req - 1, out_words.data_ptr(), n_envs)
            return eng.counters().sum(0)

        episode_dev()
        cx.barrier()
        t0 = time.perf_counter()
        cnt = episode_dev()
        cx.barrier()
        t_dev = cx.max_over_ranks(time.perf_counter() - t0)
        assert int(cnt[0]) == n_envs * (n_req - 1)
        e2e_dev = {"value": world * n_envs * (n_req - 1) / t_dev, "unit": UNIT, "seconds": t_dev,
                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n_envs * chunk,
                   "note": "requests generated on the device (Philox), decisions downloaded; not the contract's e2e"}
        del out_words
    eng.close()
    del pinned, trace

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n_envs), "chunk_requests_per_env_per_step": chunk,
                   "episode_requests": n_req, "l2_policy": "inputs larger than L2 (per-GPU env state + trace "
                   f"touched per step >= {n_envs * (tb.n_links * 64 + 16 * chunk) / 1e6:.0f} MB)",
                   "trace": f"CPython-random-exact streams, seeds {BASE_SEED}+i, generated on host in {t_gen:.1f}s"},
        "roofline": roofline, "parity_sample": parity, "cpu_baseline": None, "e2e": e2e, "e2e_device_requests": e2e_dev,
        "gpu_launches": 2 * K, "clocks": clocks,
    }
    return line


# ------------------------------------------------------------------------------------------------
# C3: JOCN-style load sweep, 1 M envs sharded over the GPUs (strong scaling), per-load counters all-reduced
# ------------------------------------------------------------------------------------------------
def run_c3(cx, args):
    import numpy as np

    from optical_networking_gym_b200 import sharding
    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.pipeline import PipelinedEpisodes
    from optical_networking_gym_b200.tables import StaticTables
    from optical_networking_gym_b200.tracegen import TraceGenerator
    from oracle import checker

    torch, rank, world = cx.torch, cx.rank, cx.world
    loads = [100.0, 200.0, 300.0, 400.0, 500.0]          # graph_load.py:18-19
    L = 1000                                              # requests per episode (README command) -> 999 steps
    n_launch = 3                                          # 333 requests per env per launch
    per_total = args.c3_envs_per_load
    r0, r1 = sharding.shard_range(per_total, rank, world)
    per = r1 - r0
    n_envs = len(loads) * per
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", "tables_nobel-eu_320.npz"))
    load_vec = np.repeat(np.asarray(loads, np.float64), per)
    eng = Engine(tb, n_envs, L, device=cx.local)
    eng.set_groups(len(loads))
    stream = cx.stream()
    steps = [(L - 1) // n_launch] * n_launch
    steps[-1] += (L - 1) - sum(steps)

    def prepare():
        eng.reset()
        eng.generate_trace(L, load_vec, seed=BASE_SEED + 3, env_offset=rank * n_envs)

    def sweep(ev=None):
        for i, s in enumerate(steps):
            eng.step_first_fit(s)
            if ev is not None:
                ev[i + 1].record(stream)
        c = eng.counters_tensor().clone()
        if world > 1:
            cx.dist.all_reduce(c, op=cx.dist.ReduceOp.SUM)   # the path's only collective: per-load counters
        return c

    prepare(); sweep()                                   # warm-up episode (3 launches + the collective)
    prepare()
    cx.barrier()
    ev = cx.events(n_launch + 2)
    ev[0].record(stream)
    c = sweep(ev)
    ev[n_launch + 1].record(stream)
    cx.barrier()
    elapsed_ms = cx.max_over_ranks(ev[0].elapsed_time(ev[n_launch + 1]))
    launch_s = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(n_launch)) * 1e-3 / n_launch
    counters = c.cpu().numpy()                           # [loads][32], whole job
    cd = counter_dict(counters.sum(0))
    total_steps = len(loads) * per_total * (L - 1)
    assert cd["decided"] == total_steps and cd["errors"] == 0, (cd["decided"], total_steps)
    own = counter_dict(eng.counters().sum(0))            # this rank's launches, for the per-GPU roofline
    roofline = make_roofline(cx, own, 320, n_envs * (L - 1) / n_launch, launch_s, "k_step_policy<320,6,5,first_fit>")
    blocking = {}
    for i, l in enumerate(loads):
        p = float(counters[i, 0] - counters[i, 1]) / float(counters[i, 0])
        blocking[str(int(l))] = {"service_blocking_rate": p, "ci95": 1.96 * (p * (1 - p) / float(counters[i, 0])) ** 0.5,
                                 "bit_rate_blocking_rate": float(counters[i, 3] - counters[i, 4]) / float(counters[i, 3])}
    k_per_load = -(-args.parity_envs // len(loads))        # every load point is sampled
    sample = [g * per + j for g in range(len(loads)) for j in checker.spread_sample(per, k_per_load)]
    parity = merge_parity(cx, dict(checker.replay_first_fit(tb, eng, sample, L - 1), checker="oracle/qrmsa_oracle.c"))
    eng.close()

    # e2e: CPython-exact replay streams in pinned host memory, uploaded by env slices; only the counters come back
    e2e = None
    if not args.no_e2e:
        t_gen = time.time()
        pinned = cx.pinned_trace(L, n_envs)
        gen = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, load_vec, base_seed=BASE_SEED + rank * n_envs)
        gen.next(L, out=[p.numpy() for p in pinned])
        gen.close()
        t_gen = time.time() - t_gen
        spl = 8 if per % 8 == 0 else 1                   # slices per load point: a slice never straddles two loads
        pipe = PipelinedEpisodes(tb, n_envs, L, slices=len(loads) * spl, device=cx.local)

        def episode():
            each = pipe.run(pinned, None, launch_steps=512, per_slice=True)       # [slices][1][32]
            per_load = each[:, 0, :].reshape(len(loads), spl, -1).sum(1)
            return sharding.allreduce_counters(per_load)

        episode()
        cx.barrier()
        t0 = time.perf_counter()
        cnt = episode()
        cx.barrier()
        t_e2e = cx.max_over_ranks(time.perf_counter() - t0)
        assert int(cnt[:, 0].sum()) == total_steps
        pipe.close()
        e2e = {"value": total_steps / t_e2e, "unit": UNIT, "seconds": t_e2e,
               "h2d_bytes_per_step": 11 * n_envs * L // n_launch, "d2h_bytes_per_step": len(loads) * 32 * 8,
               "note": f"whole sweep from reset through the host-buffer C-ABI calls: {n_envs} envs per rank x {L} "
                       f"CPython-exact requests uploaded from pinned memory on {len(loads) * spl} env slices, schedule "
                       f"build, 999 steps, per-load counters back and all-reduced (host generation {t_gen:.1f}s, outside)"}
        del pinned
    return {"workload": f"JOCN-style load sweep, nobel-eu/320/k=5 first-fit, loads {[int(l) for l in loads]} x "
                        f"{per_total} envs = {len(loads) * per_total} envs over {world} GPU(s), episodes of {L} requests "
                        "from the empty network, requests drawn on the device (Philox)",
            "scaling": "strong", "value": total_steps / (elapsed_ms * 1e-3), "unit": UNIT, "seconds": elapsed_ms * 1e-3,
            "steps": n_launch, "warmup": n_launch, "envs_per_gpu": n_envs, "gpu_launches": 2 * n_launch,
            "timed_region": "3 launches of 333 requests per env + the all-reduce of the per-load counter matrix",
            "roofline": roofline, "e2e": e2e, "parity_sample": parity, "blocking": blocking}


# ------------------------------------------------------------------------------------------------
# C4: germany50 / 640 slots / load 800, 524,288 envs per GPU
# ------------------------------------------------------------------------------------------------
def run_c4(cx, args):
    import numpy as np

    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.pipeline import PipelinedEpisodes
    from optical_networking_gym_b200.tables import StaticTables
    from optical_networking_gym_b200.tracegen import TraceGenerator
    from oracle import checker

    torch, rank, world = cx.torch, cx.rank, cx.world
    n_envs, load, S = args.c4_envs, 800.0, 640
    prefill, chunk, K, Wm = 2400, 128, 8, 3             # 2,400 requests = 3 mean holding times of arrivals at load 800
    n_req = prefill + (K + Wm) * chunk + 1
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", "tables_germany50_640.npz"))
    eng = Engine(tb, n_envs, n_req, device=cx.local)
    stream = cx.stream()
    eng.reset()
    eng.generate_trace(n_req, load, seed=BASE_SEED + 4, env_offset=rank * n_envs)
    done = 0
    while done < prefill:
        eng.step_first_fit(min(600, prefill - done)); done += min(600, prefill - done)
    for _ in range(Wm):
        eng.step_first_fit(chunk)
    cx.barrier()
    c0 = eng.counters().sum(0)
    ev = cx.events(K + 1)
    ev[0].record(stream)
    for i in range(K):
        eng.step_first_fit(chunk)
        ev[i + 1].record(stream)
    cx.barrier()
    elapsed_ms = cx.max_over_ranks(ev[0].elapsed_time(ev[K]))
    launch_s = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(K)) * 1e-3 / K
    delta = eng.counters().sum(0) - c0
    own = counter_dict(delta)                             # this rank's launches, for the per-GPU roofline
    cd = counter_dict(cx.sum_over_ranks(delta.tolist()))
    assert cd["decided"] == world * n_envs * chunk * K and cd["errors"] == 0
    roofline = make_roofline(cx, own, S, n_envs * chunk, launch_s, "k_step_policy<640,6,5,first_fit>")
    parity = checker.replay_first_fit(tb, eng, checker.spread_sample(n_envs, args.parity_envs), prefill + (K + Wm) * chunk)
    parity = merge_parity(cx, dict(parity, checker="oracle/qrmsa_oracle.c"))
    state_gb = n_envs * (tb.n_links * (32 * 4 + 320 * 4 + 320 * 2) + 24 * n_req) / 1e9
    eng.close()

    e2e = None
    if not args.no_e2e:
        L = 1000
        t_gen = time.time()
        pinned = cx.pinned_trace(L, n_envs)
        gen = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, load, base_seed=BASE_SEED + rank * n_envs)
        gen.next(L, out=[p.numpy() for p in pinned])
        gen.close()
        t_gen = time.time() - t_gen
        out_actions = torch.empty((L - 1, n_envs), dtype=torch.int32, pin_memory=True)
        pipe = PipelinedEpisodes(tb, n_envs, L, slices=32, device=cx.local)

        def episode():
            return cx.sum_over_ranks(pipe.run(pinned, out_actions, launch_steps=512).sum(0).tolist())

        episode()
        cx.barrier()
        t0 = time.perf_counter()
        cnt = episode()
        cx.barrier()
        t_e2e = cx.max_over_ranks(time.perf_counter() - t0)
        assert int(cnt[0]) == world * n_envs * (L - 1)
        pipe.close()
        e2e = {"value": world * n_envs * (L - 1) / t_e2e, "unit": UNIT, "seconds": t_e2e,
               "h2d_bytes_per_step": 11 * n_envs * chunk, "d2h_bytes_per_step": 4 * n_envs * chunk,
               "note": f"whole {L}-request episodes from reset (the reference README's episode length) through the "
                       "host-buffer C-ABI calls on 32 env slices: CPython-exact streams uploaded from pinned memory, "
                       "schedule build, every step, every decision and the counters back, counters all-reduced; bytes per "
                       f"{chunk} requests per env (host generation {t_gen:.1f}s, outside)"}
        del pinned, out_actions
    return {"workload": f"germany50/640-slot (8 THz)/k=5/6-mod first-fit, load 800, launch 1 dBm, {n_envs} envs per GPU "
                        f"({world * n_envs} envs), steady state after a {prefill}-request prefill, requests drawn on the "
                        "device (Philox)",
            "scaling": "weak", "value": cd["decided"] / (elapsed_ms * 1e-3), "unit": UNIT, "ms_per_step": elapsed_ms / K,
            "steps": K, "warmup": Wm, "chunk_requests_per_env_per_step": chunk, "env_state_gb_per_gpu": state_gb,
            "gpu_launches": 2 * K, "roofline": roofline, "e2e": e2e, "parity_sample": parity}


# ------------------------------------------------------------------------------------------------
# C5: PPO rollout with on-device observation + action mask and a torch policy in the loop
# ------------------------------------------------------------------------------------------------
def run_c5(cx, args):
    import numpy as np

    from examples.ppo_rollout import make_policy, rollout
    from optical_networking_gym_b200.env import BatchedQRMSAEnv
    from optical_networking_gym_b200.tables import StaticTables
    from oracle import checker

    torch, rank, world = cx.torch, cx.rank, cx.world
    n_envs, n_steps, Wm = args.c5_envs, args.c5_steps, 3
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", "tables_nsfnet_320.npz"))
    env = BatchedQRMSAEnv(tb, n_envs, num_spectrum_resources=320, episode_length=n_steps + Wm + 1, load=210.0,
                          bit_rates=BIT_RATES, launch_power_dbm=1.0, gen_observation=True, seed=10 + rank * n_envs,
                          device=cx.local, reset=False)
    policy = make_policy(env.observation_space.shape[0], env.action_space.n, cx.dev, seed=rank)
    stream = cx.stream()
    buf = {"action": torch.zeros((n_steps + Wm, n_envs), dtype=torch.int64, device=cx.dev),
           "status": torch.zeros((n_steps + Wm, n_envs), dtype=torch.uint8, device=cx.dev),
           "reward": torch.zeros((n_steps + Wm, n_envs), dtype=torch.float32, device=cx.dev)}

    def sub(b, a, n):
        return {k: v[a:a + n] for k, v in b.items()}

    env.reset()
    rollout(env, policy, Wm, sub(buf, 0, Wm), seed=1000 + rank)   # warm-up steps of the same episode
    cx.barrier()
    e0, e1 = cx.events(2)
    e0.record(stream)
    rollout(env, policy, n_steps, sub(buf, Wm, n_steps), seed=1000 + rank, first_step=Wm)
    c = env.engine.counters_tensor().clone()
    if world > 1:
        cx.dist.all_reduce(c, op=cx.dist.ReduceOp.SUM)
    e1.record(stream)
    cx.barrier()
    elapsed_ms = cx.max_over_ranks(e0.elapsed_time(e1))
    cnt = c.cpu().numpy().sum(0)
    # a masked action is accepted or is the reject action -- except on the mask's own rounding edge: the reference
    # validates a candidate when round((gsnr - thr) / |thr|, 10) >= 0 (osnr.pyx:366), i.e. down to ~1e-9 dB BELOW the
    # threshold, where its step() then raises ValueError (qrmsa.pyx:925-929).  Such steps are counted, not hidden.
    st = buf["status"][Wm:]
    status_counts = cx.sum_over_ranks(torch.bincount(st.flatten().to(torch.int64), minlength=5)[:5].tolist())
    # parity: the sampled envs' whole episode (warm-up + timed steps) through the oracle's step(), then the final
    # observation and action mask
    sample = checker.spread_sample(n_envs, args.parity_envs)
    idx = torch.tensor(sample, device=cx.dev)
    acts = buf["action"][:, idx].cpu().numpy()
    stats = buf["status"][:, idx].cpu().numpy()
    fo, fm = env._obs[idx].cpu().numpy(), env.action_masks()[idx].cpu().numpy()
    parity = checker.replay_actions(tb, env.engine, sample, acts, stats, n_steps + Wm + 1, fo, fm)
    parity = merge_parity(cx, dict(parity, checker="oracle/qrmsa_oracle.c (step, observation, mask)"))

    # e2e: the same rollout where every step's rewards are read back to pinned host memory and the episode's request
    # streams are uploaded from pinned host memory by reset()
    e2e = None
    if not args.no_e2e:
        host_reward = torch.empty((n_steps, n_envs), dtype=torch.float32, pin_memory=True)
        cx.barrier()
        t0 = time.perf_counter()
        env.reset()                                       # next episode: host streams -> device (H2D), schedule build
        rollout(env, policy, n_steps, None, host_reward, seed=2000 + rank)
        cc = env.engine.counters_tensor().clone()
        if world > 1:
            cx.dist.all_reduce(cc, op=cx.dist.ReduceOp.SUM)
        cx.barrier()
        t_e2e = cx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n_envs * n_steps / t_e2e, "unit": UNIT, "seconds": t_e2e,
               "h2d_bytes_per_step": 11 * n_envs, "d2h_bytes_per_step": 4 * n_envs,
               "note": "reset() (CPython-exact request streams generated on the host and uploaded from pinned memory, "
                       "schedule build) + the rollout with each step's rewards copied to pinned host memory + the "
                       "counters' all-reduce; actions are born on the device (the policy runs there)"}
        del host_reward
    obs_dim, n_act = env.observation_space.shape[0], env.action_space.n
    env.close()
    per_step = n_envs * (obs_dim * 4 + n_act)             # bytes written by the observation kernel per rollout step
    return {"workload": f"PPO-style rollout, NSFNET/320/k=5, load 210, {n_envs} envs per GPU x {n_steps} steps, observation "
                        f"f32[{obs_dim}] + action mask u8[{n_act}] built on the device every step, torch MLP "
                        f"{obs_dim}-512-256-128-{n_act} (bf16), masked categorical sample by qrmsa_sample_masked_actions, rollout buffer on the device",
            "scaling": "weak", "value": world * n_envs * n_steps / (elapsed_ms * 1e-3), "unit": UNIT,
            "ms_per_step": elapsed_ms / n_steps, "steps": n_steps, "warmup": Wm, "gpu_launches": 4 * n_steps,
            "accepted": int(cnt[1]), "decided": int(cnt[0]),
            "step_status_counts": {"accepted": int(status_counts[0]), "reject_action": int(status_counts[1]),
                                   "not_free": int(status_counts[2]), "low_gsnr_on_mask_rounding_edge": int(status_counts[3]),
                                   "idle": int(status_counts[4])},
            "roofline": {"bound": "hbm", "achieved": per_step * n_steps / (elapsed_ms * 1e-3) / 1e9, "peak": cx.peak,
                         "unit": "GB/s", "frac": per_step * n_steps / (elapsed_ms * 1e-3) / 1e9 / cx.peak, "traffic": None,
                         "kernel": "k_observation_links + k_step_action + k_sample_masked (+ the policy's four torch GEMMs inside the step)",
                         "algorithmic_bytes_per_env_step": per_step / n_envs, "peak_source": cx.peak_src},
            "e2e": e2e, "parity_sample": parity}


# ------------------------------------------------------------------------------------------------
# C1: single env, NSFNET, 10,000 step()s -- the compiled reference on ONE core, the same trace on the device
# ------------------------------------------------------------------------------------------------
class C1Background:
    """Runs the single-core reference leg in a child process while the GPU legs run (it needs one host core for
    about a minute); joined before anything that times host cores."""

    def __init__(self, n_steps=10_000):
        self.n_steps = n_steps
        self.path = os.path.join(tempfile.mkdtemp(prefix="qrmsa_c1_"), "c1.npz")
        self.proc = None
        try:
            from oracle import ref_harness

            if ref_harness.available():
                code = (f"import sys; sys.path.insert(0, {ROOT!r}); from oracle import ref_bench; "
                        f"ref_bench.c1_single_core({self.path!r}, {n_steps})")
                self.proc = subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        except Exception:
            self.proc = None

    def finish(self, cx, timeout=600):
        import numpy as np

        from optical_networking_gym_b200 import _lib
        from optical_networking_gym_b200.engine import Engine, unpack_bitmaps
        from optical_networking_gym_b200.tables import StaticTables
        from oracle import ref_bench

        if self.proc is None:
            return {"workload": "C1", "unavailable": "oracle/_ref is not built on this box"}
        try:
            _, err = self.proc.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            return {"workload": "C1", "unavailable": "single-core reference run exceeded its time limit"}
        if self.proc.returncode != 0:
            return {"workload": "C1", "unavailable": "reference run failed: " + err.decode()[-300:]}
        g = np.load(self.path)
        n = self.n_steps
        tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", "tables_nsfnet_320.npz"))
        eng = Engine(tb, 1, n + 1, device=cx.local)
        eng.reset()
        eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in ("src", "dst", "rate", "arrival", "holding")])
        stream = cx.stream()
        eng.step_first_fit(3)                                # warm-up launches are part of the episode
        e0, e1 = cx.events(2)
        e0.record(stream)
        eng.step_first_fit(n - 3)
        e1.record(stream)
        cx.torch.cuda.synchronize()
        dev_s = e0.elapsed_time(e1) * 1e-3
        words = eng.actions_host(0, n).view(np.uint32)[:, 0]
        act = (words & _lib.ACTION_MASK).astype(np.int64)
        d = np.flatnonzero(act != g["action"])
        mism = exc = bm = 0
        if len(d):
            if words[int(d[0])] & _lib.FLAG_NEAR_THRESHOLD:
                exc = 1
            else:
                mism = 1
        elif not np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], g["final_slots"]):
            bm = 1
        eng.close()
        secs = float(g["seconds"])
        return {"workload": "single env, NSFNET (nsfnet_chen), k=5, 320 slots, first-fit, load 300, 10,000 step() calls, seed 50",
                "cpu_reference": {"value": n / secs, "unit": UNIT, "cores": 1, "kind": "reference", "seconds": secs,
                                  "cpu_model": ref_bench.cpu_model(),
                                  "sample": "the compiled reference's own loop heuristic(env); env.step(a), one process, "
                                            "run beside the GPU legs of this bench"},
                "value": (n - 3) / dev_s, "unit": UNIT, "seconds": dev_s,
                "note": "ONE env on the device = one warp of one SM: the latency-bound corner, reported for completeness",
                "parity_sample": {"envs": 1, "steps": n, "mismatches": mism, "excused": exc, "bitmap_mismatches": bm,
                                  "checker": "the compiled reference itself (oracle/_ref), same trace"}}


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.lib:
        from optical_networking_gym_b200 import _lib

        _lib.LIB_PATH = os.path.abspath(args.lib)
        _lib.needs_build = lambda: False
    cx = Cx(rank, local_rank, world)
    want = [] if args.configs.lower() == "none" else [c.strip().upper() for c in args.configs.split(",") if c.strip()]
    c1 = C1Background() if ("C1" in want and rank == 0 and world == 1) else None
    t_all = time.time()
    line = run_headline(cx, args)
    configs, timings = {}, {"headline_s": time.time() - t_all}
    for name, fn in (("C3", run_c3), ("C4", run_c4), ("C5", run_c5)):
        if name not in want:
            continue
        t0 = time.time()
        try:
            res = fn(cx, args)
        except AssertionError:
            raise
        except Exception as ex:   # a config that cannot run on this box (memory, host RAM) is reported, never hidden
            res = {"unavailable": repr(ex)[:300]}
            torch.cuda.empty_cache()
        timings[name + "_s"] = time.time() - t0
        configs[name] = res
    if c1 is not None:
        t0 = time.time()
        configs["C1"] = c1.finish(cx)
        timings["C1_wait_s"] = time.time() - t0

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
            cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample, "seconds": secs,
                   "cpu_model": ref_bench.cpu_model()}
            if "C1" in configs and "cpu_reference" in configs["C1"]:
                cpu["config1_single_core"] = configs["C1"]["cpu_reference"]
        except Exception as ex:  # the baseline is reported, never a gate
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    if rank == 0:
        line["cpu_baseline"] = cpu
        line["configs"] = configs
        line["bench_seconds"] = dict(timings, total_s=time.time() - t_all)
        print(json.dumps(line), flush=True)
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
