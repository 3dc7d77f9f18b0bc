"""Drive the COMPILED, UNMODIFIED reference (oracle/_ref) deterministically.   TEST INFRASTRUCTURE ONLY.

Used by tools/gen_golden.py (golden-vector generation, this container), by
tests/ (when oracle/_ref is present) and by `bench.py --impl reference` /
`cpu_baseline` (timing the reference's own Cython path on host cores).
Never imported by the product package.

Facts relied upon (reference file:line):
  * the env draws requests from `self.rng = random.Random()` (qrmsa.pyx:241),
    OS-entropy seeded; the `seed=` kwarg only feeds an unused numpy generator
    (qrmsa.pyx:347).  The module looks `random.Random` up at call time, so
    patching the stdlib attribute while the env is constructed pins the stream.
  * the benchmark loop is `action,_,_ = heuristic(env); env.step(action)`
    (examples/JOCN_Benchmark_2024/graph_load.py:161-163), heuristic =
    heuristic_shortest_available_path_first_fit_best_modulation
    (heuristics/heuristics.py:923).
  * JOCN modulation table: graph_load.py:252-295; get_topology arguments
    (80 km spans, 0.2 dB/km, NF 4.5 dB, k=5): graph_load.py:306-314.
"""
from __future__ import annotations

import contextlib
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUB_DIR = os.path.join(HERE, "gymnasium_stub")
TOPO_DIR = os.environ.get("QRMSA_TOPOLOGY_DIR") or (
    os.path.join(REF_DIR, "topologies") if os.path.isdir(os.path.join(REF_DIR, "topologies"))
    else "/root/reference/examples/topologies")

TOPOLOGY_FILES = {
    "nsfnet": "nsfnet_chen.txt",
    "nobel-eu": "nobel-eu.xml",
    "germany50": "germany50.xml",
    "ring4": "ring_4.txt",
    "nobel-us": "nobel-us.xml",
}

# (name, maximum_length, spectral_efficiency, minimum_osnr, inband_xt)  -- graph_load.py:252-295
JOCN_MODULATIONS = (
    ("BPSK", 100_000, 1, 3.71, -14),
    ("QPSK", 2_000, 2, 6.72, -17),
    ("8QAM", 1_000, 3, 10.84, -20),
    ("16QAM", 500, 4, 13.24, -23),
    ("32QAM", 250, 5, 16.16, -26),
    ("64QAM", 125, 6, 19.01, -29),
)


def available() -> bool:
    import glob

    return bool(glob.glob(os.path.join(REF_DIR, "optical_networking_gym", "envs", "qrmsa*.so")))


def _ensure_path():
    try:
        import gymnasium  # noqa: F401
    except ImportError:
        if STUB_DIR not in sys.path:
            sys.path.insert(0, STUB_DIR)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


def import_reference():
    """Returns (topology_module, qrmsa_module, heuristics_module, osnr_module)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built; run `python oracle/build_ref.py` where /root/reference exists")
    _ensure_path()
    from optical_networking_gym import topology as ref_topology
    from optical_networking_gym.envs import qrmsa as ref_qrmsa
    from optical_networking_gym.heuristics import heuristics as ref_heuristics
    from optical_networking_gym.core import osnr as ref_osnr

    return ref_topology, ref_qrmsa, ref_heuristics, ref_osnr


@contextlib.contextmanager
def seeded_random(seed: int):
    """While active, `random.Random()` (no args) is seeded with `seed` (qrmsa.pyx:241)."""
    orig = random.Random

    class _Seeded(orig):
        def __init__(self, x=None):
            super().__init__(seed if x is None else x)

    random.Random = _Seeded
    try:
        yield
    finally:
        random.Random = orig


# the ONDM 2025 set (reference test_qrmsa.py:11-54, examples/ONDM_2025): same spectral efficiencies, other thresholds
ONDM_MODULATIONS = (("BPSK", 100_000, 1, 12.6, -14), ("QPSK", 2_000, 2, 12.6, -17), ("8QAM", 1_000, 3, 18.6, -20),
                    ("16QAM", 500, 4, 22.4, -23), ("32QAM", 250, 5, 26.4, -26), ("64QAM", 125, 6, 30.4, -29))


def make_topology(name: str, k_paths: int = 5, max_span_km: float = 80, att_db_km: float = 0.2, nf_db: float = 4.5,
                  modulations=None):
    ref_topology, _, _, _ = import_reference()
    mods = tuple(
        ref_topology.Modulation(name=n, maximum_length=ml, spectral_efficiency=se, minimum_osnr=mo, inband_xt=xt)
        for (n, ml, se, mo, xt) in (modulations or JOCN_MODULATIONS)
    )
    path = os.path.join(TOPO_DIR, TOPOLOGY_FILES[name])
    return ref_topology.get_topology(path, None, mods, max_span_km, att_db_km, nf_db, k_paths)


def env_kwargs(topology, n_slots=320, load=300.0, episode_length=1000, launch_power_dbm=1.0, seed=50,
               bit_rates=(10, 40, 100, 400, 1000), margin=0.0, gen_observation=False, k_paths=5,
               bit_rate_probabilities=None):
    """The JOCN benchmark configuration (graph_load.py:316-336)."""
    return dict(
        topology=topology,
        seed=seed,
        allow_rejection=True,
        load=load,
        episode_length=episode_length,
        num_spectrum_resources=n_slots,
        launch_power_dbm=launch_power_dbm,
        bandwidth=n_slots * 12.5e9,
        frequency_start=3e8 / 1565e-9,
        frequency_slot_bandwidth=12.5e9,
        bit_rate_selection="discrete",
        bit_rates=bit_rates,
        bit_rate_probabilities=bit_rate_probabilities,
        margin=margin,
        measure_disruptions=False,
        file_name="",
        k_paths=k_paths,
        modulations_to_consider=6,
        defragmentation=False,
        n_defrag_services=0,
        gen_observation=gen_observation,
    )


def make_env(topology, rng_seed: int, **kw):
    """Reference QRMSAEnv whose request stream is random.Random(rng_seed).  The ctor calls reset()
    once (qrmsa.pyx:414-415), which consumes request 0 of the stream."""
    _, ref_qrmsa, _, _ = import_reference()
    with seeded_random(rng_seed):
        env = ref_qrmsa.QRMSAEnv(**env_kwargs(topology, **kw))
    return env


def first_fit_heuristic():
    _, _, h, _ = import_reference()
    return h.heuristic_shortest_available_path_first_fit_best_modulation


def run_first_fit(topology, rng_seed: int, n_steps: int, record: bool = True, snapshot_every: int = 0,
                  heuristic_name: str = "heuristic_shortest_available_path_first_fit_best_modulation", **kw):
    """Run `n_steps` of heuristic+step on one reference env (one episode, no intermediate reset).

    Returns a dict of numpy arrays:
      trace:  src,dst (node indices), rate_idx, arrival f32, holding f32 for requests 0..n_steps
      steps:  action, accepted, gsnr (info['osnr'], 0.0 on reject)
      qot:    every GSNR the heuristic evaluated: (step, gsnr, threshold)
      slots:  final available_slots int32[E][S]; optional snapshots
    """
    _, _, heur_mod, _ = import_reference()
    kw = dict(kw)
    kw["episode_length"] = n_steps + 1
    env = make_env(topology, rng_seed, **kw)
    heuristic = getattr(heur_mod, heuristic_name)
    node_index = {n: i for i, n in enumerate(topology.graph["node_indices"])}
    bit_rates = list(env.bit_rates)

    qot_log = []
    orig_osnr = heur_mod.calculate_osnr
    cur_step = [0]

    if record:
        def wrapped(e, service):
            out = orig_osnr(e, service)
            qot_log.append((cur_step[0], out[0], service.current_modulation.minimum_osnr + e.margin))
            return out

        heur_mod.calculate_osnr = wrapped

    def req_fields(svc):
        return (node_index[svc.source], node_index[svc.destination], bit_rates.index(int(svc.bit_rate)),
                np.float32(svc.arrival_time), np.float32(svc.holding_time))

    trace = [req_fields(env.current_service)]
    actions = np.zeros(n_steps, np.int64)
    accepted = np.zeros(n_steps, np.uint8)
    gsnr = np.zeros(n_steps, np.float64)
    snapshots, snap_steps = [], []
    try:
        for t in range(n_steps):
            cur_step[0] = t
            a, _, _ = heuristic(env)
            svc = env.current_service
            _, _, done, _, info = env.step(a)
            actions[t] = a
            accepted[t] = 1 if svc.accepted else 0
            gsnr[t] = info["osnr"]
            trace.append(req_fields(env.current_service))
            if snapshot_every and (t + 1) % snapshot_every == 0:
                snapshots.append(np.array(env.topology.graph["available_slots"], dtype=np.uint8))
                snap_steps.append(t + 1)
            assert done == (t == n_steps - 1)
    finally:
        heur_mod.calculate_osnr = orig_osnr
    tr = np.array(trace, dtype=[("src", "u1"), ("dst", "u1"), ("rate", "u1"), ("arrival", "f4"), ("holding", "f4")])
    out = dict(
        src=tr["src"].copy(), dst=tr["dst"].copy(), rate=tr["rate"].copy(),
        arrival=tr["arrival"].copy(), holding=tr["holding"].copy(),
        action=actions, accepted=accepted, gsnr=gsnr,
        final_slots=np.array(env.topology.graph["available_slots"], dtype=np.uint8),
    )
    if record:
        q = np.array(qot_log, dtype=np.float64).reshape(-1, 3)
        out.update(qot_step=q[:, 0].astype(np.int32), qot_gsnr=q[:, 1].copy(), qot_thr=q[:, 2].copy())
    if snapshot_every:
        out.update(snap_steps=np.array(snap_steps, np.int32), snap_slots=np.array(snapshots, np.uint8))
    return out, env
