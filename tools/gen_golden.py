#!/usr/bin/env python
"""Generate tests/golden/* from the COMPILED reference (oracle/_ref).

Runs only where /root/reference exists (this container); the fixtures it writes
are committed so the GPU box and later sessions need neither the reference nor
the 1-minute Cython build.  Everything written here is OUTPUT of the reference
(request traces, decisions, GSNR values, slot matrices) or tables derived from
its topology object -- no reference source.

    python oracle/build_ref.py && python tools/gen_golden.py [--only NAME]

Fixtures
  tables_<topo>_<S>.npz      static tables exported from the reference topology object
  run_<tag>.npz              one env, first-fit heuristic + step (graph_load.py:161-163 loop):
                             trace, action/accepted/gsnr per step, every QoT check, slot snapshots
  multi_<tag>.npz            many short envs (seeds base..base+n-1) for batched parity
  rl_<tag>.npz               env.step() driven by externally chosen actions (valid, invalid, reject)
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from optical_networking_gym_b200.tables import StaticTables  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
BIT_RATES = (10, 40, 100, 400, 1000)

SINGLE = [
    # tag, topology, S, load, seed, steps, snapshot_every
    ("nobel-eu_320_l300_s50", "nobel-eu", 320, 300.0, 50, 3000, 500),
    ("nsfnet_320_l300_s50", "nsfnet", 320, 300.0, 50, 10000, 2000),   # BASELINE config 1
    ("germany50_640_l800_s52", "germany50", 640, 800.0, 52, 1500, 500),
    ("nobel-eu_320_l500_s7", "nobel-eu", 320, 500.0, 7, 1500, 500),
    ("ring4_320_l60_s3", "ring4", 320, 60.0, 3, 600, 200),
]
POLICY = [
    # tag, topology, S, load, seed, steps, reference heuristic
    ("lb_nobel-eu_320_l400_s9", "nobel-eu", 320, 400.0, 9, 2000, "load_balancing_best_modulation"),
    ("lb_nsfnet_320_l300_s4", "nsfnet", 320, 300.0, 4, 1500, "load_balancing_best_modulation"),
    ("lbff_nobel-eu_320_l400_s13", "nobel-eu", 320, 400.0, 13, 2000, "heuristic_load_balancing_first_fit"),
    ("lbff_nsfnet_320_l300_s8", "nsfnet", 320, 300.0, 8, 1500, "heuristic_load_balancing_first_fit"),
]
EXHAUSTIVE = [
    # tag, topology, S, load, seed, steps, reference heuristic: every valid start of every (path, modulation) is QoT-checked
    # (about 2 s of reference time per step), so the recordings are short; per step they keep the number of checks
    # and the gap between the best and the second-best acceptable GSNR instead of the full QoT log
    ("hsnr_nsfnet_320_l300_s21", "nsfnet", 320, 300.0, 21, 150, "heuristic_highest_snr"),
    ("hsnr_nobel-eu_320_l400_s5", "nobel-eu", 320, 400.0, 5, 90, "heuristic_highest_snr"),
]
VARIANTS = [
    # first-fit runs away from the JOCN configuration: tag, topology, topology kwargs, env kwargs, seed, steps.
    # Each writes its own tables fixture (tables_<tag>.npz) because modulations / margin / power / k / spans differ.
    ("var_ondm_nsfnet", "nsfnet", dict(modulations="ONDM"), dict(n_slots=320, load=250.0, launch_power_dbm=0.0), 31, 2000),
    ("var_margin_nobel-eu", "nobel-eu", dict(), dict(n_slots=320, load=350.0, launch_power_dbm=-1.0, margin=1.5,
                                                    bit_rates=(40, 100, 400), bit_rate_probabilities=[0.5, 0.3, 0.2]), 32, 2000),
    ("var_k3_nsfnet_160", "nsfnet", dict(k_paths=3, max_span_km=100, att_db_km=0.22, nf_db=5.5),
     dict(n_slots=160, load=120.0, launch_power_dbm=2.0, k_paths=3, bit_rates=(10, 100, 400)), 33, 2000),
]
MULTI = [
    # tag, topology, S, load, base_seed, n_envs, steps
    ("nobel-eu_320_l300_b50", "nobel-eu", 320, 300.0, 50, 64, 400),
    ("germany50_640_l800_b50", "germany50", 640, 800.0, 50, 8, 300),
]
RL = [
    ("nsfnet_320_l210_s11", "nsfnet", 320, 210.0, 11, 400),
]
OBS = [
    # tag, topology, S, load, seed, steps   (gen_observation=True: ~3 s per step in the reference)
    ("nsfnet_320_l210_s21", "nsfnet", 320, 210.0, 21, 60),
    ("nobel-eu_320_l400_s8", "nobel-eu", 320, 400.0, 8, 24),   # (ring_4 has < k paths per pair: the reference itself raises there)
]


def tables_for(topo, S):
    return StaticTables.from_topology(topo, num_spectrum_resources=S, bit_rates=BIT_RATES, launch_power_dbm=1.0,
                                      margin=0.0, k_paths=5, modulations_to_consider=6)


def gen_rl(topo, tb, S, load, seed, n_steps):
    """Drive reference env.step() with a mix of first-fit, random-valid, invalid and reject actions."""
    heur = rh.first_fit_heuristic()
    env = rh.make_env(topo, seed, n_slots=S, load=load, episode_length=n_steps + 1)
    node_index = {n: i for i, n in enumerate(topo.graph["node_indices"])}
    rates = list(env.bit_rates)
    rng = np.random.default_rng(seed)
    rec = dict(src=[], dst=[], rate=[], arrival=[], holding=[], action=[], status=[], reward=[], gsnr=[], term=[])
    snaps = []

    def log_req(svc):
        rec["src"].append(node_index[svc.source]); rec["dst"].append(node_index[svc.destination])
        rec["rate"].append(rates.index(int(svc.bit_rate)))
        rec["arrival"].append(np.float32(svc.arrival_time)); rec["holding"].append(np.float32(svc.holding_time))

    log_req(env.current_service)
    n_actions = env.action_space.n
    consumed = 0
    while consumed < n_steps:
        kind = rng.integers(10)
        a_ff, _, _ = heur(env)
        if kind < 6:
            a = a_ff
        elif kind < 7:
            a = n_actions - 1
        else:
            a = int(rng.integers(n_actions - 1))  # arbitrary: usually not free or GSNR too low
        svc = env.current_service
        before = svc.service_id
        try:
            _, reward, term, _, info = env.step(int(a))
            if "osnr" not in info:      # early return: path/slot not free, request not consumed (qrmsa.pyx:886-897)
                status, g = 2, 0.0
            else:
                status, g = (1 if a == n_actions - 1 else 0), info["osnr"]
        except ValueError:
            # GSNR below threshold: the reference raises (qrmsa.pyx:925-929) and leaves the state untouched
            status, reward, term, g = 3, 0.0, False, float("nan")
        rec["action"].append(a); rec["status"].append(status); rec["reward"].append(reward)
        rec["gsnr"].append(g); rec["term"].append(bool(term))
        if status in (0, 1):
            consumed += 1
            assert env.current_service.service_id == before + 1
            log_req(env.current_service)
        else:
            assert env.current_service.service_id == before
        snaps.append(np.packbits(np.array(env.topology.graph["available_slots"], np.uint8), axis=1))
    out = {k: np.array(v) for k, v in rec.items()}
    out["src"] = out["src"].astype(np.uint8); out["dst"] = out["dst"].astype(np.uint8)
    out["rate"] = out["rate"].astype(np.uint8)
    out["arrival"] = out["arrival"].astype(np.float32); out["holding"] = out["holding"].astype(np.float32)
    out["action"] = out["action"].astype(np.int64); out["status"] = out["status"].astype(np.uint8)
    out["slots_packed_last"] = snaps[-1]
    out["slots_packed_every"] = np.array(snaps[::25])
    out["final_slots"] = np.array(env.topology.graph["available_slots"], np.uint8)
    return out


def gen_obs(topo, S, load, seed, n_steps, modulations_to_consider=6):
    """gen_observation=True (qrmsa.pyx:583-781): observation vector + GSNR-validated action mask per step,
    driven by a seeded mix of mask-sampled and first-fit actions."""
    heur = rh.first_fit_heuristic()
    # a few hundred first-fit steps without observations would be faster, but the env cannot switch modes;
    # a higher load fills the network within the recorded steps instead
    if modulations_to_consider == 6:
        env = rh.make_env(topo, seed, n_slots=S, load=load, episode_length=n_steps + 1, gen_observation=True)
    else:   # modulations_to_consider < len(modulations): max_modulation_idx moves with the request (qrmsa.pyx:543-581)
        _, ref_qrmsa, _, _ = rh.import_reference()
        kw = rh.env_kwargs(topo, n_slots=S, load=load, episode_length=n_steps + 1, gen_observation=True)
        kw["modulations_to_consider"] = modulations_to_consider
        with rh.seeded_random(seed):
            env = ref_qrmsa.QRMSAEnv(**kw)
    node_index = {n: i for i, n in enumerate(topo.graph["node_indices"])}
    rates = list(env.bit_rates)
    rng = np.random.default_rng(seed)
    obs0, info0 = env.reset()
    rec = dict(src=[], dst=[], rate=[], arrival=[], holding=[], action=[], reward=[], obs=[obs0], mask=[info0["mask"]],
               max_mod=[int(env.max_modulation_idx)])

    def log_req(svc):
        rec["src"].append(node_index[svc.source]); rec["dst"].append(node_index[svc.destination])
        rec["rate"].append(rates.index(int(svc.bit_rate)))
        rec["arrival"].append(np.float32(svc.arrival_time)); rec["holding"].append(np.float32(svc.holding_time))

    log_req(env.current_service)
    mask = info0["mask"]
    for t in range(n_steps):
        valid = np.flatnonzero(mask[:-1])
        if len(valid) and rng.integers(4) > 0:
            a = int(rng.choice(valid))
        elif len(valid):
            a = int(valid[0])
        else:
            a = len(mask) - 1
        obs, reward, term, _, info = env.step(a)
        mask = info["mask"]
        rec["action"].append(a); rec["reward"].append(reward); rec["obs"].append(obs); rec["mask"].append(mask)
        rec["max_mod"].append(int(env.max_modulation_idx))
        log_req(env.current_service)
    out = dict(src=np.array(rec["src"], np.uint8), dst=np.array(rec["dst"], np.uint8), rate=np.array(rec["rate"], np.uint8),
               arrival=np.array(rec["arrival"], np.float32), holding=np.array(rec["holding"], np.float32),
               action=np.array(rec["action"], np.int64), reward=np.array(rec["reward"], np.float64),
               obs=np.array(rec["obs"], np.float32), mask=np.packbits(np.array(rec["mask"], np.uint8), axis=1),
               n_actions=np.int64(len(mask)), max_mod=np.array(rec["max_mod"], np.int32),
               final_slots=np.array(env.topology.graph["available_slots"], np.uint8))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    topo_cache = {}

    def topo_of(name):
        if name not in topo_cache:
            topo_cache[name] = rh.make_topology(name)
        return topo_cache[name]

    for tag, name, S, load, seed, steps, snap in SINGLE:
        if args.only and args.only not in tag:
            continue
        t0 = time.time()
        topo = topo_of(name)
        tb = tables_for(topo, S)
        tb.save(os.path.join(GOLDEN, f"tables_{name}_{S}.npz"))
        out, _ = rh.run_first_fit(topo, seed, steps, snapshot_every=snap, n_slots=S, load=load)
        out["snap_slots"] = np.packbits(out["snap_slots"], axis=2)
        out["meta_load"] = np.float64(load); out["meta_seed"] = np.int64(seed)
        np.savez_compressed(os.path.join(GOLDEN, f"run_{tag}.npz"), **out)
        print(f"run_{tag}: {steps} steps, accept {out['accepted'].mean():.3f}, {time.time() - t0:.1f}s")

    for tag, name, S, load, seed, steps, hname in POLICY:
        if args.only and args.only not in ("policy_" + tag):
            continue
        t0 = time.time()
        topo = topo_of(name)
        tables_for(topo, S).save(os.path.join(GOLDEN, f"tables_{name}_{S}.npz"))
        out, _ = rh.run_first_fit(topo, seed, steps, heuristic_name=hname, n_slots=S, load=load)
        out["meta_load"] = np.float64(load); out["meta_seed"] = np.int64(seed)
        np.savez_compressed(os.path.join(GOLDEN, f"policy_{tag}.npz"), **out)
        print(f"policy_{tag}: {steps} steps, accept {out['accepted'].mean():.3f}, {time.time() - t0:.1f}s")

    for tag, name, tkw, ekw, seed, steps in VARIANTS:
        if args.only and args.only not in tag:
            continue
        t0 = time.time()
        tkw = dict(tkw)
        if tkw.get("modulations") == "ONDM":
            tkw["modulations"] = rh.ONDM_MODULATIONS
        topo = rh.make_topology(name, **tkw)
        tb = StaticTables.from_topology(topo, num_spectrum_resources=ekw["n_slots"],
                                        bit_rates=ekw.get("bit_rates", BIT_RATES),
                                        launch_power_dbm=ekw.get("launch_power_dbm", 1.0), margin=ekw.get("margin", 0.0),
                                        k_paths=ekw.get("k_paths", 5), modulations_to_consider=6)
        tb.save(os.path.join(GOLDEN, f"tables_{tag}.npz"))
        out, _ = rh.run_first_fit(topo, seed, steps, **ekw)
        out["meta_load"] = np.float64(ekw["load"]); out["meta_seed"] = np.int64(seed)
        np.savez_compressed(os.path.join(GOLDEN, f"run_{tag}.npz"), **out)
        print(f"run_{tag}: {steps} steps, accept {out['accepted'].mean():.3f}, {len(out['qot_gsnr'])} QoT checks, "
              f"{time.time() - t0:.1f}s")

    for tag, name, S, load, seed, steps, hname in EXHAUSTIVE:
        if args.only and args.only not in ("policy_" + tag):
            continue
        t0 = time.time()
        topo = topo_of(name)
        tables_for(topo, S).save(os.path.join(GOLDEN, f"tables_{name}_{S}.npz"))
        out, _ = rh.run_first_fit(topo, seed, steps, heuristic_name=hname, n_slots=S, load=load)
        qs, qg, qt = out.pop("qot_step"), out.pop("qot_gsnr"), out.pop("qot_thr")
        n_checks = np.bincount(qs, minlength=steps).astype(np.int32)
        gap = np.full(steps, np.inf)
        thr_margin = np.full(steps, np.inf)
        for t in range(steps):
            sel = qs == t
            g, th = qg[sel], qt[sel]
            if len(g):
                thr_margin[t] = np.abs(g - th).min()
            ok = np.sort(g[g >= th])[::-1]
            if len(ok) >= 2:
                gap[t] = ok[0] - ok[1]
        out.update(n_checks=n_checks, best_gap_db=gap, thr_margin_db=thr_margin,
                   meta_load=np.float64(load), meta_seed=np.int64(seed))
        np.savez_compressed(os.path.join(GOLDEN, f"policy_{tag}.npz"), **out)
        print(f"policy_{tag}: {steps} steps, accept {out['accepted'].mean():.3f}, min gap {gap.min():.3e} dB, "
              f"{time.time() - t0:.1f}s")

    for tag, name, S, load, base, n_envs, steps in MULTI:
        if args.only and args.only not in tag:
            continue
        t0 = time.time()
        topo = topo_of(name)
        tables_for(topo, S).save(os.path.join(GOLDEN, f"tables_{name}_{S}.npz"))
        keys = ("src", "dst", "rate", "arrival", "holding", "action", "accepted", "gsnr", "final_slots")
        acc = {k: [] for k in keys}
        qn, qg, qt, qs = [], [], [], []
        for i in range(n_envs):
            out, _ = rh.run_first_fit(topo, base + i, steps, n_slots=S, load=load)
            for k in keys:
                acc[k].append(out[k])
            qn.append(len(out["qot_gsnr"])); qg.append(out["qot_gsnr"]); qt.append(out["qot_thr"]); qs.append(out["qot_step"])
        d = {k: np.array(v) for k, v in acc.items()}
        d["final_slots"] = np.packbits(d["final_slots"], axis=2)
        d["qot_count"] = np.array(qn, np.int32)
        d["qot_gsnr"] = np.concatenate(qg); d["qot_thr"] = np.concatenate(qt); d["qot_step"] = np.concatenate(qs)
        d["meta_load"] = np.float64(load); d["meta_seed"] = np.int64(base)
        np.savez_compressed(os.path.join(GOLDEN, f"multi_{tag}.npz"), **d)
        print(f"multi_{tag}: {n_envs} envs x {steps} steps, {time.time() - t0:.1f}s")

    for tag, name, S, load, seed, steps in RL:
        if args.only and args.only not in tag:
            continue
        t0 = time.time()
        topo = topo_of(name)
        tb = tables_for(topo, S)
        tb.save(os.path.join(GOLDEN, f"tables_{name}_{S}.npz"))
        out = gen_rl(topo, tb, S, load, seed, steps)
        out["meta_load"] = np.float64(load); out["meta_seed"] = np.int64(seed)
        np.savez_compressed(os.path.join(GOLDEN, f"rl_{tag}.npz"), **out)
        st = out["status"]
        print(f"rl_{tag}: {len(st)} calls, status counts {np.bincount(st, minlength=4)}, {time.time() - t0:.1f}s")


    for tag, name, S, load, seed, steps in OBS:
        if args.only and args.only not in ("obs_" + tag):
            continue
        t0 = time.time()
        topo = topo_of(name)
        tables_for(topo, S).save(os.path.join(GOLDEN, f"tables_{name}_{S}.npz"))
        out = gen_obs(topo, S, load, seed, steps)
        out["meta_load"] = np.float64(load); out["meta_seed"] = np.int64(seed)
        np.savez_compressed(os.path.join(GOLDEN, f"obs_{tag}.npz"), **out)
        print(f"obs_{tag}: {steps} steps, mean valid actions {np.unpackbits(out['mask'], axis=1).sum(1).mean():.0f}, "
              f"{time.time() - t0:.1f}s")


    if not args.only or args.only in "csv_nsfnet":
        # per-service CSV written by the reference itself (qrmsa.pyx:387-406, :967-990)
        import glob, tempfile
        topo = topo_of("nsfnet")
        tmp = tempfile.mkdtemp()
        rh_kw = dict(n_slots=320, load=300.0)
        kw = rh.env_kwargs(topo, episode_length=401, **rh_kw)
        kw["file_name"] = os.path.join(tmp, "svc")
        _, ref_qrmsa, _, _ = rh.import_reference()
        with rh.seeded_random(77):
            env = ref_qrmsa.QRMSAEnv(**kw)
        heur = rh.first_fit_heuristic()
        node_index = {n: i for i, n in enumerate(topo.graph["node_indices"])}
        rates = list(env.bit_rates)
        tr = []
        def req(svc):
            return (node_index[svc.source], node_index[svc.destination], rates.index(int(svc.bit_rate)),
                    np.float32(svc.arrival_time), np.float32(svc.holding_time))
        tr.append(req(env.current_service))
        for t in range(400):
            a, _, _ = heur(env)
            env.step(a)
            tr.append(req(env.current_service))
        text = open(glob.glob(os.path.join(tmp, "*.csv"))[0]).read()
        arr = np.array(tr, dtype=[("src", "u1"), ("dst", "u1"), ("rate", "u1"), ("arrival", "f4"), ("holding", "f4")])
        np.savez_compressed(os.path.join(GOLDEN, "csv_nsfnet_320_l300_s77.npz"), csv=np.array(text),
                            **{k: arr[k].copy() for k in arr.dtype.names})
        print("csv_nsfnet_320_l300_s77:", len(text.splitlines()), "lines")


    if not args.only or args.only in "obs_mc2_nsfnet_320_l260_s5":
        t0 = time.time()
        out = gen_obs(topo_of("nsfnet"), 320, 260.0, 5, 40, modulations_to_consider=2)
        out["meta_load"] = np.float64(260.0); out["meta_seed"] = np.int64(5)
        np.savez_compressed(os.path.join(GOLDEN, "obs_mc2_nsfnet_320_l260_s5.npz"), **out)
        print(f"obs_mc2_nsfnet_320_l260_s5: 40 steps, max_modulation_idx values {sorted(set(out['max_mod'].tolist()))}, {time.time() - t0:.1f}s")

    # measure_disruptions / defragmentation (qrmsa.pyx:937-952, :1113-1122, :1545-1639), first-fit heuristic
    FEATS = [("feat_disrupt_nsfnet_320_l600_s7", "nsfnet", 600.0, 7, 400, dict(measure_disruptions=True)),
             ("feat_defrag_nsfnet_320_l300_s9_n5", "nsfnet", 300.0, 9, 500, dict(defragmentation=True, n_defrag_services=5)),
             ("feat_defrag_nsfnet_320_l150_s11_n0", "nsfnet", 150.0, 11, 250, dict(defragmentation=True, n_defrag_services=0)),
             ("feat_both_nobel-eu_320_l400_s13_n3", "nobel-eu", 400.0, 13, 300,
              dict(measure_disruptions=True, defragmentation=True, n_defrag_services=3))]
    for tag, name, load, seed, steps, feats in FEATS:
        if args.only and args.only not in tag:
            continue
        import glob, tempfile
        t0 = time.time()
        topo = topo_of(name)
        tmp = tempfile.mkdtemp()
        kw = rh.env_kwargs(topo, n_slots=320, load=load, episode_length=steps + 1)
        kw.update(feats)
        kw["file_name"] = os.path.join(tmp, "svc")      # the CSV carries the per-step disrupted count (qrmsa.pyx:983)
        _, ref_qrmsa, _, _ = rh.import_reference()
        with rh.seeded_random(seed):
            env = ref_qrmsa.QRMSAEnv(**kw)
        heur = rh.first_fit_heuristic()
        node_index = {n: i for i, n in enumerate(topo.graph["node_indices"])}
        rates = list(env.bit_rates)

        def req(svc):
            return (node_index[svc.source], node_index[svc.destination], rates.index(int(svc.bit_rate)),
                    np.float32(svc.arrival_time), np.float32(svc.holding_time))

        tr = [req(env.current_service)]
        act, accd, dis_ratio, realloc, cycles, acc_total = [], [], [], [], [], []
        for t in range(steps):
            a, _, _ = heur(env)
            svc = env.current_service
            _, _, _, _, info = env.step(a)
            act.append(a); accd.append(1 if svc.accepted else 0)
            dis_ratio.append(info["disrupted_services"]); realloc.append(info["episode_service_realocations"])
            cycles.append(info["episode_defrag_cicles"]); acc_total.append(info["episode_services_accepted"])
            tr.append(req(env.current_service))
        lines = open(glob.glob(os.path.join(tmp, "*.csv"))[0]).read().splitlines()[2:]
        local = np.array([int(l.split(",")[11]) if l.split(",")[4] != "-1" else 0 for l in lines], np.int32)
        arr = np.array(tr, dtype=[("src", "u1"), ("dst", "u1"), ("rate", "u1"), ("arrival", "f4"), ("holding", "f4")])
        cum = np.rint(np.array(dis_ratio) * np.maximum(np.array(acc_total), 1)).astype(np.int64)
        assert np.array_equal(np.cumsum(local), cum), "CSV column and info ratio disagree"
        np.savez_compressed(os.path.join(GOLDEN, f"{tag}.npz"), action=np.array(act, np.int64), accepted=np.array(accd, np.uint8),
                            disrupted_local=local, disrupted_ratio=np.array(dis_ratio), realocations=np.array(realloc, np.int64),
                            defrag_cicles=np.array(cycles, np.int64),
                            final_slots=np.array(env.topology.graph["available_slots"], dtype=np.uint8),
                            meta_load=np.float64(load), meta_seed=np.int64(seed),
                            feat=np.array([int(feats.get("measure_disruptions", False)), int(feats.get("defragmentation", False)),
                                           int(feats.get("n_defrag_services", 0))], np.int64),
                            **{k: arr[k].copy() for k in arr.dtype.names})
        print(f"{tag}: {steps} steps, accepted {sum(accd)}, disrupted {int(cum[-1])}, realocations {realloc[-1]}, "
              f"defrag cycles {cycles[-1]}, {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
