"""The reference's OWN heuristics (unchanged, imported from oracle/_ref) driving the B200 `QRMSAEnv` side by side with the
compiled reference env: same constructor kwargs, same seed, same heuristic function object called on both
(optical_networking_gym_b200.compat.patch_reference_heuristics).  Also the constructor / reset surface that round 1
refused: `file_name=` (per-service CSV), `bit_rate_selection="continuous"`, `reset(options={"only_episode_counters":
True})`, and the three values of `calculate_osnr`."""
import os

import numpy as np
import pytest

from oracle import ref_harness as rh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")]


def _pair(seed, L, load=300.0, topo_name="nsfnet", **over):
    """(reference env, B200 env) with the same kwargs and the same request stream."""
    from optical_networking_gym_b200.env import QRMSAEnv

    _, ref_qrmsa, _, _ = rh.import_reference()
    topo = rh.make_topology(topo_name)
    kw = rh.env_kwargs(topo, n_slots=320, load=load, episode_length=L)
    kw.update(over)
    with rh.seeded_random(seed):
        ref = ref_qrmsa.QRMSAEnv(**kw)
    kw2 = dict(kw)
    ctor_seed = kw2.pop("seed")
    env = QRMSAEnv(seed=seed, **kw2)
    return ref, env, ctor_seed


def _same_request(cs, rs):
    assert (cs.source, cs.destination, cs.service_id) == (rs.source, rs.destination, rs.service_id)
    assert cs.bit_rate == rs.bit_rate and cs.arrival_time == rs.arrival_time and cs.holding_time == rs.holding_time


@pytest.mark.parametrize("name,steps", [("heuristic_shortest_available_path_first_fit_best_modulation", 150),
                                        ("heuristic_mscl_sequential_simplified", 120),
                                        ("best_modulation_load_balancing", 120),
                                        ("heuristic_load_balancing_first_fit", 80),
                                        ("heuristic_psr", 12),
                                        ("heuristic_highest_snr", 25),
                                        ("shortest_available_path_lowest_spectrum_best_modulation", 100),
                                        ("load_balancing_best_modulation", 80),
                                        # (heuristic_mscl itself takes minutes per request on the reference env: not run here)
                                        ("heuristic_mscl_simplified", 40),
                                        # exact fit takes a free block of exactly n slots without a guard slot: step() then
                                        # answers "not free" and does not consume the request (qrmsa.pyx:886-897), so the
                                        # reference's own loop stops advancing -- the B200 env must stall the same way
                                        ("heuristic_exact_fit", 40)])
def test_reference_heuristic_on_both_envs(name, steps):
    from optical_networking_gym_b200.compat import patch_reference_heuristics

    _, _, H, _ = rh.import_reference()
    patch_reference_heuristics(H)
    h = getattr(H, name)
    ref, env, _ = _pair(4242, steps + 1)
    n_acc = 0
    for t in range(steps):
        _same_request(env.current_service, ref.current_service)
        out_r, out_b = h(ref), h(env)
        a_r = out_r[0] if isinstance(out_r, tuple) else out_r
        a_b = out_b[0] if isinstance(out_b, tuple) else out_b
        assert a_b == a_r, f"{name}: step {t}: B200 env {out_b}, reference env {out_r}"
        if isinstance(out_r, tuple):
            assert tuple(out_b[1:]) == tuple(out_r[1:])
        o_r, o_b = ref.step(a_r), env.step(a_b)
        assert o_b[1] == o_r[1] and o_b[2] == o_r[2]
        assert ("osnr" in o_b[4]) == ("osnr" in o_r[4])   # (absent from the "not free" answer, qrmsa.pyx:893)
        if "osnr" in o_r[4]:
            assert o_b[4]["osnr"] == pytest.approx(o_r[4]["osnr"], abs=1e-3)
        else:
            assert {k: o_b[4][k] for k in ("blocked_due_to_resources", "blocked_due_to_osnr", "rejected")} == \
                   {k: o_r[4][k] for k in ("blocked_due_to_resources", "blocked_due_to_osnr", "rejected")}
        n_acc += a_r != ref.action_space.n - 1
    assert n_acc > steps // 2
    if name == "heuristic_exact_fit":   # (n_acc counts calls: the stalled request is offered again and again)
        assert env.current_service.service_id == ref.current_service.service_id < steps
    assert np.array_equal(env.available_slots_matrix(), np.asarray(ref.topology.graph["available_slots"]))
    env.close()


def test_mask_heuristic_and_calculate_osnr_triple():
    """The mask-based policy of heuristics.py:419 on both envs (gen_observation=True: observation and GSNR-validated mask
    from qrmsa_observation), and calculate_osnr's (GSNR, ASE, NLI) for the candidate each env just provisioned."""
    from optical_networking_gym_b200.compat import patch_reference_heuristics

    _, _, H, ref_osnr = rh.import_reference()
    patch_reference_heuristics(H)
    ref, env, _ = _pair(99, 12, load=210.0, gen_observation=True)
    obs_r, info_r = ref.reset()
    obs_b, info_b = env.reset()
    for t in range(8):
        _same_request(env.current_service, ref.current_service)
        assert np.array_equal(info_b["mask"], info_r["mask"]), f"mask differs at step {t}"
        assert np.abs(obs_b - obs_r).max() < 2e-6
        a_r = H.shortest_available_path_first_fit_best_modulation(info_r["mask"])
        a_b = H.shortest_available_path_first_fit_best_modulation(info_b["mask"])
        assert a_r == a_b
        svc_r, svc_b = ref.current_service, env.current_service
        if a_r != ref.action_space.n - 1:
            # the candidate written on the current service the way the heuristics do it (heuristics.py:943-955)
            for e_, svc in ((ref, svc_r), (env, svc_b)):
                route, mod, slot = e_.encoded_decimal_to_array(a_r)
                svc.path = e_.k_shortest_paths[svc.source, svc.destination][route]
                svc.current_modulation = e_.modulations[mod]
                svc.number_slots = e_.get_number_slots(svc, svc.current_modulation)
                svc.initial_slot = slot
                svc.center_frequency = e_.frequency_start + e_.frequency_slot_bandwidth * slot + e_.frequency_slot_bandwidth * (svc.number_slots / 2)
                svc.bandwidth = e_.frequency_slot_bandwidth * svc.number_slots
                svc.launch_power = e_.launch_power
            g = np.array(H.calculate_osnr(env, svc_b))
            g_ref = np.array(ref_osnr.calculate_osnr(ref, svc_r))
            assert g.shape == g_ref.shape == (3,)
            assert np.abs(g - g_ref).max() < 1e-6, (g, g_ref)
        obs_r, _, _, _, info_r = ref.step(a_r)
        obs_b, _, _, _, info_b = env.step(a_b)
        if svc_r.accepted:
            assert svc_b.accepted
            assert svc_b.OSNR == pytest.approx(svc_r.OSNR, abs=1e-6)
            assert svc_b.ASE == pytest.approx(svc_r.ASE, abs=1e-6) and svc_b.NLI == pytest.approx(svc_r.NLI, abs=1e-6)
    env.close()


def test_continuous_bit_rates_vs_reference():
    """bit_rate_selection="continuous": rng.randint(lower, higher) per request (qrmsa.pyx:246-254, :1086-1087)."""
    from optical_networking_gym_b200.heuristics import heuristic_shortest_available_path_first_fit_best_modulation as h_b200

    ref, env, _ = _pair(777, 121, bit_rate_selection="continuous", bit_rate_lower_bound=25.0, bit_rate_higher_bound=100.0)
    h_ref = rh.first_fit_heuristic()
    rates = set()
    for t in range(120):
        _same_request(env.current_service, ref.current_service)
        rates.add(env.current_service.bit_rate)
        a_b, a_r = h_b200(env)[0], h_ref(ref)[0]
        assert a_b == a_r, t
        o_b, o_r = env.step(a_b), ref.step(a_r)
        assert o_b[1] == o_r[1] and o_b[2] == o_r[2]
    assert len(rates) > 30 and min(rates) >= 25 and max(rates) <= 100
    assert np.array_equal(env.available_slots_matrix(), np.asarray(ref.topology.graph["available_slots"]))
    env.close()


def test_only_episode_counters_reset_vs_reference():
    """reset(options={"only_episode_counters": True}) (qrmsa.pyx:427-464, used by examples/ONDM_2025/ppo_debugger.py:164):
    the episode counters restart on the live network and -- the release heap being emptied -- the services running at
    that moment are never released."""
    from optical_networking_gym_b200.heuristics import heuristic_shortest_available_path_first_fit_best_modulation as h_b200

    L = 60
    ref, env, _ = _pair(31, L, load=400.0)
    h_ref = rh.first_fit_heuristic()
    for ep in range(3):
        # the first episode ends after L - 1 steps (reset drew request 0); a continued one counts from zero without
        # drawing a request, so it ends after L
        for t in range(L - 1 if ep == 0 else L):
            _same_request(env.current_service, ref.current_service)
            a_b, a_r = h_b200(env)[0], h_ref(ref)[0]
            assert a_b == a_r, (ep, t)
            o_b, o_r = env.step(a_b), ref.step(a_r)
            assert o_b[1] == o_r[1] and o_b[2] == o_r[2]
            for k in ("episode_services_accepted", "episode_service_blocking_rate", "service_blocking_rate"):
                assert o_b[4][k] == pytest.approx(o_r[4][k], abs=1e-12), (ep, t, k)
        assert o_b[2] and o_r[2]
        assert np.array_equal(env.available_slots_matrix(), np.asarray(ref.topology.graph["available_slots"]))
        ob, ib = env.reset(options={"only_episode_counters": True})
        orr, ir = ref.reset(options={"only_episode_counters": True})
        assert ib == ir == {} and ob.shape == orr.shape
        assert env.episode_services_processed == 0
    # the network kept every service of the earlier episodes: it is fuller than one episode alone could make it
    assert int((env.available_slots_matrix() == 0).sum()) == int((np.asarray(ref.topology.graph["available_slots"]) == 0).sum())
    env.close()


def test_file_name_writes_the_reference_csv(tmp_path):
    """QRMSAEnv(file_name=...) (qrmsa.pyx:387-406, :967-990): same file name rule, header and lines."""
    from optical_networking_gym_b200.heuristics import heuristic_shortest_available_path_first_fit_best_modulation as h_b200

    # two directories so that the two implementations do not write the same file
    from optical_networking_gym_b200.env import QRMSAEnv
    _, ref_qrmsa, _, _ = rh.import_reference()
    topo = rh.make_topology("nsfnet")
    kw = rh.env_kwargs(topo, n_slots=320, load=300.0, episode_length=81)
    ctor_seed = kw["seed"]
    with rh.seeded_random(5):
        ref = ref_qrmsa.QRMSAEnv(**dict(kw, file_name=str(tmp_path / "ref" / "svc")))
    kw2 = dict(kw); kw2.pop("seed")
    env = QRMSAEnv(seed=5, **dict(kw2, file_name=str(tmp_path / "b200" / "svc")))
    h_ref = rh.first_fit_heuristic()
    for t in range(80):
        a_b, a_r = h_b200(env)[0], h_ref(ref)[0]
        assert a_b == a_r
        env.step(a_b); ref.step(a_r)
    env.close()       # (the reference flushes every line and keeps its file open)
    # the file name carries the constructor's seed argument: the B200 env was given the stream seed, the reference its own
    ref_file = str(tmp_path / "ref" / f"svc_{topo.graph['name']}_1.0_300.0_{ctor_seed}.csv")
    assert os.path.exists(ref_file)
    assert env.final_file_name == str(tmp_path / "b200" / f"svc_{topo.graph['name']}_1.0_300.0_5.csv")
    got, want = open(env.final_file_name).read().splitlines(), open(ref_file).read().splitlines()
    assert got[:2] == want[:2] and len(got) == len(want) == 82
    for a, b in zip(got[2:], want[2:]):
        fa, fb = a.split(","), b.split(",")
        assert fa[:8] == fb[:8] and fa[11:] == fb[11:], (a, b)       # ids, endpoints, rate, path, modulation, counts: exact text
        for i in (8, 9, 10):
            assert float(fa[i]) == pytest.approx(float(fb[i]), abs=1e-6)


def test_two_modulations_to_consider_env_vs_live_reference():
    """QRMSAEnv(modulations_to_consider=2, gen_observation=True) next to the reference env (the ONDM PPO setting,
    examples/ONDM_2025/new_train_multi_ppo.py:101): action space, masks, max_modulation_idx and the decode of the mask-chosen
    action follow the reference request by request."""
    _, _, H, _ = rh.import_reference()
    ref, env, _ = _pair(2025, 10, load=260.0, gen_observation=True, modulations_to_consider=2)
    assert env.action_space.n == ref.action_space.n == 5 * 2 * 320 + 1
    obs_r, info_r = ref.reset()
    obs_b, info_b = env.reset()
    for t in range(7):
        _same_request(env.current_service, ref.current_service)
        assert env.max_modulation_idx == ref.max_modulation_idx, t
        assert np.array_equal(info_b["mask"], info_r["mask"]), f"mask differs at step {t}"
        assert np.abs(obs_b - obs_r).max() < 2e-6
        a = H.shortest_available_path_first_fit_best_modulation(info_r["mask"])
        assert env.encoded_decimal_to_array(a) == list(ref.encoded_decimal_to_array(a))
        obs_r, rw_r, term_r, _, info_r = ref.step(a)
        obs_b, rw_b, term_b, _, info_b = env.step(a)
        assert rw_b == rw_r and term_b == term_r
        assert info_b["chosen_slot"] == info_r["chosen_slot"] and info_b["osnr"] == pytest.approx(info_r["osnr"], abs=1e-3)
    assert np.array_equal(env.available_slots_matrix(), np.asarray(ref.topology.graph["available_slots"]))
    env.close()
