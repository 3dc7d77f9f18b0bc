"""The oracle (oracle/qrmsa_oracle.c) against every golden vector recorded from the compiled reference.
CPU only.  Bar: bit-exact actions / accept flags / slot matrices, GSNR to 1e-9 dB."""
import numpy as np
import pytest

from helpers import TRACE_KEYS, load_golden, load_tables, parse_tag
from oracle import oracle as orc

SINGLE = ["run_nobel-eu_320_l300_s50", "run_nsfnet_320_l300_s50", "run_germany50_640_l800_s52",
          "run_nobel-eu_320_l500_s7", "run_ring4_320_l60_s3"]
MULTI = ["multi_nobel-eu_320_l300_b50", "multi_germany50_640_l800_b50"]


@pytest.mark.parametrize("tag", SINGLE)
def test_oracle_single(tag):
    topo, S = parse_tag(tag)
    tb, g = load_tables(topo, S), load_golden(tag)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*[g[k] for k in TRACE_KEYS])
    done = 0
    acts, gs = [], []
    for step, snap in zip(g["snap_steps"], g["snap_slots"]):
        r = o.run_first_fit(int(step) - done, log_qot=False)
        acts.append(r["action"]); gs.append(r["gsnr"])
        done = int(step)
        assert np.array_equal(o.slots(), np.unpackbits(snap, axis=1)[:, :S]), f"snapshot {step}"
    r = o.run_first_fit(n - done, log_qot=False)
    acts.append(r["action"]); gs.append(r["gsnr"])
    assert np.array_equal(np.concatenate(acts), g["action"])
    assert np.abs(np.concatenate(gs) - g["gsnr"]).max() < 1e-9
    assert np.array_equal(o.slots(), g["final_slots"])
    c = o.counters()
    assert c["accepted"] == int(g["accepted"].sum()) and c["ep_processed"] == n + 1


def test_oracle_qot_log_matches_reference():
    tag = "run_nobel-eu_320_l500_s7"
    topo, S = parse_tag(tag)
    tb, g = load_tables(topo, S), load_golden(tag)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*[g[k] for k in TRACE_KEYS])
    r = o.run_first_fit(n, log_qot=True)
    assert np.array_equal(r["qot_step"], g["qot_step"])
    assert np.array_equal(r["qot_thr"], g["qot_thr"])
    assert np.abs(r["qot_gsnr"] - g["qot_gsnr"]).max() < 1e-9


@pytest.mark.parametrize("tag", MULTI)
def test_oracle_multi(tag):
    topo, S = parse_tag(tag)
    tb, g = load_tables(topo, S), load_golden(tag)
    n_envs, n = g["action"].shape
    ref_slots = np.unpackbits(g["final_slots"], axis=2)[:, :, :S]
    for e in range(n_envs):
        o = orc.OracleEnv(tb, n + 1)
        o.reset(*[g[k][e] for k in TRACE_KEYS])
        r = o.run_first_fit(n, log_qot=False)
        assert np.array_equal(r["action"], g["action"][e]), f"env {e}"
        assert np.array_equal(o.slots(), ref_slots[e])
        assert np.abs(r["gsnr"] - g["gsnr"][e]).max() < 1e-9


def test_oracle_step_action_vs_reference():
    """env.step() with external actions: accepted / reject / not-free (request not consumed) / low GSNR (raise)."""
    g = load_golden("rl_nsfnet_320_l210_s11")
    tb = load_tables("nsfnet", 320)
    n_req = len(g["src"])
    o = orc.OracleEnv(tb, n_req)
    o.reset(*[g[k] for k in TRACE_KEYS])
    for i, (a, st, rw, gs) in enumerate(zip(g["action"], g["status"], g["reward"], g["gsnr"])):
        status, reward, gsnr, term = o.step_action(int(a), n_req)
        assert status == st, f"call {i}: status {status} != {st}"
        if st != 3:
            assert reward == pytest.approx(rw, abs=1e-12)
        if st == 0:
            assert gsnr == pytest.approx(gs, abs=1e-9)
        assert term == bool(g["term"][i])
    assert np.array_equal(o.slots(), g["final_slots"])


def test_oracle_edge_cases():
    tb = load_tables("ring4", 320)
    # a single request: no step possible (a step needs the following request)
    o = orc.OracleEnv(tb, 1)
    o.reset(np.array([0], np.uint8), np.array([1], np.uint8), np.array([0], np.uint8),
            np.array([1.0], np.float32), np.array([5.0], np.float32))
    with pytest.raises(RuntimeError):
        o.run_first_fit(1)
    # zero holding time: released at the next arrival, network returns to all-free
    n = 50
    src = np.zeros(n, np.uint8); dst = np.full(n, 2, np.uint8); rate = np.full(n, 4, np.uint8)
    arr = np.arange(1, n + 1, dtype=np.float32); hold = np.zeros(n, np.float32)
    o = orc.OracleEnv(tb, n)
    o.reset(src, dst, rate, arr, hold)
    r = o.run_first_fit(n - 1)
    assert r["accepted"].all() and (r["action"] == r["action"][0]).all()
    assert o.slots().all()


@pytest.mark.parametrize("tag,topo", [("obs_nsfnet_320_l210_s21", "nsfnet"), ("obs_nobel-eu_320_l400_s8", "nobel-eu")])
def test_oracle_observation_and_mask_vs_reference(tag, topo):
    """gen_observation=True: observation vector and GSNR-validated action mask, every step of the recording."""
    import os
    from helpers import GOLDEN

    if not os.path.exists(os.path.join(GOLDEN, tag + ".npz")):
        pytest.skip("fixture not generated")
    g = load_golden(tag)
    tb = load_tables(topo, 320)
    n_req, n_act = len(g["src"]), int(g["n_actions"])
    mask_ref = np.unpackbits(g["mask"], axis=1)[:, :n_act]
    o = orc.OracleEnv(tb, n_req)
    o.reset(*[g[k] for k in TRACE_KEYS])
    for t in range(len(g["action"]) + 1):
        obs, mask = o.observation()
        assert np.abs(obs - g["obs"][t]).max() <= 1e-7, f"obs at step {t}"
        assert np.array_equal(mask, mask_ref[t]), f"mask at step {t}"
        if t < len(g["action"]):
            st, rw, _, _ = o.step_action(int(g["action"][t]), n_req)
            assert st in (0, 1) and rw == pytest.approx(float(g["reward"][t]), abs=1e-12)
    assert np.array_equal(o.slots(), g["final_slots"])


@pytest.mark.parametrize("tag,topo", [("policy_lb_nobel-eu_320_l400_s9", "nobel-eu"), ("policy_lb_nsfnet_320_l300_s4", "nsfnet")])
def test_oracle_load_balancing_vs_reference(tag, topo):
    """load_balancing_best_modulation (heuristics.py:547-627), the reference benchmark's heuristic #4."""
    tb, g = load_tables(topo, 320), load_golden(tag)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*[g[k] for k in TRACE_KEYS])
    r = o.run_first_fit(n, policy=1)
    assert np.array_equal(r["action"], g["action"])
    assert np.abs(r["gsnr"] - g["gsnr"]).max() < 1e-9
    assert np.array_equal(r["qot_step"], g["qot_step"]) and np.abs(r["qot_gsnr"] - g["qot_gsnr"]).max() < 1e-9
    assert np.array_equal(o.slots(), g["final_slots"])


@pytest.mark.parametrize("tag,topo", [("policy_hsnr_nsfnet_320_l300_s21", "nsfnet"), ("policy_hsnr_nobel-eu_320_l400_s5", "nobel-eu")])
def test_oracle_highest_snr_vs_reference(tag, topo):
    """heuristic_highest_snr (heuristics.py:272-328), the reference benchmark's heuristic #2: every valid start of every
    (path, modulation) is QoT-checked; the recording keeps the number of checks per step instead of the full log."""
    import os
    from helpers import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, tag + ".npz")):
        pytest.skip("recording not generated")
    tb, g = load_tables(topo, 320), load_golden(tag)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*[g[k] for k in TRACE_KEYS])
    r = o.run_first_fit(n, policy=2)
    assert np.array_equal(r["action"], g["action"])
    assert np.abs(r["gsnr"] - g["gsnr"]).max() < 1e-9
    assert np.array_equal(np.bincount(r["qot_step"], minlength=n), g["n_checks"])
    assert np.array_equal(o.slots(), g["final_slots"])


@pytest.mark.parametrize("tag,topo", [("policy_lbff_nobel-eu_320_l400_s13", "nobel-eu"), ("policy_lbff_nsfnet_320_l300_s8", "nsfnet")])
def test_oracle_lb_first_fit_vs_reference(tag, topo):
    """heuristic_load_balancing_first_fit (heuristics.py:202-270): paths ordered by occupied fraction, then first fit."""
    tb, g = load_tables(topo, 320), load_golden(tag)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*[g[k] for k in TRACE_KEYS])
    r = o.run_first_fit(n, policy=3)
    assert np.array_equal(r["action"], g["action"])
    assert np.abs(r["gsnr"] - g["gsnr"]).max() < 1e-9
    assert np.array_equal(r["qot_step"], g["qot_step"]) and np.abs(r["qot_gsnr"] - g["qot_gsnr"]).max() < 1e-9
    assert np.array_equal(o.slots(), g["final_slots"])


VARIANTS = ["var_ondm_nsfnet", "var_margin_nobel-eu", "var_k3_nsfnet_160"]


@pytest.mark.parametrize("tag", VARIANTS)
def test_oracle_vs_reference_configuration_variants(tag):
    """Away from the JOCN configuration: the ONDM modulation set (two modulations share a threshold), a 1.5 dB margin
    with another launch power and a skewed bit-rate mix, k = 3 paths on 160 slots with other span parameters."""
    import os
    from helpers import GOLDEN
    from optical_networking_gym_b200.tables import StaticTables

    tb = StaticTables.load(os.path.join(GOLDEN, f"tables_{tag}.npz"))
    g = load_golden("run_" + tag)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    o.reset(*[g[k] for k in TRACE_KEYS])
    r = o.run_first_fit(n)
    assert np.array_equal(r["action"], g["action"])
    assert np.array_equal(r["accepted"], g["accepted"])
    assert np.abs(r["gsnr"] - g["gsnr"]).max() < 1e-9
    assert np.array_equal(r["qot_step"], g["qot_step"]) and np.abs(r["qot_gsnr"] - g["qot_gsnr"]).max() < 1e-9
    assert np.array_equal(o.slots(), g["final_slots"])


FEATS = ["feat_disrupt_nsfnet_320_l600_s7", "feat_defrag_nsfnet_320_l300_s9_n5", "feat_defrag_nsfnet_320_l150_s11_n0",
         "feat_both_nobel-eu_320_l400_s13_n3"]


@pytest.mark.parametrize("tag", FEATS)
def test_oracle_disruptions_and_defragmentation_vs_reference(tag):
    """measure_disruptions (qrmsa.pyx:937-952) and defragmentation (:1113-1122, :1545-1639) restated in the oracle:
    decisions, per-step disrupted counts (the CSV column), reallocation / cycle counters and final slots against
    recordings of the compiled reference."""
    g = load_golden(tag)
    tb = load_tables(tag.split("_")[2], 320)
    n = len(g["action"])
    o = orc.OracleEnv(tb, n + 1)
    md, df, nd = (int(x) for x in g["feat"])
    o.set_features(md, df, nd)
    o.reset(*[g[k] for k in TRACE_KEYS])
    for t in range(n):
        # step()'s info is assembled before the next request is drawn (qrmsa.pyx:996-1052): the defragmentation counters
        # it reports are those of the previous step's release phase
        before = o.feature_counters()
        assert before["episode_service_realocations"] == int(g["realocations"][t]), f"step {t}"
        assert before["episode_defrag_cicles"] == int(g["defrag_cicles"][t]), f"step {t}"
        r = o.run_first_fit(1, log_qot=False)
        assert int(r["action"][0]) == int(g["action"][t]), f"step {t}"
        assert o.feature_counters()["last_step_disrupted"] == int(g["disrupted_local"][t]), f"step {t}"
    assert np.array_equal(o.slots(), g["final_slots"])
    assert o.feature_counters()["disrupted_services"] == int(g["disrupted_local"].sum())


def test_oracle_observation_with_fewer_modulations_vs_reference():
    """modulations_to_consider = 2 of 6 (examples/ONDM_2025/new_train_multi_ppo.py:101): observation() moves
    max_modulation_idx with every request (qrmsa.pyx:543-581, :680), the two blocks of a path stand for modulations
    max_idx and max_idx - 1 (:716-719) and step() decodes the action with it (:821-829)."""
    g = load_golden("obs_mc2_nsfnet_320_l260_s5")
    tb = load_tables("nsfnet", 320).replace(mods_to_consider=2)
    n_req, n_act = len(g["src"]), int(g["n_actions"])
    assert n_act == tb.n_actions == 5 * 2 * 320 + 1
    mask_ref = np.unpackbits(g["mask"], axis=1)[:, :n_act]
    o = orc.OracleEnv(tb, n_req)
    o.reset(*[g[k] for k in TRACE_KEYS])
    for t in range(len(g["action"]) + 1):
        obs, mask = o.observation()
        assert o.max_modulation_idx == int(g["max_mod"][t]), f"step {t}"
        assert np.abs(obs - g["obs"][t]).max() < 2e-6, f"obs at step {t}"
        assert np.array_equal(mask, mask_ref[t]), f"mask at step {t}"
        if t < len(g["action"]):
            st, rw, _, _ = o.step_action(int(g["action"][t]), n_req)
            assert st in (0, 1) and rw == pytest.approx(float(g["reward"][t]), abs=1e-9), f"step {t}"
    assert np.array_equal(o.slots(), g["final_slots"])
