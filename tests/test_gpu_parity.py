"""CUDA path vs the compiled reference's recorded outputs (tests/golden) and vs the oracle.

Everything goes through the C ABI (optical_networking_gym_b200.engine.Engine -> libqrmsa_b200.so).
Bar: bit-exact path / modulation / initial-slot / accept decisions and slot bitmaps; GSNR within
1e-3 dB; decisions with a GSNR within 1e-3 dB of a threshold are flagged, not counted.
"""
import numpy as np
import pytest

from helpers import TRACE_KEYS, check_flags_against_recording, compare_decisions, load_golden, load_tables, parse_tag

pytestmark = pytest.mark.gpu

GSNR_TOL_DB = 1e-3

SINGLE = ["run_nobel-eu_320_l300_s50", "run_nsfnet_320_l300_s50", "run_germany50_640_l800_s52",
          "run_nobel-eu_320_l500_s7", "run_ring4_320_l60_s3"]
MULTI = ["multi_nobel-eu_320_l300_b50", "multi_germany50_640_l800_b50"]


def _engine(tb, n_envs, n_req):
    from optical_networking_gym_b200.engine import Engine

    eng = Engine(tb, n_envs, n_req)
    eng.enable_gsnr_log(True)
    return eng


def _run(eng, g, multi, chunks):
    from optical_networking_gym_b200 import _lib

    tr = [np.ascontiguousarray(g[k].T if multi else g[k][:, None]) for k in TRACE_KEYS]
    n_req = tr[0].shape[0]
    eng.reset()
    eng.load_trace_host(*tr)
    done = 0
    for c in chunks:
        eng.step_first_fit(c)
        done += c
    assert done == n_req - 1
    words = eng.actions_host(0, n_req - 1)
    actions = (words & _lib.ACTION_MASK).astype(np.int64).T        # [env][step]
    flagged = ((words.view(np.uint32) & _lib.FLAG_NEAR_THRESHOLD) != 0).T
    accepted = ((words.view(np.uint32) & _lib.FLAG_ACCEPTED) != 0).T
    gsnr = eng.gsnr_host(0, n_req - 1).T
    return actions, flagged, accepted, gsnr


@pytest.mark.parametrize("tag", SINGLE)
def test_single_env_vs_reference(tag):
    from optical_networking_gym_b200.engine import unpack_bitmaps

    topo, S = parse_tag(tag)
    tb = load_tables(topo, S)
    g = load_golden(tag)
    n_steps = len(g["action"])
    eng = _engine(tb, 1, n_steps + 1)
    # uneven chunking exercises state hand-over between launches
    chunks = [1, 2, 37] + [n_steps - 40]
    actions, flagged, accepted, gsnr = _run(eng, g, False, chunks)
    n_cmp, n_exc = compare_decisions(actions, g["action"][None], flagged, tag)
    # the fixtures are fixed files: no env of them diverges, so nothing below is conditional
    assert n_exc == 0 and n_cmp == n_steps
    assert np.array_equal(accepted[0], g["accepted"].astype(bool))
    assert np.abs(gsnr[0] - g["gsnr"]).max() < GSNR_TOL_DB
    slots = unpack_bitmaps(eng.export_bitmaps(0, 1), S)[0]
    assert np.array_equal(slots, g["final_slots"])
    st = eng.env_state()[0]
    assert st[0] == n_steps and st[1] == int(g["accepted"].sum()) and st[3] == 0
    c = eng.counters_dict()
    assert c["decided"] == n_steps and c["accepted"] == int(g["accepted"].sum())
    # every QoT check of the reference is either evaluated or refused on the empty-network bound
    assert c["gn_evals"] + c["gn_pruned"] == len(g["qot_gsnr"])
    assert c["errors"] == 0
    # flags, both ways: a recorded check within 1e-3 dB of its threshold <=> our flag on that step
    assert check_flags_against_recording(flagged[0], g, GSNR_TOL_DB, tag) == c["near_threshold"]
    eng.close()


@pytest.mark.parametrize("tag", MULTI)
def test_batched_envs_vs_reference(tag):
    from optical_networking_gym_b200.engine import unpack_bitmaps

    topo, S = parse_tag(tag)
    tb = load_tables(topo, S)
    g = load_golden(tag)
    n_envs, n_steps = g["action"].shape
    eng = _engine(tb, n_envs, n_steps + 1)
    actions, flagged, accepted, gsnr = _run(eng, g, True, [n_steps])
    n_cmp, n_exc = compare_decisions(actions, g["action"], flagged, tag)
    assert n_exc == 0 and n_cmp == n_envs * n_steps          # fixed fixtures: every env agrees to the last step
    assert np.array_equal(accepted, g["accepted"].astype(bool))
    slots = unpack_bitmaps(eng.export_bitmaps(0, n_envs), S)
    ref_slots = np.unpackbits(g["final_slots"], axis=2)[:, :, :S]
    off = np.concatenate([[0], np.cumsum(g["qot_count"])])
    n_flags = 0
    for e in range(n_envs):
        assert np.array_equal(slots[e], ref_slots[e]), f"env {e}: slot bitmap differs"
        assert np.abs(gsnr[e] - g["gsnr"][e]).max() < GSNR_TOL_DB
        ge = {k: g[k][off[e]:off[e + 1]] for k in ("qot_step", "qot_gsnr", "qot_thr")}
        n_flags += check_flags_against_recording(flagged[e], ge, GSNR_TOL_DB, f"{tag} env {e}")
    c = eng.counters_dict()
    assert c["near_threshold"] == n_flags and c["errors"] == 0
    assert c["gn_evals"] + c["gn_pruned"] == len(g["qot_gsnr"])
    eng.close()


def test_snapshots_mid_episode():
    """Slot matrices after 500, 1000, ... steps equal the reference's snapshots."""
    from optical_networking_gym_b200.engine import unpack_bitmaps

    tag = "run_nobel-eu_320_l300_s50"
    topo, S = parse_tag(tag)
    tb, g = load_tables(topo, S), load_golden(tag)
    n_steps = len(g["action"])
    eng = _engine(tb, 1, n_steps + 1)
    tr = [np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS]
    eng.reset(); eng.load_trace_host(*tr)
    done = 0
    for step, snap in zip(g["snap_steps"], g["snap_slots"]):
        eng.step_first_fit(int(step) - done)
        done = int(step)
        slots = unpack_bitmaps(eng.export_bitmaps(0, 1), S)[0]
        assert np.array_equal(slots, np.unpackbits(snap, axis=1)[:, :S]), f"snapshot at step {step}"
    eng.close()


@pytest.mark.parametrize("tag", ["policy_lb_nobel-eu_320_l400_s9", "policy_lb_nsfnet_320_l300_s4"])
def test_load_balancing_policy_vs_reference(tag):
    """qrmsa_step_heuristic(QRMSA_POLICY_LOAD_BALANCING) against the reference's load_balancing_best_modulation."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import unpack_bitmaps

    topo = tag.split("_")[2]
    tb, g = load_tables(topo, 320), load_golden(tag)
    n = len(g["action"])
    eng = _engine(tb, 1, n + 1)
    eng.reset(); eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    for c in (3, 500, n - 503):
        eng.step_heuristic("load_balancing", c)
    words = eng.actions_host(0, n)
    actions = (words & _lib.ACTION_MASK).astype(np.int64).T
    flagged = ((words.view(np.uint32) & _lib.FLAG_NEAR_THRESHOLD) != 0).T
    n_cmp, n_exc = compare_decisions(actions, g["action"][None], flagged, tag)
    assert n_exc == 0 and n_cmp == n
    assert np.abs(eng.gsnr_host(0, n).T[0] - g["gsnr"]).max() < GSNR_TOL_DB
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], g["final_slots"])
    c = eng.counters_dict()
    assert c["gn_evals"] + c["gn_pruned"] == len(g["qot_gsnr"]) and c["errors"] == 0
    assert check_flags_against_recording(flagged[0], g, GSNR_TOL_DB, tag) == c["near_threshold"]
    eng.close()


def test_blocked_by_flags_and_log_counters():
    """Rejected requests carry the heuristic's (blocked_due_to_resources, blocked_due_to_osnr) pair in the action word
    (heuristics.py:966); the counters k_count_decisions derives from the log must equal a recount on the host."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine

    tag = "multi_nobel-eu_320_l300_b50"
    topo, S = parse_tag(tag)
    tb = load_tables(topo, S)
    g = load_golden(tag)
    tr = [np.ascontiguousarray(g[k].T) for k in TRACE_KEYS]
    n_req, n_envs = tr[0].shape
    eng = Engine(tb, n_envs, n_req)
    eng.reset()
    eng.load_trace_host(*tr)
    for c in (100, n_req - 1 - 100):
        eng.step_first_fit(c)
    w = eng.actions_host(0, n_req - 1).view(np.uint32)
    acc = (w & _lib.FLAG_ACCEPTED) != 0
    res = (w & _lib.FLAG_BLOCKED_RESOURCES) != 0
    osn = (w & _lib.FLAG_BLOCKED_OSNR) != 0
    assert ((w & _lib.FLAG_DECIDED) != 0).all()
    assert not (acc & (res | osn)).any()          # the pair is only reported with the reject action
    assert ((res | osn) | acc).all()              # a reject always has a cause
    c = eng.counters_dict()
    assert c["decided"] == w.size and c["accepted"] == int(acc.sum()) and c["rejected"] == int((~acc).sum())
    assert c["blocked_resources"] == int(res.sum()) and c["blocked_osnr"] == int(osn.sum())
    assert c["near_threshold"] == int(((w & _lib.FLAG_NEAR_THRESHOLD) != 0).sum())
    rates = np.asarray(tb.bit_rates)[tr[2][: n_req - 1]]
    assert c["rate_requested_milli"] == int(np.rint(rates * 1000).sum())
    assert c["rate_provisioned_milli"] == int(np.rint(rates[acc] * 1000).sum())
    m_idx = (tb.n_mods - 1) - ((w & _lib.ACTION_MASK) // tb.n_slots) % tb.n_mods
    assert np.array_equal(c["mod_hist"][: tb.n_mods], np.bincount(m_idx[acc], minlength=tb.n_mods))
    assert np.array_equal(eng.env_state()[:, 1], acc.sum(0))
    eng.close()


@pytest.mark.parametrize("tag", ["policy_hsnr_nsfnet_320_l300_s21", "policy_hsnr_nobel-eu_320_l400_s5"])
def test_highest_snr_policy_vs_reference(tag):
    """qrmsa_step_heuristic(QRMSA_POLICY_HIGHEST_SNR) against the reference's heuristic_highest_snr.  A decision is
    excused only after a step flagged NEAR_TIE (runner-up within 1e-6 dB of the winner: the reference compares GSNR in
    dB, where several 1/GSNR values collapse) or NEAR_THRESHOLD (a check within 1e-3 dB of its threshold that could
    change the outcome)."""
    import os
    from helpers import GOLDEN
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import unpack_bitmaps

    if not os.path.exists(os.path.join(GOLDEN, tag + ".npz")):
        pytest.skip("recording not generated")
    topo = tag.split("_")[2]
    tb, g = load_tables(topo, 320), load_golden(tag)
    n = len(g["action"])
    eng = _engine(tb, 1, n + 1)
    eng.reset(); eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    for c in (2, 31, n - 33):
        eng.step_heuristic("highest_snr", c)
    words = eng.actions_host(0, n)
    w = words.view(np.uint32)
    actions = (words & _lib.ACTION_MASK).astype(np.int64).T
    flagged = ((w & (_lib.FLAG_NEAR_THRESHOLD | _lib.FLAG_NEAR_TIE)) != 0).T
    n_cmp, n_exc = compare_decisions(actions, g["action"][None], flagged, tag)
    # the tie flag must not be raised for the same channel met under several modulations (exact, deterministic ties)
    assert int(((w & _lib.FLAG_NEAR_TIE) != 0).sum()) <= int((g["best_gap_db"] < 2e-6).sum())
    assert n_exc == 0 and n_cmp == n
    assert np.abs(eng.gsnr_host(0, n).T[0] - g["gsnr"]).max() < GSNR_TOL_DB
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], g["final_slots"])
    c = eng.counters_dict()
    assert c["gn_evals"] == int(g["n_checks"].sum())         # every QoT check of the reference, none skipped
    assert c["decided"] == n and c["accepted"] == int(g["accepted"].sum()) and c["errors"] == 0
    eng.close()


@pytest.mark.parametrize("topo,n_slots,load,n_envs,n", [("nobel-eu", 320, 450.0, 7, 60), ("germany50", 640, 800.0, 5, 40),
                                                       ("var_k3_nsfnet", 160, 150.0, 6, 120), ("nsfnet", 100, 70.0, 6, 120)])
def test_highest_snr_policy_batched_vs_oracle(topo, n_slots, load, n_envs, n):
    """Several envs per launch against the oracle on CPython-exact traces: the link-major kernel (spectra up to 320
    slots, k_step_highest_snr_links) and the general one (640 slots, k_step_highest_snr)."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine, unpack_bitmaps
    from optical_networking_gym_b200.tracegen import TraceGenerator
    from oracle import oracle as orc

    tb = load_tables(topo, n_slots) if (topo, n_slots) != ("nsfnet", 100) else load_tables("nsfnet", 320).replace(n_slots=100)
    tr = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, load, base_seed=77).next(n + 1)
    eng = Engine(tb, n_envs, n + 1)
    eng.reset(); eng.load_trace_host(*tr)
    eng.step_heuristic("highest_snr", n)
    words = eng.actions_host(0, n)
    actions = (words & _lib.ACTION_MASK).astype(np.int64).T
    flagged = ((words.view(np.uint32) & (_lib.FLAG_NEAR_THRESHOLD | _lib.FLAG_NEAR_TIE)) != 0).T
    slots = unpack_bitmaps(eng.export_bitmaps(0, n_envs), n_slots)
    ref = []
    for e in range(n_envs):
        o = orc.OracleEnv(tb, n + 1)
        o.reset(*[a[:, e] for a in tr])
        ref.append(o.run_first_fit(n, log_qot=False, policy=2)["action"])
        if np.array_equal(ref[-1], actions[e]):
            assert np.array_equal(o.slots(), slots[e])
    compare_decisions(actions, np.array(ref), flagged, "highest_snr batched")
    assert eng.counters_dict()["errors"] == 0
    eng.close()


@pytest.mark.parametrize("tag", ["policy_lbff_nobel-eu_320_l400_s13", "policy_lbff_nsfnet_320_l300_s8"])
def test_lb_first_fit_policy_vs_reference(tag):
    """qrmsa_step_heuristic(QRMSA_POLICY_LB_FIRST_FIT) against the reference's heuristic_load_balancing_first_fit."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import unpack_bitmaps

    topo = tag.split("_")[2]
    tb, g = load_tables(topo, 320), load_golden(tag)
    n = len(g["action"])
    eng = _engine(tb, 1, n + 1)
    eng.reset(); eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    for c in (5, 700, n - 705):
        eng.step_heuristic("load_balancing_first_fit", c)
    words = eng.actions_host(0, n)
    w = words.view(np.uint32)
    actions = (words & _lib.ACTION_MASK).astype(np.int64).T
    flagged = ((w & _lib.FLAG_NEAR_THRESHOLD) != 0).T
    n_cmp, n_exc = compare_decisions(actions, g["action"][None], flagged, tag)
    assert n_exc == 0 and n_cmp == n
    assert np.abs(eng.gsnr_host(0, n).T[0] - g["gsnr"]).max() < GSNR_TOL_DB
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], g["final_slots"])
    c = eng.counters_dict()
    assert c["gn_evals"] + c["gn_pruned"] == len(g["qot_gsnr"]) and c["errors"] == 0
    assert check_flags_against_recording(flagged[0], g, GSNR_TOL_DB, tag) == c["near_threshold"]
    rej = (w & _lib.FLAG_ACCEPTED) == 0                    # a reject always reports (True, False), heuristics.py:270
    assert ((w[rej] & _lib.FLAG_BLOCKED_RESOURCES) != 0).all() and ((w[rej] & _lib.FLAG_BLOCKED_OSNR) == 0).all()
    eng.close()


@pytest.mark.parametrize("tag", ["var_ondm_nsfnet", "var_margin_nobel-eu", "var_k3_nsfnet_160"])
def test_configuration_variants_vs_reference(tag):
    """The fused first-fit step away from the JOCN configuration (other modulation thresholds, margin, launch power,
    bit-rate mix, k, slot count, span parameters): decisions, GSNR, bitmaps and QoT-check count against the reference."""
    import os
    from helpers import GOLDEN
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import unpack_bitmaps
    from optical_networking_gym_b200.tables import StaticTables

    tb = StaticTables.load(os.path.join(GOLDEN, f"tables_{tag}.npz"))
    g = load_golden("run_" + tag)
    n = len(g["action"])
    eng = _engine(tb, 1, n + 1)
    actions, flagged, accepted, gsnr = _run(eng, g, False, [7, 900, n - 907])
    n_cmp, n_exc = compare_decisions(actions, g["action"][None], flagged, tag)
    assert n_exc == 0 and n_cmp == n
    assert np.array_equal(accepted[0], g["accepted"].astype(bool))
    assert np.abs(gsnr[0] - g["gsnr"]).max() < GSNR_TOL_DB
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), tb.n_slots)[0], g["final_slots"])
    c = eng.counters_dict()
    assert c["gn_evals"] + c["gn_pruned"] == len(g["qot_gsnr"]) and c["errors"] == 0
    assert check_flags_against_recording(flagged[0], g, GSNR_TOL_DB, tag) == c["near_threshold"]
    eng.close()
