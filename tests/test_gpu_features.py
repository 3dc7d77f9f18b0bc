"""measure_disruptions and defragmentation on the device (qrmsa_set_features; reference envs/qrmsa.pyx:937-952,
:1113-1122, :1545-1639) against recordings of the compiled reference, against the oracle on fresh traces, and -- where
oracle/_ref is present -- side by side with the reference env through the QRMSAEnv API."""
import numpy as np
import pytest

from helpers import TRACE_KEYS, load_golden, load_tables
from oracle import oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.gpu

FEATS = ["feat_disrupt_nsfnet_320_l600_s7", "feat_defrag_nsfnet_320_l300_s9_n5", "feat_defrag_nsfnet_320_l150_s11_n0",
         "feat_both_nobel-eu_320_l400_s13_n3"]


def _engine(tag):
    from optical_networking_gym_b200.engine import Engine

    g = load_golden(tag)
    tb = load_tables(tag.split("_")[2], 320)
    n = len(g["action"])
    eng = Engine(tb, 1, n + 1)
    md, df, nd = (int(x) for x in g["feat"])
    eng.set_features(md, df, nd)
    eng.reset()
    eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    return eng, g, tb, n, md


@pytest.mark.parametrize("tag", FEATS)
def test_fused_first_fit_with_features_vs_reference(tag):
    """One request per launch, so that every step's disrupted count and the counters can be compared."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import unpack_bitmaps

    eng, g, tb, n, md = _engine(tag)
    for t in range(n):
        c = eng.counters_dict()
        # info is assembled before the next request is drawn (qrmsa.pyx:996-1052)
        assert c["service_reallocations"] == int(g["realocations"][t]) and c["defrag_cycles"] == int(g["defrag_cicles"][t]), t
        eng.step_first_fit(1)
        if md:
            assert int(eng.step_disrupted()[0]) == int(g["disrupted_local"][t]), f"step {t}"
    words = eng.actions_host(0, n).view(np.uint32)[:, 0]
    accepted = (words & _lib.FLAG_ACCEPTED) != 0
    assert np.array_equal(accepted, g["accepted"].astype(bool))
    # defragmentation rewrites the start slot of a moved service in its action word: path and modulation digits stay
    S = tb.n_slots
    assert np.array_equal((words & _lib.ACTION_MASK)[accepted] // S, g["action"][accepted] // S)
    if not int(g["feat"][1]):
        assert np.array_equal((words & _lib.ACTION_MASK).astype(np.int64), g["action"])
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), S)[0], g["final_slots"])
    c = eng.counters_dict()
    assert c["disrupted_services"] == int(g["disrupted_local"].sum()) and c["errors"] == 0
    assert int(((words & _lib.FLAG_DISRUPTED) != 0).sum()) == int(g["disrupted_local"].sum())
    eng.close()


@pytest.mark.parametrize("tag", FEATS)
def test_features_multi_step_launch_and_step_action(tag):
    """The same run in three launches of the fused kernel, and once more through qrmsa_step_action with the recorded
    actions: same final state and counters."""
    import torch
    from optical_networking_gym_b200.engine import unpack_bitmaps

    eng, g, tb, n, md = _engine(tag)
    for c in (7, n // 2, n - 7 - n // 2):
        eng.step_first_fit(c)
    S = tb.n_slots
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), S)[0], g["final_slots"])
    c1 = eng.counters_dict()
    assert c1["disrupted_services"] == int(g["disrupted_local"].sum()) and c1["errors"] == 0
    # step_action: the actions the reference took, one by one (a moved service keeps its recorded start: the action was
    # valid when it was taken)
    eng.reset()
    eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    a = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.zeros(1, dtype=torch.uint8, device="cuda")
    for t in range(n):
        a[0] = int(g["action"][t])
        eng.step_action(a, None, st, None, None)
        assert int(st[0]) == (0 if g["accepted"][t] else 1), f"step {t}"
        if md:
            assert int(eng.step_disrupted()[0]) == int(g["disrupted_local"][t]), f"step {t}"
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), S)[0], g["final_slots"])
    c2 = eng.counters_dict()
    for k in ("disrupted_services", "defrag_cycles", "service_reallocations", "accepted", "releases"):
        assert c1[k] == c2[k], k
    assert c2["service_reallocations"] >= int(g["realocations"][-1]) and c2["defrag_cycles"] >= int(g["defrag_cicles"][-1])
    eng.close()


def test_features_batched_vs_oracle():
    """Many envs per launch with both switches on, against the oracle on CPython-exact traces: decisions (path,
    modulation, acceptance), final bitmaps, disrupted / reallocation / cycle counters."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine, unpack_bitmaps
    from optical_networking_gym_b200.tracegen import TraceGenerator

    tb = load_tables("nobel-eu", 320)
    n_envs, n = 37, 220
    tr = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, 450.0, base_seed=4321).next(n + 1)
    eng = Engine(tb, n_envs, n + 1)
    eng.set_features(True, True, 4)
    eng.reset(); eng.load_trace_host(*tr)
    for c in (1, 100, n - 101):
        eng.step_first_fit(c)
    words = eng.actions_host(0, n).view(np.uint32)
    slots = unpack_bitmaps(eng.export_bitmaps(0, n_envs), 320)
    tot = dict(dis=0, mv=0, cyc=0)
    n_flagged = 0
    for e in range(n_envs):
        o = orc.OracleEnv(tb, n + 1)
        o.set_features(True, True, 4)
        o.reset(*[a[:, e] for a in tr])
        ref = o.run_first_fit(n, log_qot=False)
        w = words[:, e]
        if (w & _lib.FLAG_NEAR_THRESHOLD).any():      # a near-threshold check anywhere may legitimately change what follows
            n_flagged += 1
            if not np.array_equal(ref["accepted"].astype(bool), (w & _lib.FLAG_ACCEPTED) != 0):
                continue
        assert np.array_equal(ref["accepted"].astype(bool), (w & _lib.FLAG_ACCEPTED) != 0), e
        assert np.array_equal(o.slots(), slots[e]), f"env {e}: bitmaps differ"
        starts = np.array([o.service_start(i) for i in range(n)])
        acc = ref["accepted"].astype(bool)
        assert np.array_equal((w & _lib.ACTION_MASK)[acc] % 320, starts[acc]), f"env {e}: current start slots differ"
        fc = o.feature_counters()
        tot["dis"] += fc["disrupted_services"]; tot["mv"] += fc["episode_service_realocations"]; tot["cyc"] += fc["episode_defrag_cicles"]
    c = eng.counters_dict()
    assert c["errors"] == 0
    if n_flagged == 0:
        assert (c["disrupted_services"], c["service_reallocations"], c["defrag_cycles"]) == (tot["dis"], tot["mv"], tot["cyc"])
    assert tot["mv"] > 0 and tot["cyc"] > 0
    eng.close()


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_env_api_with_features_vs_live_reference():
    """QRMSAEnv(measure_disruptions=True, defragmentation=True, n_defrag_services=3) next to the reference env: actions,
    rewards and the disruption / defragmentation info keys (including the truncated episode ratio, qrmsa.pyx:1038-1041)."""
    from optical_networking_gym_b200.env import QRMSAEnv
    from optical_networking_gym_b200.heuristics import heuristic_shortest_available_path_first_fit_best_modulation as h_b200

    _, ref_qrmsa, _, _ = rh.import_reference()
    topo = rh.make_topology("nsfnet")
    kw = rh.env_kwargs(topo, n_slots=320, load=600.0, episode_length=161)
    kw.update(measure_disruptions=True, defragmentation=True, n_defrag_services=3)
    with rh.seeded_random(17):
        ref = ref_qrmsa.QRMSAEnv(**kw)
    kw2 = dict(kw); kw2.pop("seed")
    env = QRMSAEnv(seed=17, **kw2)
    h_ref = rh.first_fit_heuristic()
    seen_disrupted = False
    for t in range(160):
        a_b, a_r = h_b200(env)[0], h_ref(ref)[0]
        assert a_b == a_r, t
        o_b, o_r = env.step(a_b), ref.step(a_r)
        assert o_b[1] == o_r[1] and o_b[2] == o_r[2]
        for k in ("disrupted_services", "episode_disrupted_services", "episode_defrag_cicles", "episode_service_realocations",
                  "episode_services_accepted"):
            assert o_b[4][k] == pytest.approx(o_r[4][k], abs=1e-12), (t, k)
        seen_disrupted |= o_r[4]["disrupted_services"] > 0
    assert seen_disrupted and o_r[4]["episode_service_realocations"] > 0
    assert np.array_equal(env.available_slots_matrix(), np.asarray(ref.topology.graph["available_slots"]))
    env.close()
