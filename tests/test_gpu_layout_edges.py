"""Layout boundaries of the device state against the oracle: slot counts on both sides of the 64-/128-byte row switch
(479 / 480), the smallest and the largest supported spectrum, env counts that are not multiples of the warp or CTA
size, a one-request-ahead trace (every launch decides one request), and the error paths that must fail loudly."""
import numpy as np
import pytest

from helpers import load_tables

pytestmark = pytest.mark.gpu


def _check(tb, n_envs, n, load, seed, chunks):
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine, unpack_bitmaps
    from optical_networking_gym_b200.tracegen import TraceGenerator
    from oracle import oracle as orc

    tr = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, load, base_seed=seed).next(n + 1)
    eng = Engine(tb, n_envs, n + 1)
    eng.reset(); eng.load_trace_host(*tr)
    done = 0
    for c in chunks:
        eng.step_first_fit(c); done += c
    assert done == n
    words = eng.actions_host(0, n)
    actions = (words & _lib.ACTION_MASK).astype(np.int64).T
    flagged = ((words.view(np.uint32) & _lib.FLAG_NEAR_THRESHOLD) != 0).T
    slots = unpack_bitmaps(eng.export_bitmaps(0, n_envs), tb.n_slots)
    n_rej = 0
    for e in range(n_envs):
        o = orc.OracleEnv(tb, n + 1)
        o.reset(*[a[:, e] for a in tr])
        ref = o.run_first_fit(n, log_qot=False)
        if not np.array_equal(ref["action"], actions[e]):
            d = int(np.flatnonzero(ref["action"] != actions[e])[0])
            assert flagged[e, d], f"S={tb.n_slots} env {e}: unflagged mismatch at step {d}"
            continue
        assert np.array_equal(o.slots(), slots[e]), f"S={tb.n_slots} env {e}: bitmap mismatch"
        for l in range(tb.n_links):
            assert sorted(map(tuple, eng.export_link_list(e, l))) == sorted(map(tuple, o.link_list(l)))
        n_rej += int((ref["accepted"] == 0).sum())
    c = eng.counters_dict()
    assert c["decided"] == n_envs * n and c["errors"] == 0
    eng.close()
    return n_rej


@pytest.mark.parametrize("S,load", [(64, 60.0), (479, 420.0), (480, 420.0), (800, 800.0)])
def test_slot_count_boundaries(S, load):
    """479 slots: 64-byte rows, one-byte positions; 480: 128-byte rows, two-byte positions; 64 / 800: the extremes (at 800 slots the GN tables fill the 227 KB of shared memory)."""
    tb = load_tables("nsfnet", 320).replace(n_slots=S)
    rej = _check(tb, 5, 260, load, 1000 + S, (1, 130, 129))
    assert rej > 0            # the spectrum end (guard-slot rule at the last slot) and rejections were exercised


@pytest.mark.parametrize("n_envs", [1, 31, 33, 150])
def test_ragged_env_counts(n_envs):
    tb = load_tables("ring4", 320)
    _check(tb, n_envs, 120, 60.0, 7, (120,))


def test_one_request_per_launch_and_idle_launches():
    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.tracegen import TraceGenerator

    tb = load_tables("ring4", 320)
    n = 40
    _check(tb, 3, n, 80.0, 11, (1,) * n)
    # stepping past the end of the loaded trace decides nothing and raises nothing
    tr = TraceGenerator(2, tb.n_nodes, tb.n_rates, 50.0, base_seed=3).next(6)
    eng = Engine(tb, 2, 6)
    eng.reset(); eng.load_trace_host(*tr)
    eng.step_first_fit(50)
    assert eng.counters_dict()["decided"] == 2 * 5
    eng.step_first_fit(50)
    assert eng.counters_dict()["decided"] == 2 * 5 and (eng.env_state()[:, 0] == 5).all()
    eng.close()


def test_loud_failures():
    from optical_networking_gym_b200._lib import QRMSAError
    from optical_networking_gym_b200.engine import Engine

    tb = load_tables("ring4", 320)
    with pytest.raises(QRMSAError):
        Engine(tb.replace(n_slots=961), 1, 10)            # beyond one bitmap word per lane + the virtual slot
    with pytest.raises(QRMSAError):
        Engine(tb, 1, 20000)                               # episode longer than the in-shared-memory schedule sort
    eng = Engine(tb, 2, 10)
    with pytest.raises(QRMSAError):
        eng.step_first_fit(1)                              # no trace loaded
    with pytest.raises(QRMSAError):
        eng.step_heuristic(7, 1)                           # unknown policy
    eng.close()


@pytest.mark.parametrize("topo,S,policy,load", [("nobel-eu", 320, "first_fit", 300.0), ("nsfnet", 320, "load_balancing", 250.0),
                                               ("nobel-eu", 320, "load_balancing_first_fit", 300.0),
                                               ("germany50", 640, "first_fit", 800.0)])
def test_shared_memory_staging_matches_global_state(topo, S, policy, load):
    """The step kernel keeps the link rows, the request / schedule stream chunks and a compact path table in shared
    memory when they fit (qrmsa_create; level 2), only the stream chunks when the rows do not fit (germany50/640: the
    tables alone take 183 KB; level 1), and qrmsa_set_staging(0) reads everything through L1.  Same action words,
    bitmaps, channel lists, env state and counters at every level, across launch boundaries (rows are written back
    after every launch)."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.tracegen import TraceGenerator

    tb = load_tables(topo, S)
    n_envs, n = 70, 400
    tr = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, load, base_seed=99).next(n + 1)
    out = {}
    for level in (2, 1, 0):
        eng = Engine(tb, n_envs, n + 1)
        eng.set_staging(level)
        eng.reset(); eng.load_trace_host(*tr)
        for c in (1, 7, 200, n - 208):
            eng.step_heuristic(_lib.POLICIES[policy], c)
        lists = [sorted(map(tuple, eng.export_link_list(e, l))) for e in (0, n_envs - 1) for l in range(tb.n_links)]
        out[level] = (eng.actions_host(0, n).copy(), eng.export_bitmaps(0, n_envs).copy(), eng.env_state().copy(),
                      eng.counters().copy(), lists)
        c = eng.counters_dict()
        assert c["decided"] == n_envs * n and c["errors"] == 0
        eng.close()
    for level in (1, 0):
        for i, what in enumerate(("action words", "bitmaps", "env state", "counters")):
            assert np.array_equal(out[2][i], out[level][i]), f"{what} differ between staging levels 2 and {level}"
        assert out[2][4] == out[level][4], "channel lists differ"
