"""Host request generator (csrc/tracegen.cpp) vs traces recorded from the compiled reference and vs CPython's
own `random` driven in the reference's draw order (oracle.generate_trace_python).  CPU only."""
import numpy as np
import pytest

from helpers import TRACE_KEYS, load_golden, load_tables, parse_tag
from oracle import oracle as orc
from optical_networking_gym_b200.tracegen import TraceGenerator


@pytest.mark.parametrize("tag", ["run_nobel-eu_320_l300_s50", "run_nsfnet_320_l300_s50", "run_germany50_640_l800_s52",
                                 "run_nobel-eu_320_l500_s7", "run_ring4_320_l60_s3"])
def test_single_stream_vs_reference(tag):
    topo, S = parse_tag(tag)
    tb, g = load_tables(topo, S), load_golden(tag)
    tg = TraceGenerator(1, tb.n_nodes, tb.n_rates, float(g["meta_load"]), base_seed=int(g["meta_seed"]))
    n = len(g["src"])
    a = tg.next(n // 3)
    b = tg.next(n - n // 3)          # the clock and the MT state persist across calls
    for i, k in enumerate(TRACE_KEYS):
        assert np.array_equal(np.concatenate([a[i], b[i]])[:, 0], g[k]), k


def test_batch_streams_vs_reference():
    g = load_golden("multi_nobel-eu_320_l300_b50")
    tb = load_tables("nobel-eu", 320)
    n_envs, n = g["src"].shape
    for threads in (1, 3):
        tg = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, 300.0, base_seed=int(g["meta_seed"]), n_threads=threads)
        out = tg.next(n)
        for i, k in enumerate(TRACE_KEYS):
            assert np.array_equal(out[i].T, g[k]), k


@pytest.mark.parametrize("seed,load,nodes,rates", [(0, 10.0, 4, 3), (1, 777.5, 14, 5), (2 ** 31 - 1, 50.0, 50, 2),
                                                    (2 ** 32 + 5, 300.0, 28, 5), (123456789012, 1.5, 7, 1)])
def test_vs_cpython_random(seed, load, nodes, rates):
    ref, _, _ = orc.generate_trace_python(nodes, rates, load, 10800.0, seed, 400)
    out = TraceGenerator(1, nodes, rates, load, base_seed=seed).next(400)
    for i, k in enumerate(TRACE_KEYS):
        assert np.array_equal(out[i][:, 0], ref[k]), k


def test_per_env_loads_and_seeds():
    loads = np.array([100.0, 200.0, 300.0, 400.0])
    out = TraceGenerator(4, 14, 5, loads, base_seed=9).next(100)
    for e in range(4):
        ref, _, _ = orc.generate_trace_python(14, 5, float(loads[e]), 10800.0, 9 + e, 100)
        for i, k in enumerate(TRACE_KEYS):
            assert np.array_equal(out[i][:, e], ref[k])


def test_properties():
    out = TraceGenerator(8, 28, 5, 300.0, base_seed=3).next(2000)
    src, dst, rate, arr, hold = out
    assert (src != dst).all() and src.max() < 28 and dst.max() < 28 and rate.max() < 5
    assert (np.diff(arr, axis=0) >= 0).all() and (hold >= 0).all()
    assert abs(np.diff(arr, axis=0).mean() - 10800.0 / 300.0) < 1.5


def test_philox_known_answers_and_restatement_properties():
    """The numpy restatement of the on-device generator: Philox4x32-10 against the Random123 known-answer vectors
    (kat_vectors: philox4x32 10), then determinism / shard independence of the request streams built on it."""
    from oracle import oracle as orc
    from optical_networking_gym_b200.tracegen import choice_tables

    kat = [((0, 0, 0, 0), (0, 0), "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for c, k, want in kat:
        out = orc.philox4x32_10([np.array([x]) for x in c], k)
        assert " ".join("%08x" % int(o[0]) for o in out) == want
    tabs = choice_tables(14, 5)
    a = orc.generate_trace_philox(12, 40, 210.0, 77, *tabs)
    b = orc.generate_trace_philox(4, 40, 210.0, 77, *tabs, env_offset=8)
    assert np.array_equal(a[0][:, 8:], b[0]) and np.array_equal(a[3][:, 8:], b[3])   # streams follow the global env index
    c = orc.generate_trace_philox(12, 15, 210.0, 77, *tabs, pos0=25)
    assert np.array_equal(a[1][25:], c[1])                                               # and the request index
    assert (a[0] != a[1]).all() and (np.diff(a[3].astype(np.float64), axis=0) >= 0).all()


def test_skewed_bit_rate_mix_vs_reference():
    """bit_rate_probabilities = [0.5, 0.3, 0.2] (qrmsa.pyx:256-257, :1089): the generator's `choices` bisection against
    the stream the reference drew for the margin / bit-rate-mix recording."""
    import os
    from helpers import GOLDEN
    from optical_networking_gym_b200.tables import StaticTables

    tag = "var_margin_nobel-eu"
    tb = StaticTables.load(os.path.join(GOLDEN, f"tables_{tag}.npz"))
    g = load_golden("run_" + tag)
    tg = TraceGenerator(1, tb.n_nodes, tb.n_rates, float(g["meta_load"]), base_seed=int(g["meta_seed"]),
                        bit_rate_probabilities=[0.5, 0.3, 0.2])
    out = tg.next(len(g["src"]))
    for i, k in enumerate(TRACE_KEYS):
        assert np.array_equal(out[i][:, 0], g[k]), k
    assert abs((g["rate"] == 0).mean() - 0.5) < 0.05


def test_randint_rates_vs_cpython():
    """bit_rate_selection="continuous" (qrmsa.pyx:246-254): the fifth draw of a request is rng.randint(lower, higher),
    restated in csrc/tracegen.cpp (_randbelow_with_getrandbits) and checked against CPython's own random.Random."""
    from oracle import oracle as orc

    g = TraceGenerator(4, 14, 76, 210.0, base_seed=321, n_threads=1, randint_rates=(25, 100))
    tr = g.next(400)
    for e in range(4):
        ref, _, _ = orc.generate_trace_python(14, 76, 210.0, 10800.0, 321 + e, 400, randint_rates=(25, 100))
        for k, a in zip(("src", "dst", "rate", "arrival", "holding"), tr):
            assert np.array_equal(ref[k], a[:, e]), (e, k)
    assert tr[2].min() >= 0 and tr[2].max() <= 75
    g.close()
