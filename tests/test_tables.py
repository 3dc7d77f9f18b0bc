import numpy as np
import pytest

from helpers import load_tables
from optical_networking_gym_b200.tables import StaticTables


def test_roundtrip(tmp_path):
    tb = load_tables("nobel-eu", 320)
    p = str(tmp_path / "t.npz")
    tb.save(p)
    tb2 = StaticTables.load(p)
    for k in StaticTables._ARRAYS:
        assert np.array_equal(getattr(tb, k), getattr(tb2, k))
    for k in StaticTables._SCALARS + StaticTables._FLOATS:
        assert getattr(tb, k) == getattr(tb2, k)


@pytest.mark.parametrize("topo,S,N,E,hmax", [("nsfnet", 320, 14, 22, 9), ("nobel-eu", 320, 28, 41, 9),
                                             ("germany50", 640, 50, 88, 14)])
def test_dimensions_match_survey(topo, S, N, E, hmax):
    tb = load_tables(topo, S)
    assert (tb.n_nodes, tb.n_links, tb.n_slots, tb.k_paths, tb.n_mods) == (N, E, S, 5, 6)
    assert tb.max_hops == hmax
    hops = tb.path_hops.reshape(N, N, 5)
    for a in range(N):
        assert (hops[a, a] == 0).all()
        for b in range(N):
            if a != b:
                assert (hops[a, b] >= 1).all()                       # every pair has exactly k paths
                assert np.array_equal(hops[a, b], hops[b, a])         # ksp[n1,n2] is ksp[n2,n1] (topology.pyx:354-355)
    # slots needed: ceil(rate / (SE * 12.5))  
    need = tb.slots_needed.reshape(5, 6)
    assert need[4].tolist() == [80, 40, 27, 20, 16, 14] and need[0].tolist() == [1] * 6
    assert tb.n_actions == 5 * 6 * S + 1


def test_from_topology_matches_golden_tables():
    from oracle import ref_harness as rh

    if not rh.available():
        pytest.skip("oracle/_ref not built")
    topo = rh.make_topology("nsfnet")
    tb = StaticTables.from_topology(topo, num_spectrum_resources=320, bit_rates=(10, 40, 100, 400, 1000),
                                    launch_power_dbm=1.0)
    g = load_tables("nsfnet", 320)
    for k in StaticTables._ARRAYS:
        assert np.array_equal(getattr(tb, k), getattr(g, k)), k
