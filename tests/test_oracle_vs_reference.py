"""Live check of the oracle against the COMPILED reference (oracle/_ref), where it is present: a seed and
load that no committed fixture holds.  Skipped when oracle/_ref was not built."""
import numpy as np
import pytest

from helpers import TRACE_KEYS
from oracle import oracle as orc
from oracle import ref_harness as rh
from optical_networking_gym_b200.tables import StaticTables

pytestmark = pytest.mark.reference


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_oracle_matches_live_reference():
    topo = rh.make_topology("nsfnet")
    out, _ = rh.run_first_fit(topo, 4242, 250, n_slots=320, load=420.0)
    tb = StaticTables.from_topology(topo, num_spectrum_resources=320, bit_rates=(10, 40, 100, 400, 1000),
                                    launch_power_dbm=1.0)
    o = orc.OracleEnv(tb, 251)
    o.reset(*[out[k] for k in TRACE_KEYS])
    r = o.run_first_fit(250)
    assert np.array_equal(r["action"], out["action"])
    assert np.abs(r["gsnr"] - out["gsnr"]).max() < 1e-9
    assert np.array_equal(o.slots(), out["final_slots"])
    ref, _, _ = orc.generate_trace_python(tb.n_nodes, 5, 420.0, 10800.0, 4242, 251)
    for k in TRACE_KEYS:
        assert np.array_equal(ref[k], out[k])
