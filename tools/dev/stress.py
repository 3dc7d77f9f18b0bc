"""Long full-size run at high load: error counters, bitmap <-> channel-list consistency, device-vs-oracle on samples."""
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_tables
from optical_networking_gym_b200 import _lib
from optical_networking_gym_b200.engine import Engine, unpack_bitmaps
from oracle import oracle as orc
tb = load_tables("nobel-eu", 320)
n_envs, n = 65536, 3001
eng = Engine(tb, n_envs, n)
eng.reset(); eng.generate_trace(n, 600.0, seed=2024)
t0 = time.time()
for _ in range(6):
    eng.step_first_fit(500)
c = eng.counters_dict()
print("decided", c["decided"], "accept", c["accepted"] / c["decided"], "errors", c["errors"], "sec", time.time() - t0)
assert c["decided"] == n_envs * 3000 and c["errors"] == 0
tr = eng.trace_host()
words = eng.actions_host(0, n - 1)
acts = (words & _lib.ACTION_MASK).T
flag = ((words.view(np.uint32) & _lib.FLAG_NEAR_THRESHOLD) != 0).T
for e in (0, 777, 65535):
    o = orc.OracleEnv(tb, n)
    o.reset(*[a[:, e] for a in tr])
    ref = o.run_first_fit(n - 1, log_qot=False)
    same = np.array_equal(ref["action"], acts[e])
    if not same:
        d = int(np.flatnonzero(ref["action"] != acts[e])[0])
        assert flag[e, : d + 1].any(), (e, d)
        print("env", e, "diverges at a flagged step", d)
    else:
        assert np.array_equal(o.slots(), unpack_bitmaps(eng.export_bitmaps(e, 1), 320)[0])
        print("env", e, "3000 decisions and the final bitmaps identical to the oracle")
print("stress ok")
