import sys, time, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_tables
from optical_networking_gym_b200 import _lib
if len(sys.argv) > 1:   # kernel experiments: a prebuilt library (tools/dev/build_variant.sh)
    import os
    _lib.LIB_PATH = os.path.abspath(sys.argv[1]); _lib.needs_build = lambda: False
from optical_networking_gym_b200.engine import Engine
tb = load_tables("nobel-eu", 320)
n_envs, n = (int(sys.argv[2]) if len(sys.argv) > 2 else 4096), 400
eng = Engine(tb, n_envs, n + 1)
eng.reset(); eng.generate_trace(n + 1, 300.0, seed=1)
eng.step_heuristic("highest_snr", 300)   # (launch 0: the fill; launch 1, timed below, is the one the ncu capture takes)
torch.cuda.synchronize()
c0 = eng.counters_dict()
t0 = time.time()
eng.step_heuristic("highest_snr", 100)
c = eng.counters_dict()
dt = time.time() - t0
import hashlib
dig = hashlib.sha256(eng.actions_host(0, 400).tobytes()).hexdigest()[:16]
print(sys.argv[1] if len(sys.argv) > 1 else "in-tree", "actions sha", dig, "highest_snr:", n_envs * 100 / dt, "env-steps/s;", (c["gn_evals"] - c0["gn_evals"]) / (c["decided"] - c0["decided"]), "QoT checks per request; accepted", c["accepted"] / c["decided"])
