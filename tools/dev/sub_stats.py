import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from optical_networking_gym_b200.engine import Engine
from optical_networking_gym_b200.tables import StaticTables
from optical_networking_gym_b200.tracegen import TraceGenerator
tb = StaticTables.load('tests/golden/tables_nobel-eu_320.npz')
n_envs, T = 8192, 1400
tg = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, 300.0, base_seed=50)
tr = tg.next(T)
eng = Engine(tb, n_envs, T, device=0)
eng.reset(); eng.load_trace_host(*tr)
eng.step_first_fit(1000)
c0 = np.array(eng.counters(), dtype=np.int64).reshape(-1)
eng.step_first_fit(128)
c1 = np.array(eng.counters(), dtype=np.int64).reshape(-1)
d = c1 - c0
steps = d[0]
print('steps', steps, 'per env-step: rounds*8', d[25]*8/steps, 'cands/round', d[26]/d[25], 'alive/round', d[29]/d[25], 'search iters/round', d[27]/d[25], 'gn iters/round', d[28]/d[25], 'evals/step', d[9]/steps)
