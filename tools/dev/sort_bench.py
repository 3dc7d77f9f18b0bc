import sys, time, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_tables
from optical_networking_gym_b200.engine import Engine
tb = load_tables("nobel-eu", 320)
eng = Engine(tb, 65536, 3817)
eng.reset()
for _ in range(2): eng.generate_trace(3817, 300.0, seed=1)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): eng.generate_trace(3817, 300.0, seed=1)
torch.cuda.synchronize()
print("generate + schedule:", (time.perf_counter() - t0) / 3 * 1e3, "ms")
