import sys, time, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_tables
from optical_networking_gym_b200 import _lib
if len(sys.argv) > 2:   # kernel experiments: a prebuilt library (tools/dev/build_variant.sh)
    import os
    _lib.LIB_PATH = os.path.abspath(sys.argv[2]); _lib.needs_build = lambda: False
from optical_networking_gym_b200.env import BatchedQRMSAEnv
tb = load_tables("nsfnet", 320)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
env = BatchedQRMSAEnv(tb, n, num_spectrum_resources=320, episode_length=400, load=210.0, bit_rates=(10, 40, 100, 400, 1000),
                      launch_power_dbm=1.0, gen_observation=False, seed=10, request_source="device")
env.step_first_fit(300)
obs = torch.zeros((n, env.observation_space.shape[0]), dtype=torch.float32, device="cuda")
mask = torch.zeros((n, env.action_space.n), dtype=torch.uint8, device="cuda")
for _ in range(2): env.engine.observation(obs, mask)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): env.engine.observation(obs, mask)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
import hashlib
dig = hashlib.sha256(obs.cpu().numpy().tobytes() + mask.cpu().numpy().tobytes()).hexdigest()[:16]
if len(sys.argv) > 3:   # keep a sample of the outputs: variants are compared here afterwards (features to 2e-6, mask bit for bit)
    np.savez_compressed(sys.argv[3], obs=obs[:1024].cpu().numpy(), mask=np.packbits(mask[:1024].cpu().numpy(), axis=1))
print(f"{sys.argv[2] if len(sys.argv) > 2 else 'in-tree'} sha {dig} k_observation: {n} envs in {dt*1e3:.2f} ms = {n/dt:,.0f} env/s; valid actions per env {float(mask.sum())/n:.0f}")
