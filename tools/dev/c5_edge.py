"""Find and explain a masked action that step() refuses in the config-5 rollout (bench rank r of N): is it the mask's
rounding edge (osnr.pyx:366: valid down to ~1e-9 dB below the threshold) or something else?"""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_tables
from examples.ppo_rollout import make_policy, policy_logits
from optical_networking_gym_b200.env import BatchedQRMSAEnv
from optical_networking_gym_b200.sampling import sample_masked_actions

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_envs, n_steps, Wm = 16384, 1024, 3
tb = load_tables("nsfnet", 320)
env = BatchedQRMSAEnv(tb, n_envs, num_spectrum_resources=320, episode_length=n_steps + Wm + 1, load=210.0,
                      bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, gen_observation=True, seed=10 + rank * n_envs,
                      reset=False)
policy = make_policy(env.observation_space.shape[0], env.action_space.n, torch.device("cuda"), seed=rank)
env.reset()
obs, mask = env._obs, env.action_masks()
S, M = 320, 6
with torch.no_grad():
    for t in range(n_steps + Wm):
        seed, step = (1000 + rank, t) if t < Wm else (1000 + rank, t)   # bench: warm-up steps 0..2 then first_step=Wm
        a = sample_masked_actions(policy_logits(policy, obs), mask, seed, step)
        mask_before = mask.clone()
        obs, rw, term, _, info = env.step(a)
        st = info["status"]
        bad = torch.nonzero((st != 0) & (st != 1)).flatten()
        if len(bad):
            e = int(bad[0]); act = int(a[e])
            p, mi, s = act // (M * S), (act // S) % M, act % S
            m = M - 1 - mi
            print(f"step {t}: env {e} action {act} (path {p}, modulation {m}, slot {s}) status {int(st[e])}, mask bit {int(mask_before[e, act])}")
            src, dst, rate, arr, hold, words = env.engine.env_log(e)
            cur = int(env.engine.env_state()[e, 0])
            n = int(tb.slots_needed.reshape(-1, M)[rate[cur], m])
            g = env.engine.probe_gsnr(e, int(src[cur]), int(dst[cur]), p, s, n)
            thr = float(tb.mod_min_osnr[m])
            print(f"  request {cur}: GSNR {g!r} dB, threshold {thr!r} dB, difference {g - thr:.3e} dB; "
                  f"mask rule round((g-thr)/|thr|, 10) = {round((g - thr) / abs(thr), 10)!r}")
            break
        mask = info["mask"]
    else:
        print("no refused action in this rollout")
