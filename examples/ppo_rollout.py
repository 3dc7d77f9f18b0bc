#!/usr/bin/env python
"""BASELINE config 5: rollout with on-device action masks and a PyTorch policy consuming them.

    python examples/ppo_rollout.py [--envs 16384] [--steps 64] [--topology nsfnet]

Mirrors the consumer of reference examples/ONDM_2025/train_multi_masked_ppo.py (MaskablePPO over
SubprocVecEnv of 14 envs, policy pi=[512,256,128], :410-444): observation float32[368] and mask uint8[9601]
stay in HBM, the MLP 368->512->256->128->9601 samples a masked categorical action, `BatchedQRMSAEnv.step`
applies it.  Policy weights are random (no training here): this measures the environment side of the loop.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from optical_networking_gym_b200.env import BatchedQRMSAEnv
    from optical_networking_gym_b200.tables import StaticTables

    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--topology", default="nsfnet")
    ap.add_argument("--load", type=float, default=210.0)   # train_multi_masked_ppo.py:381
    args = ap.parse_args()
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", f"tables_{args.topology}_320.npz"))
    env = BatchedQRMSAEnv(tb, args.envs, num_spectrum_resources=320, episode_length=args.steps + 1, load=args.load,
                          bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, gen_observation=True, seed=10)
    obs_dim, n_act = env.observation_space.shape[0], env.action_space.n
    dev = torch.device("cuda")
    policy = torch.nn.Sequential(torch.nn.Linear(obs_dim, 512), torch.nn.Tanh(), torch.nn.Linear(512, 256), torch.nn.Tanh(),
                                 torch.nn.Linear(256, 128), torch.nn.Tanh(), torch.nn.Linear(128, n_act)).to(dev).bfloat16()
    obs, info = env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_env = 0.0
    total_reward = torch.zeros((), device=dev)
    with torch.no_grad():
        for _ in range(args.steps):
            logits = policy(obs.bfloat16()).float()
            logits.masked_fill_(info["mask"] == 0, float("-inf"))
            # masked categorical sample by the Gumbel-max trick: argmax(logits - log E), E ~ Exp(1); no normalisation pass
            action = (logits - torch.empty_like(logits).exponential_().log_()).argmax(dim=1)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            obs, reward, term, trunc, info = env.step(action)
            torch.cuda.synchronize(); t_env += time.perf_counter() - t1
            total_reward += reward.sum()
    dt = time.perf_counter() - t0
    c = env.counters()
    print(f"{args.envs} envs x {args.steps} steps on {args.topology}: {args.envs * args.steps / dt:,.0f} env-steps/s "
          f"(env side {args.envs * args.steps / t_env:,.0f}/s), accepted {c['accepted']}/{c['decided']}, "
          f"status errors {c['errors']}, mean reward {float(total_reward) / (args.envs * args.steps):.4f}")
    env.close()


if __name__ == "__main__":
    main()
