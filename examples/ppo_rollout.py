#!/usr/bin/env python
"""BASELINE config 5: rollout with on-device action masks and a PyTorch policy consuming them.

    python examples/ppo_rollout.py [--envs 16384] [--steps 1024] [--topology nsfnet]
    torchrun --nproc-per-node N examples/ppo_rollout.py ...        # --envs per GPU, one policy replica per rank

Mirrors the consumer of reference examples/ONDM_2025/train_multi_masked_ppo.py (MaskablePPO over
SubprocVecEnv of 14 envs, policy pi=[512,256,128], :410-444): observation float32[368] and mask uint8[9601]
stay in HBM, the MLP 368->512->256->128->9601 samples a masked categorical action, `BatchedQRMSAEnv.step`
applies it, and the rollout buffer (actions, rewards, statuses) is filled on the device.  Policy weights are random
(no training here): this measures the environment side of the loop with the policy in it.  Across ranks the only
traffic is the all-reduce of the counters at the end of the rollout.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_policy(obs_dim: int, n_actions: int, device, seed: int = 0):
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    # the action head is padded to a multiple of 8 outputs (9,601 -> 9,608): an odd row length takes the GEMM off its
    # aligned path; `policy_logits` returns the [n_envs, n_actions] view of the padded rows and the sampler reads it in place
    padded = (n_actions + 7) // 8 * 8
    layers, dims = [], [obs_dim, 512, 256, 128, padded]          # train_multi_masked_ppo.py:440-441
    for i in range(4):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        with torch.no_grad():
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) / dims[i] ** 0.5)
            lin.bias.zero_()
        layers += [lin] + ([torch.nn.Tanh()] if i < 3 else [])
    net = torch.nn.Sequential(*layers).to(device).bfloat16()
    net.n_actions = n_actions
    return net


def policy_logits(policy, obs):
    """bf16 logits [n_envs, n_actions] (a view of the padded action head's output)."""
    return policy(obs.bfloat16())[:, : policy.n_actions]


def rollout(env, policy, n_steps: int, buffers: dict = None, host_reward=None, seed: int = 0, first_step: int = 0,
            torch_sampler: bool = False):
    """n_steps of obs -> policy -> masked categorical sample -> env.step for every env of `env` (a BatchedQRMSAEnv with
    gen_observation=True, already reset).  `buffers` (optional) receives the rollout buffer rows on the device:
    "action" int64 [n_steps, n_envs], "status" uint8, "reward" float32.  `host_reward` (optional, pinned float32
    [n_steps, n_envs]) gets every step's rewards by an asynchronous device->host copy (the end-to-end leg)."""
    import torch

    from optical_networking_gym_b200.sampling import sample_masked_actions

    obs, mask = env._obs, env.action_masks()
    total_reward = torch.zeros((), device=obs.device, dtype=torch.float64)
    with torch.no_grad():
        for t in range(n_steps):
            logits = policy_logits(policy, obs)                  # bf16 [n_envs, n_actions], read in place by the sampler
            if torch_sampler:   # the same sample spelled with torch ops (seven passes over an fp32 copy of the logits)
                lf = logits.float().masked_fill_(mask == 0, float("-inf"))
                action = (lf - torch.empty_like(lf).exponential_().log_()).argmax(dim=1)
            else:               # one pass over the logits and the mask (Gumbel-max, Philox keyed by seed / step / env / action)
                action = sample_masked_actions(logits, mask, seed, first_step + t)
            obs, reward, term, trunc, info = env.step(action)
            mask = info["mask"]
            if buffers is not None:
                buffers["action"][t].copy_(action)
                buffers["status"][t].copy_(info["status"])
                buffers["reward"][t].copy_(reward)
            if host_reward is not None:
                host_reward[t].copy_(reward, non_blocking=True)
            total_reward += reward.sum()
    return total_reward


def main():
    import torch
    import torch.distributed as dist

    from optical_networking_gym_b200 import sharding
    from optical_networking_gym_b200.env import BatchedQRMSAEnv
    from optical_networking_gym_b200.tables import StaticTables

    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384, help="envs per GPU")
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--topology", default="nsfnet")
    ap.add_argument("--load", type=float, default=210.0)   # train_multi_masked_ppo.py:381
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tb = StaticTables.load(os.path.join(ROOT, "tests", "golden", f"tables_{args.topology}_320.npz"))
    env = BatchedQRMSAEnv(tb, args.envs, num_spectrum_resources=320, episode_length=args.steps + 1, load=args.load,
                          bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, gen_observation=True,
                          seed=10 + rank * args.envs, device=local)
    dev = torch.device("cuda", local)
    policy = make_policy(env.observation_space.shape[0], env.action_space.n, dev, seed=rank)
    env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    total_reward = rollout(env, policy, args.steps)
    counters = sharding.allreduce_counters(env.engine.counters())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        c = counters.sum(0)
        print(f"{world} GPU(s) x {args.envs} envs x {args.steps} steps on {args.topology}: "
              f"{world * args.envs * args.steps / dt:,.0f} env-steps/s, accepted {int(c[1])}/{int(c[0])}, "
              f"status errors {int(c[15])}, mean reward on rank 0 {float(total_reward) / (args.envs * args.steps):.4f}")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
