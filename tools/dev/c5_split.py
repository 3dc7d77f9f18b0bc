"""Where a config-5 rollout step goes: observation kernel, step_action, the policy's GEMMs, the masked sampler."""
import sys, time, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import load_tables
from examples.ppo_rollout import make_policy, policy_logits
from optical_networking_gym_b200.env import BatchedQRMSAEnv
from optical_networking_gym_b200.sampling import sample_masked_actions

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
tb = load_tables("nsfnet", 320)
env = BatchedQRMSAEnv(tb, n, num_spectrum_resources=320, episode_length=600, load=210.0, bit_rates=(10, 40, 100, 400, 1000),
                      launch_power_dbm=1.0, gen_observation=True, seed=10)
policy = make_policy(env.observation_space.shape[0], env.action_space.n, torch.device("cuda"))
eng = env.engine
eng.step_first_fit(300)
obs, mask = env._obs, env._mask
eng.observation(obs, mask)

def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3

with torch.no_grad():
    logits = policy_logits(policy, obs)
    act = sample_masked_actions(logits, mask, 1, 0)
    print(f"observation kernel      {timeit(lambda: eng.observation(obs, mask)):.3f} ms")
    print(f"policy GEMMs (bf16)     {timeit(lambda: policy_logits(policy, obs)):.3f} ms")
    print(f"masked sampler          {timeit(lambda: sample_masked_actions(logits, mask, 1, 0)):.3f} ms")
    rw = torch.zeros(n, dtype=torch.float32, device="cuda"); st = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rej = torch.full((n,), env.action_space.n - 1, dtype=torch.int64, device="cuda")
    print(f"step_action (reject)    {timeit(lambda: eng.step_action(rej, rw, st, None, None), 5):.3f} ms")
    def full():
        lg = policy_logits(policy, obs); a = sample_masked_actions(lg, mask, 1, 0); env.step(a)
    print(f"whole step              {timeit(full, 20):.3f} ms for {n} envs = {n / timeit(full, 20) * 1e3:,.0f} env-steps/s")
