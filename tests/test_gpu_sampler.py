"""qrmsa_sample_masked_actions (csrc/qrmsa_sampler.cuh) against its numpy restatement: the sampled action is always a
valid one, equals the restatement's argmax (ties in float32 `logf` rounding aside), is reproducible and follows the
masked softmax distribution."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_sampler_matches_numpy_restatement(dtype):
    import torch
    from optical_networking_gym_b200.sampling import sample_masked_actions

    rng = np.random.default_rng(3)
    n_envs, n_actions = 777, 9601                      # the NSFNET action space of BASELINE config 5, odd row length
    logits = torch.from_numpy(rng.normal(0, 2, (n_envs, n_actions)).astype(np.float32)).cuda().to(getattr(torch, dtype))
    mask_np = (rng.random((n_envs, n_actions)) < 0.03).astype(np.uint8)
    mask_np[:, -1] = 1                                 # the reject action is always valid (qrmsa.pyx:766)
    mask_np[5] = 0; mask_np[5, -1] = 1                 # only the reject action
    mask_np[6] = 0                                     # no valid action at all: cannot come from k_observation
    mask = torch.from_numpy(mask_np).cuda()
    a = sample_masked_actions(logits, mask, seed=1234, step=7).cpu().numpy()
    b = sample_masked_actions(logits, mask, seed=1234, step=7).cpu().numpy()
    c = sample_masked_actions(logits, mask, seed=1234, step=8).cpu().numpy()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert (mask_np[np.arange(n_envs), a] == 1)[np.arange(n_envs) != 6].all()
    assert a[5] == n_actions - 1 and a[6] == n_actions - 1
    ref, keys = orc.sample_masked_actions_numpy(logits.float().cpu().numpy(), mask_np, 1234, 7)
    diff = np.flatnonzero(a != ref)
    assert len(diff) <= n_envs // 100                  # logf on the device vs numpy: last-bit differences can swap a near tie
    for e in diff:
        assert abs(keys[e, a[e]] - keys[e, ref[e]]) < 1e-4


def test_sampler_distribution_and_strides():
    import torch
    from optical_networking_gym_b200.sampling import sample_masked_actions

    n_envs, n_actions = 20000, 12
    lg = torch.tensor([0.0, 1.0, -1.0, 2.0, 0.5, 0.0, 3.0, -2.0, 1.5, 0.0, 0.2, 0.1])
    big = torch.zeros((n_envs, 16), dtype=torch.float32, device="cuda")      # row stride 16 > n_actions
    big[:, :n_actions] = lg.cuda()
    logits = big[:, :n_actions]
    mask = torch.ones((n_envs, n_actions), dtype=torch.uint8, device="cuda")
    mask[:, 6] = 0                                                              # the most likely action is masked out
    counts = np.zeros(n_actions)
    for step in range(5):
        a = sample_masked_actions(logits, mask, seed=99, step=step).cpu().numpy()
        counts += np.bincount(a, minlength=n_actions)
    assert counts[6] == 0
    p = np.exp(lg.numpy()); p[6] = 0; p /= p.sum()
    n = counts.sum()
    z = (counts - n * p)[p > 0] / np.sqrt(n * p * (1 - p))[p > 0]
    assert np.abs(z).max() < 5.0, z
