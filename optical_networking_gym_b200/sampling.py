"""Masked categorical sampling on the device: the consumer side of `action_masks()`.

The reference's RL consumer is sb3_contrib's MaskablePPO over the wrapper's `action_masks()`
(wrappers/qrmsa_gym.py:74-75, examples/ONDM_2025/train_multi_masked_ppo.py:410-444): sample
a ~ Categorical(softmax(logits) restricted to mask == 1).  `sample_masked_actions` draws that sample for every env of a
batch in ONE pass over the policy's logits (float32 or bfloat16, left where the policy wrote them) and the uint8 mask
`qrmsa_observation` wrote -- hand-written sm_100a kernel behind `qrmsa_sample_masked_actions` (include/qrmsa_b200.h),
Gumbel-max over a Philox stream keyed by (seed, step) and counted by (env, action).
"""
from __future__ import annotations

from . import _lib
from ._lib import check


def sample_masked_actions(logits, mask, seed: int, step: int, out=None, stream=None):
    """logits: CUDA float32 / bfloat16 [n_envs, n_actions] (rows may be strided); mask: CUDA uint8 [n_envs, n_actions];
    returns int64 CUDA [n_envs].  Deterministic in (seed, step, env, action)."""
    import torch

    if logits.dim() != 2 or mask.shape != logits.shape or not logits.is_cuda or not mask.is_cuda:
        raise ValueError("logits and mask must be CUDA tensors of the same [n_envs, n_actions] shape")
    if mask.dtype != torch.uint8 or logits.stride(1) != 1 or mask.stride(1) != 1:
        raise ValueError("mask must be uint8 and both tensors contiguous along the action axis")
    if logits.dtype == torch.float32:
        dt = 0
    elif logits.dtype == torch.bfloat16:
        dt = 1
    else:
        raise ValueError("logits must be float32 or bfloat16")
    n_envs, n_actions = logits.shape
    if out is None:
        out = torch.empty(n_envs, dtype=torch.int64, device=logits.device)
    dev = logits.device.index if logits.device.index is not None else torch.cuda.current_device()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    check(_lib.load().qrmsa_sample_masked_actions(logits.data_ptr(), dt, mask.data_ptr(), int(n_envs), int(n_actions),
                                                  int(logits.stride(0)), int(mask.stride(0)), int(seed) & (2 ** 64 - 1),
                                                  int(step), out.data_ptr(), int(dev), getattr(st, "cuda_stream", st) or None))
    return out
