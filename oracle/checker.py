"""Replay sampled envs of a large device batch through the oracle.   TEST INFRASTRUCTURE ONLY.

Used by tests/, __graft_entry__.smoke() and bench.py's `parity_sample` legs: after a (timed) device run, the request
stream and the decision log of K envs are copied back (`Engine.env_log`), the same streams are replayed from an empty
network through oracle/qrmsa_oracle.c, and decisions, accept flags and final slot bitmaps are compared under the
north_star rule (bit-exact; a first divergence is excused only if THAT step carries the near-threshold flag).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc

ACTION_MASK = 0x00FFFFFF
FLAG_NEAR_THRESHOLD = 0x80000000
FLAG_ACCEPTED = 0x20000000
FLAG_DECIDED = 0x40000000


def spread_sample(n_envs: int, k: int):
    """k env indices spread over [0, n_envs): both ends, warp / CTA boundaries and an even stride between."""
    k = min(k, n_envs)
    base = {0, n_envs - 1, min(31, n_envs - 1), min(32, n_envs - 1), min(1023, n_envs - 1), min(1024, n_envs - 1)}
    step = max(n_envs // max(k, 1), 1)
    i = step // 2
    while len(base) < k and i < n_envs:
        base.add(i)
        i += step
    j = 1
    while len(base) < k:
        base.add(j % n_envs)
        j += 7919
    return sorted(base)[:k] if len(base) > k else sorted(base)


def replay_first_fit(tables, engine, envs, n_decided: int, policy: int = 0, check_bitmaps: bool = True) -> dict:
    """Compare `n_decided` decisions of each env in `envs` (device run from reset) with the oracle.
    Returns {"envs", "steps", "mismatches", "excused", "bitmap_mismatches", "near_threshold_flags"}."""
    from optical_networking_gym_b200.engine import unpack_bitmaps   # plumbing only (bit unpacking)

    mism = exc = bm_bad = flags = 0
    for e in envs:
        src, dst, rate, arr, hold, words = engine.env_log(int(e), 0, n_decided + 1)
        act = (words[:n_decided] & ACTION_MASK).astype(np.int64)
        flg = (words[:n_decided] & FLAG_NEAR_THRESHOLD) != 0
        acc = (words[:n_decided] & FLAG_ACCEPTED) != 0
        assert ((words[:n_decided] & FLAG_DECIDED) != 0).all(), f"env {e}: undecided request inside the compared span"
        flags += int(flg.sum())
        o = orc.OracleEnv(tables, n_decided + 1)
        o.reset(src, dst, rate, arr, hold)
        ref = o.run_first_fit(n_decided, log_qot=False, policy=policy)
        d = np.flatnonzero(ref["action"] != act)
        if len(d):
            if flg[int(d[0])]:
                exc += 1
            else:
                mism += 1
            continue
        if not np.array_equal(ref["accepted"].astype(bool), acc):
            mism += 1
            continue
        if check_bitmaps:
            bm = unpack_bitmaps(engine.export_bitmaps(int(e), 1), tables.n_slots)[0]
            if not np.array_equal(o.slots(), bm):
                bm_bad += 1
    return {"envs": len(list(envs)), "steps": int(n_decided), "mismatches": mism, "excused": exc,
            "bitmap_mismatches": bm_bad, "near_threshold_flags": flags}


def replay_actions(tables, engine, envs, actions, statuses, episode_length: int, final_obs=None, final_mask=None) -> dict:
    """RL path: `actions` int64 [steps][len(envs)] were applied by qrmsa_step_action to the sampled envs (device
    `statuses` uint8 [steps][len(envs)]); replay them through the oracle's step and compare status per step, the final
    slot bitmaps and -- when given -- the final observation (|d| < 2e-6) and action mask (exact)."""
    from optical_networking_gym_b200.engine import unpack_bitmaps

    n_steps = actions.shape[0]
    mism = bm_bad = obs_bad = mask_bad = 0
    for j, e in enumerate(envs):
        n_req = int(engine.n_loaded)
        src, dst, rate, arr, hold, _ = engine.env_log(int(e), 0, n_req)
        o = orc.OracleEnv(tables, n_req)
        o.reset(src, dst, rate, arr, hold)
        ok = True
        for s in range(n_steps):
            st, _, _, _ = o.step_action(int(actions[s, j]), episode_length)
            if st != int(statuses[s, j]):
                ok = False
                break
        if not ok:
            mism += 1
            continue
        bm = unpack_bitmaps(engine.export_bitmaps(int(e), 1), tables.n_slots)[0]
        if not np.array_equal(o.slots(), bm):
            bm_bad += 1
            continue
        if final_obs is not None:
            ob, mk = o.observation()
            if np.abs(ob - final_obs[j]).max() > 2e-6:
                obs_bad += 1
            if not np.array_equal(mk, final_mask[j]):
                mask_bad += 1
    return {"envs": len(list(envs)), "steps": int(n_steps), "mismatches": mism, "excused": 0, "bitmap_mismatches": bm_bad,
            "obs_mismatches": obs_bad, "mask_mismatches": mask_bad}
