"""Observation + action-mask kernel (qrmsa_observation) vs the reference recording and vs the oracle."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, TRACE_KEYS, load_golden, load_tables
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

OBS_TOL = 2e-6   # float32 features; FP64 sums are accumulated in a different order than numpy's pairwise sum


def _mask_diff_allowed(tb, o, mask_dev, mask_ref):
    """Mask entries may differ only where the candidate's GSNR is within 1e-3 dB of its threshold."""
    bad = np.flatnonzero(mask_dev != mask_ref)
    S, M = tb.n_slots, tb.n_mods
    for a in bad:
        p, mi, s = a // (M * S), (a // S) % M, a % S
        m = M - 1 - mi
        n = int(tb.slots_needed.reshape(-1, M)[o_rate(o), m])
        g = o.probe_gsnr(o_src(o), o_dst(o), p, s, n)
        assert abs(g - tb.mod_min_osnr[m]) < 1e-3, f"mask differs at action {a} with GSNR {g} vs {tb.mod_min_osnr[m]}"
    return len(bad)


def o_src(o): return int(o._trace[0][o.current_request])
def o_dst(o): return int(o._trace[1][o.current_request])
def o_rate(o): return int(o._trace[2][o.current_request])


@pytest.mark.parametrize("tag,topo", [("obs_nsfnet_320_l210_s21", "nsfnet"), ("obs_nobel-eu_320_l400_s8", "nobel-eu")])
def test_observation_vs_reference_recording(tag, topo):
    import torch
    from optical_networking_gym_b200.engine import Engine

    if not os.path.exists(os.path.join(GOLDEN, tag + ".npz")):
        pytest.skip("fixture not generated")
    g = load_golden(tag)
    tb = load_tables(topo, 320)
    n_req, n_act = len(g["src"]), int(g["n_actions"])
    mask_ref = np.unpackbits(g["mask"], axis=1)[:, :n_act]
    eng = Engine(tb, 1, n_req)
    assert eng.observation_dims() == (g["obs"].shape[1], n_act)
    eng.reset(); eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    o = orc.OracleEnv(tb, n_req)
    o.reset(*[g[k] for k in TRACE_KEYS])
    dev = torch.device("cuda")
    obs = torch.zeros((1, g["obs"].shape[1]), dtype=torch.float32, device=dev)
    mask = torch.zeros((1, n_act), dtype=torch.uint8, device=dev)
    a = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.zeros(1, dtype=torch.uint8, device=dev)
    rw = torch.zeros(1, dtype=torch.float32, device=dev)
    n_diff = 0
    for t in range(len(g["action"]) + 1):
        mask.fill_(7)   # every entry must be written by the kernel
        eng.observation(obs, mask)
        torch.cuda.synchronize()
        assert np.abs(obs.cpu().numpy()[0] - g["obs"][t]).max() < OBS_TOL, f"obs at step {t}"
        n_diff += _mask_diff_allowed(tb, o, mask.cpu().numpy()[0], mask_ref[t])
        if t < len(g["action"]):
            a[0] = int(g["action"][t])
            eng.step_action(a, rw, st, None, None)
            torch.cuda.synchronize()
            assert int(st[0]) in (0, 1) and float(rw[0]) == pytest.approx(float(g["reward"][t]), abs=1e-6)
            o.step_action(int(g["action"][t]), n_req)
    assert n_diff <= 2
    eng.close()


@pytest.mark.parametrize("topo,n_slots,load", [("germany50", 640, 800.0), ("var_k3_nsfnet", 160, 150.0), ("nsfnet", 100, 70.0),
                                               ("nsfnet", 64, 40.0)])
def test_batched_observation_vs_oracle(topo, n_slots, load):
    """Several envs at different fill levels: germany50/640 (the general kernel: two c2 passes per thread, non-prunable
    paths) and NSFNET with 160 slots and k = 3 (the link-major kernel away from its 320-slot / k = 5 shape: two live
    centre groups at most, scratch sized by the staging area rather than by the start lists)."""
    import torch
    from optical_networking_gym_b200.engine import Engine
    from optical_networking_gym_b200.tracegen import TraceGenerator

    # (100 and 64 slots: the 320-slot NSFNET tables re-dimensioned -- a spectrum that ends inside a bitmap word, and the
    # smallest one, where the link-major kernel's scratch is sized by its record staging area)
    tb = load_tables(topo, n_slots) if (topo, n_slots) != ("nsfnet", 100) and (topo, n_slots) != ("nsfnet", 64) \
        else load_tables("nsfnet", 320).replace(n_slots=n_slots)
    n_envs, n_req = 5, 400
    tr = TraceGenerator(n_envs, tb.n_nodes, tb.n_rates, load, base_seed=321).next(n_req)
    eng = Engine(tb, n_envs, n_req)
    eng.reset(); eng.load_trace_host(*tr)
    obs_dim, n_act = eng.observation_dims()
    dev = torch.device("cuda")
    obs = torch.zeros((n_envs, obs_dim), dtype=torch.float32, device=dev)
    mask = torch.zeros((n_envs, n_act), dtype=torch.uint8, device=dev)
    oracles = []
    for e in range(n_envs):
        o = orc.OracleEnv(tb, n_req)
        o.reset(*[a[:, e] for a in tr])
        oracles.append(o)
    for steps in (0, 150, 249):
        if steps:
            eng.step_first_fit(steps)
            for o in oracles:
                o.run_first_fit(steps, log_qot=False)
        eng.observation(obs, mask)
        torch.cuda.synchronize()
        ho, hm = obs.cpu().numpy(), mask.cpu().numpy()
        for e, o in enumerate(oracles):
            ro, rm = o.observation()
            assert np.abs(ho[e] - ro).max() < OBS_TOL, f"env {e} after {steps} more steps"
            _mask_diff_allowed(tb, o, hm[e], rm)
            assert hm[e, -1] == 1
    eng.close()


def test_fewer_modulations_to_consider_vs_reference_recording():
    """modulations_to_consider = 2 of 6 (examples/ONDM_2025/new_train_multi_ppo.py:101): the observation kernel re-decides
    max_modulation_idx per request (qrmsa.pyx:543-581), the two blocks of a path are modulations max_idx and max_idx - 1
    (:716-719), step() decodes the action with it (:821-829).  Observation, mask, max_modulation_idx, rewards and final
    slots against a recording of the compiled reference."""
    import torch
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import Engine, unpack_bitmaps

    g = load_golden("obs_mc2_nsfnet_320_l260_s5")
    tb = load_tables("nsfnet", 320).replace(mods_to_consider=2)
    n_req, n_act = len(g["src"]), int(g["n_actions"])
    mask_ref = np.unpackbits(g["mask"], axis=1)[:, :n_act]
    eng = Engine(tb, 1, n_req)
    assert eng.observation_dims() == (g["obs"].shape[1], n_act) == (3 + 5 + 12 * 5 * 2, 5 * 2 * 320 + 1)
    eng.reset(); eng.load_trace_host(*[np.ascontiguousarray(g[k][:, None]) for k in TRACE_KEYS])
    dev = torch.device("cuda")
    obs = torch.zeros((1, g["obs"].shape[1]), dtype=torch.float32, device=dev)
    mask = torch.zeros((1, n_act), dtype=torch.uint8, device=dev)
    a = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.zeros(1, dtype=torch.uint8, device=dev)
    rw = torch.zeros(1, dtype=torch.float32, device=dev)
    seen = set()
    for t in range(len(g["action"]) + 1):
        mask.fill_(7)
        eng.observation(obs, mask)
        torch.cuda.synchronize()
        assert int(eng.max_modulation_idx()[0]) == int(g["max_mod"][t]), f"max_modulation_idx at step {t}"
        seen.add(int(g["max_mod"][t]))
        assert np.abs(obs.cpu().numpy()[0] - g["obs"][t]).max() < OBS_TOL, f"obs at step {t}"
        assert np.array_equal(mask.cpu().numpy()[0], mask_ref[t]), f"mask at step {t}"
        if t < len(g["action"]):
            a[0] = int(g["action"][t])
            eng.step_action(a, rw, st, None, None)
            torch.cuda.synchronize()
            assert int(st[0]) in (0, 1) and float(rw[0]) == pytest.approx(float(g["reward"][t]), abs=1e-6)
    assert len(seen) >= 3
    assert np.array_equal(unpack_bitmaps(eng.export_bitmaps(0, 1), 320)[0], g["final_slots"])
    with pytest.raises(_lib.QRMSAError, match="address all modulations"):
        eng.step_first_fit(1)
    eng.close()
