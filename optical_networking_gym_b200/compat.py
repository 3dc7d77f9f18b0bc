"""Running the reference's OWN, unchanged heuristics against the B200 `QRMSAEnv`.

The reference's `optical_networking_gym.heuristics.heuristics` module is tied to its Cython env by two module-level
names only (reference heuristics/heuristics.py:8-10, :15-33):

* `get_qrmsa_env(env)` unwraps `.env` chains until it meets an instance of the *Cython* `QRMSAEnv` class and raises
  otherwise;
* `calculate_osnr` is imported from `core.osnr`, which walks `topology[u][v]["running_services"]` -- Python objects
  the B200 env does not keep (its channel lists live in device memory).

Everything else the heuristics touch is the attribute / method surface `env.QRMSAEnv` mirrors (`current_service`,
`k_shortest_paths`, `modulations`, `get_available_slots`, `_get_candidates`, `get_number_slots`, `is_path_free`, ...).
`patch_reference_heuristics()` rebinds those two names to dispatchers: a B200 env is recognised and served by the
device (`qrmsa_probe_qot`), a reference env goes to the original functions -- so the same heuristic function can be
called on both implementations side by side (tests/test_gpu_reference_heuristics.py).
"""
from __future__ import annotations


def patch_reference_heuristics(module=None):
    """Rebind `get_qrmsa_env` and `calculate_osnr` in the reference's heuristics module (imported from
    `optical_networking_gym.heuristics.heuristics` when not given).  Idempotent; returns the module."""
    from .env import QRMSAEnv
    from .env import calculate_osnr as b200_calculate_osnr

    if module is None:
        import optical_networking_gym.heuristics.heuristics as module
    if getattr(module, "_b200_patched", False):
        return module
    ref_get, ref_osnr = module.get_qrmsa_env, module.calculate_osnr

    def get_qrmsa_env(env):
        e = env
        while True:
            if isinstance(e, QRMSAEnv):
                return e
            if hasattr(e, "env"):
                e = e.env
                continue
            return ref_get(env)

    def calculate_osnr(env, service, *args, **kwargs):
        if isinstance(env, QRMSAEnv):
            return b200_calculate_osnr(env, service)
        return ref_osnr(env, service, *args, **kwargs)

    module.get_qrmsa_env = get_qrmsa_env
    module.calculate_osnr = calculate_osnr
    module._b200_patched = True
    module._b200_originals = (ref_get, ref_osnr)
    return module


def unpatch_reference_heuristics(module=None):
    if module is None:
        import optical_networking_gym.heuristics.heuristics as module
    if getattr(module, "_b200_patched", False):
        module.get_qrmsa_env, module.calculate_osnr = module._b200_originals
        module._b200_patched = False
    return module
