"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

gymnasium is not installed in this image and there is no network.  The compiled
reference (oracle/_ref) imports it at module load (reference
optical_networking_gym/envs/qrmsa.pyx:10-11, wrappers/qrmsa_gym.py:4-6,
heuristics/heuristics.py:10).  Only the handful of names the reference touches
are provided; nothing here is on the product path.
"""
from . import spaces, utils, envs  # noqa: F401
from .core import Env, Wrapper  # noqa: F401


def make(id, **kwargs):  # pragma: no cover - convenience only
    from .envs.registration import registry
    import importlib

    mod, cls = registry[id]["entry_point"].split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
