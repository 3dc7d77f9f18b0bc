"""B200-native batched QRMSA environment step (drop-in for the reference's Cython path).

Layers (reference file:line in each module's docstring):
  csrc/            sm_100a CUDA kernels + the C ABI (include/qrmsa_b200.h)
  _lib.py          in-tree nvcc build + ctypes binding
  engine.py        one context per GPU: reset / load_trace / step_first_fit / step_action / counters
  tables.py        dense static tables exported from a reference `topology` graph
  tracegen.py      CPython-`random`-exact request streams for a batch of envs
  env.py           QRMSAEnv-compatible API (reset / step / action masks, heuristics' call surface)

There is no CPU fallback: without the CUDA library and a GPU every compute call raises.
"""
from .tables import StaticTables  # noqa: F401
from ._lib import QRMSAError  # noqa: F401

__all__ = ["StaticTables", "QRMSAError"]
__version__ = "0.1.0"
