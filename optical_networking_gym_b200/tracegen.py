"""Request-stream generation for a batch of envs (host side, C++ in csrc/tracegen.cpp).

Env i draws from the stream `random.Random(base_seed + i)` exactly as the reference's
QRMSAEnv._next_service / _get_node_pair do (reference envs/qrmsa.pyx:1079-1089, :1134-1148), so
"replay a request trace the reference generated from the same seeds" needs no reference at run time.

The cumulative-weight tables `random.choices` bisects are built HERE with the same numpy/itertools
expressions the reference uses (np.full(1/N), copy, zero the source, `/= np.sum`, accumulate), so the
float64 values are bit-identical without restating numpy's pairwise summation.
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check


def choice_tables(n_nodes: int, n_rates: int, node_request_probabilities=None, bit_rate_probabilities=None):
    if node_request_probabilities is None:
        w = np.full((n_nodes,), fill_value=1.0 / n_nodes, dtype=np.float64)  # qrmsa.pyx:279-284
    else:
        w = np.asarray(node_request_probabilities, dtype=np.float64)
    src_cum = np.array(list(itertools.accumulate(w)), np.float64)
    dst_cum = np.zeros((n_nodes, n_nodes), np.float64)
    for s in range(n_nodes):
        w2 = np.copy(w)                      # qrmsa.pyx:1140-1143
        w2[s] = 0.0
        w2 /= np.sum(w2)
        dst_cum[s] = list(itertools.accumulate(w2))
    if bit_rate_probabilities is None:
        bit_rate_probabilities = [1.0 / n_rates for _ in range(n_rates)]  # qrmsa.pyx:256-257
    rate_cum = np.array(list(itertools.accumulate(bit_rate_probabilities)), np.float64)
    return src_cum, dst_cum, rate_cum


class TraceGenerator:
    """Streams of (src, dst, rate index, arrival f32, holding f32); arrays are [n_requests, n_envs]."""

    def __init__(self, n_envs: int, n_nodes: int, n_rates: int, load, mean_holding_time: float = 10800.0,
                 base_seed: int = 50, node_request_probabilities=None, bit_rate_probabilities=None,
                 n_threads: int = 0, randint_rates=None):
        """randint_rates=(lower, higher): bit_rate_selection="continuous" -- the rate index drawn is
        rng.randint(lower, higher) - lower (qrmsa.pyx:246-254) instead of a choice among n_rates."""
        self.lib = _lib.load()
        self.n_envs = int(n_envs)
        self.n_threads = int(n_threads)
        load = np.broadcast_to(np.asarray(load, np.float64), (self.n_envs,)).copy()
        src_cum, dst_cum, rate_cum = choice_tables(n_nodes, n_rates, node_request_probabilities, bit_rate_probabilities)
        h = C.c_void_p()
        check(self.lib.qrmsa_tracegen_create(self.n_envs, int(base_seed), int(n_nodes), int(n_rates),
                                             load.ctypes.data, float(mean_holding_time), src_cum.ctypes.data,
                                             dst_cum.ctypes.data, rate_cum.ctypes.data, C.byref(h)))
        self._h = h
        if randint_rates is not None:
            check(self.lib.qrmsa_tracegen_set_randint_rates(self._h, int(randint_rates[0]), int(randint_rates[1])))

    def next(self, n_requests: int, out: Optional[Sequence[np.ndarray]] = None):
        """Next n_requests requests of every env.  `out` may hold 5 preallocated (e.g. pinned) arrays."""
        shape = (int(n_requests), self.n_envs)
        if out is None:
            out = (np.empty(shape, np.uint8), np.empty(shape, np.uint8), np.empty(shape, np.uint8),
                   np.empty(shape, np.float32), np.empty(shape, np.float32))
        src, dst, rate, arrival, holding = out
        for a in out:
            assert a.shape == shape and a.flags.c_contiguous
        check(self.lib.qrmsa_tracegen_next(self._h, int(n_requests), src.ctypes.data, dst.ctypes.data,
                                           rate.ctypes.data, arrival.ctypes.data, holding.ctypes.data, self.n_threads))
        return src, dst, rate, arrival, holding

    def close(self):
        if getattr(self, "_h", None):
            self.lib.qrmsa_tracegen_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
