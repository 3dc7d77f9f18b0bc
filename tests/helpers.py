"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

from optical_networking_gym_b200.tables import StaticTables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TRACE_KEYS = ("src", "dst", "rate", "arrival", "holding")


def load_tables(topo: str, n_slots: int) -> StaticTables:
    return StaticTables.load(os.path.join(GOLDEN, f"tables_{topo}_{n_slots}.npz"))


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def parse_tag(tag: str):
    """'run_nobel-eu_320_l300_s50' -> ('nobel-eu', 320)"""
    parts = tag.split("_")
    return parts[1], int(parts[2])


def first_divergence(a, b):
    d = np.flatnonzero(np.asarray(a) != np.asarray(b))
    return int(d[0]) if len(d) else -1


def compare_decisions(actions, ref_actions, flagged, what=""):
    """Bit-exact comparison with the north_star rule: a decision whose GSNR lies within 1e-3 dB of a
    threshold is FLAGGED; divergence at or after a flagged step of that env is reported, not counted.
    Returns (n_compared, n_excused)."""
    actions = np.asarray(actions); ref_actions = np.asarray(ref_actions); flagged = np.asarray(flagged, bool)
    assert actions.shape == ref_actions.shape
    if actions.ndim == 1:
        actions, ref_actions, flagged = actions[None], ref_actions[None], flagged[None]
    n_cmp = n_exc = 0
    for e in range(actions.shape[0]):
        d = first_divergence(actions[e], ref_actions[e])
        if d < 0:
            n_cmp += actions.shape[1]
            continue
        fl = np.flatnonzero(flagged[e, : d + 1])
        assert len(fl) > 0, (f"{what} env {e}: decision mismatch at step {d} "
                             f"(got {actions[e, d]}, reference {ref_actions[e, d]}) with no near-threshold flag before it")
        n_cmp += d
        n_exc += 1
    return n_cmp, n_exc
