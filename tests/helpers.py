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
    """Bit-exact comparison with the north_star rule: a decision whose GSNR lies within 1e-3 dB of a threshold is
    FLAGGED rather than counted as a mismatch.  Up to its first divergent step an env's decisions -- hence its network
    state -- are identical to the reference's, so a legitimate divergence can only come from a QoT comparison made AT
    that step: the first divergent step itself must carry the flag (an earlier flag excuses nothing).  What follows a
    flagged divergence in that env is reported, not counted.  Returns (n_compared, n_excused)."""
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
        assert flagged[e, d], (f"{what} env {e}: decision mismatch at step {d} (got {actions[e, d]}, reference "
                               f"{ref_actions[e, d]}) and that step carries no near-threshold flag")
        n_cmp += d
        n_exc += 1
    if n_exc:
        print(f"{what}: {n_exc} env(s) excused after a flagged near-threshold divergence ({n_cmp} decisions compared)")
    return n_cmp, n_exc


def check_flags_against_recording(flagged_steps, g, tol_db=1e-3, what=""):
    """Both directions of the flag rule against a recording that holds every QoT check of the reference
    (qot_step / qot_gsnr / qot_thr): every step with a recorded check strictly inside the tolerance is flagged, and
    every flagged step has a recorded check within the tolerance (1e-9 dB of slack for the two FP64 evaluations)."""
    delta = np.abs(np.asarray(g["qot_gsnr"]) - np.asarray(g["qot_thr"]))
    steps = np.asarray(g["qot_step"])
    must = set(np.unique(steps[delta < tol_db - 1e-9]).tolist())
    may = set(np.unique(steps[delta < tol_db + 1e-9]).tolist())
    got = set(int(x) for x in np.flatnonzero(np.asarray(flagged_steps, bool)))
    assert must <= got, f"{what}: near-threshold checks at steps {sorted(must - got)} were not flagged"
    assert got <= may, f"{what}: steps {sorted(got - may)} are flagged without a QoT check within {tol_db} dB"
    return len(got)
