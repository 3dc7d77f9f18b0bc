"""ctypes binding of oracle/qrmsa_oracle.c.   TEST INFRASTRUCTURE ONLY.

The oracle is the checker for the CUDA path (tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline leg).  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qrmsa_oracle.c")
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "libqrmsa_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # no -ffast-math, no -march=native: must run unchanged on the GPU box's host CPU
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


class _Tables(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_nodes", "n_links", "k_paths", "n_mods", "mods_to_consider", "n_rates", "n_slots", "max_hops")] + [
        ("path_hops", C.c_void_p), ("path_links", C.c_void_p), ("link_n_spans", C.c_void_p),
        ("link_span_len_m", C.c_void_p), ("link_alpha", C.c_void_p), ("link_nf", C.c_void_p),
        ("mod_se", C.c_void_p), ("mod_min_osnr", C.c_void_p), ("bit_rates", C.c_void_p),
        ("slots_needed", C.c_void_p),
        ("frequency_start", C.c_double), ("slot_bw", C.c_double), ("launch_power_w", C.c_double),
        ("margin_db", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.POINTER(_Tables), C.c_int]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_reset.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
        _lib.orc_run_first_fit.restype = C.c_int
        _lib.orc_run_first_fit.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
        _lib.orc_run_heuristic.restype = C.c_int
        _lib.orc_run_heuristic.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
        _lib.orc_step_action.restype = C.c_int
        _lib.orc_step_action.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int), C.c_int]
        _lib.orc_get_slots.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_reset_episode_counters.argtypes = [C.c_void_p]
        _lib.orc_max_mod_idx.restype = C.c_int
        _lib.orc_max_mod_idx.argtypes = [C.c_void_p]
        _lib.orc_set_features.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.orc_get_feature_counters.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_service_start.restype = C.c_int
        _lib.orc_service_start.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_observation.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        _lib.orc_get_link_list.restype = C.c_int
        _lib.orc_get_link_list.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_probe_gsnr.restype = C.c_double
        _lib.orc_probe_gsnr.argtypes = [C.c_void_p] + [C.c_int] * 5
        _lib.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_current_request.restype = C.c_int
        _lib.orc_current_request.argtypes = [C.c_void_p]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def link_length_range(tb):
    """min / max LINK length in km (qrmsa.pyx:676-678), from the span tables: n_spans * span_length."""
    km = tb.link_n_spans * tb.link_span_len_m / 1e3
    return float(km.min()), float(km.max())


COUNTER_NAMES = ("processed", "accepted", "ep_processed", "ep_accepted", "bl_reject", "n_gn", "n_gn_terms",
                 "n_links_read", "n_release", "hops_accepted", "heur_blk_res", "heur_blk_osnr")


class OracleEnv:
    """One scalar FP64 env.  `tables` is any object with the StaticTables attribute names."""

    def __init__(self, tables, max_requests: int = 1):
        L = lib()
        self.tables = tables
        self._keep = dict(
            path_hops=np.ascontiguousarray(tables.path_hops, np.uint8),
            path_links=np.ascontiguousarray(tables.path_links, np.uint8),
            link_n_spans=np.ascontiguousarray(tables.link_n_spans, np.int32),
            link_span_len_m=np.ascontiguousarray(tables.link_span_len_m, np.float64),
            link_alpha=np.ascontiguousarray(tables.link_alpha, np.float64),
            link_nf=np.ascontiguousarray(tables.link_nf, np.float64),
            mod_se=np.ascontiguousarray(tables.mod_se, np.int32),
            mod_min_osnr=np.ascontiguousarray(tables.mod_min_osnr, np.float64),
            bit_rates=np.ascontiguousarray(tables.bit_rates, np.float64),
            slots_needed=np.ascontiguousarray(tables.slots_needed, np.uint8),
        )
        t = _Tables()
        for n in ("n_nodes", "n_links", "k_paths", "n_mods", "mods_to_consider", "n_rates", "n_slots", "max_hops"):
            setattr(t, n, int(getattr(tables, n)))
        for n, a in self._keep.items():
            setattr(t, n, a.ctypes.data)
        t.frequency_start = tables.frequency_start
        t.slot_bw = tables.slot_bandwidth_hz
        t.launch_power_w = tables.launch_power_w
        t.margin_db = tables.margin_db
        self._h = C.c_void_p(L.orc_create(C.byref(t), int(max_requests)))
        self._trace = None
        self.E, self.S = int(tables.n_links), int(tables.n_slots)

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self, src, dst, rate, arrival, holding):
        """Wipe the network and attach a request trace (request 0 becomes current)."""
        tr = (np.ascontiguousarray(src, np.uint8), np.ascontiguousarray(dst, np.uint8),
              np.ascontiguousarray(rate, np.uint8), np.ascontiguousarray(arrival, np.float32),
              np.ascontiguousarray(holding, np.float32))
        self._trace = tr
        lib().orc_reset(self._h, *[_ptr(a) for a in tr], len(tr[0]))

    def run_first_fit(self, n_steps: int, log_qot: bool = True, policy: int = 0):
        action = np.zeros(n_steps, np.int32)
        accepted = np.zeros(n_steps, np.uint8)
        gsnr = np.zeros(n_steps, np.float64)
        per_step = self.tables.k_paths * self.tables.n_mods * (self.tables.n_slots if policy == 2 else 1)   # 2: every start
        cap = n_steps * per_step if log_qot else 0
        qs = np.zeros(max(cap, 1), np.int32)
        qg = np.zeros(max(cap, 1), np.float64)
        qt = np.zeros(max(cap, 1), np.float64)
        qn = C.c_int(0)
        rc = lib().orc_run_heuristic(self._h, int(policy), n_steps, _ptr(action), _ptr(accepted), _ptr(gsnr),
                                     _ptr(qs) if log_qot else None, _ptr(qg) if log_qot else None,
                                     _ptr(qt) if log_qot else None, cap, C.byref(qn))
        if rc != 0:
            raise RuntimeError("oracle: trace exhausted (a step needs the following request)")
        out = dict(action=action, accepted=accepted, gsnr=gsnr)
        if log_qot:
            out.update(qot_step=qs[: qn.value].copy(), qot_gsnr=qg[: qn.value].copy(), qot_thr=qt[: qn.value].copy())
        return out

    def step_action(self, action: int, episode_length: int):
        rw, g, term = C.c_double(0), C.c_double(0), C.c_int(0)
        st = lib().orc_step_action(self._h, int(action), C.byref(rw), C.byref(g), C.byref(term), int(episode_length))
        return st, rw.value, g.value, bool(term.value)

    def set_features(self, measure_disruptions=False, defragmentation=False, n_defrag_services=0):
        """measure_disruptions / defragmentation / n_defrag_services of the constructor (qrmsa.pyx:206-237)."""
        lib().orc_set_features(self._h, int(bool(measure_disruptions)), int(bool(defragmentation)), int(n_defrag_services))

    def feature_counters(self) -> dict:
        out = np.zeros(5, np.int64)
        lib().orc_get_feature_counters(self._h, _ptr(out))
        return dict(disrupted_services=int(out[0]), episode_disrupted_services=int(out[1]), episode_defrag_cicles=int(out[2]),
                    episode_service_realocations=int(out[3]), last_step_disrupted=int(out[4]))

    def service_start(self, request_id: int) -> int:
        return lib().orc_service_start(self._h, int(request_id))

    @property
    def max_modulation_idx(self) -> int:
        return lib().orc_max_mod_idx(self._h)

    def reset_episode_counters(self):
        """reset(options={"only_episode_counters": True}): pending releases dropped, network and request kept."""
        lib().orc_reset_episode_counters(self._h)

    def observation(self):
        """(obs float32[1+2+k+12kM], mask uint8[kMS+1]) for the current request (gen_observation=True mode)."""
        tb = self.tables
        mc = int(tb.mods_to_consider)
        obs = np.zeros(1 + 2 + tb.k_paths + 12 * tb.k_paths * mc, np.float32)
        mask = np.zeros(tb.k_paths * mc * tb.n_slots + 1, np.uint8)
        pl = np.ascontiguousarray(tb.path_length_km, np.float64)
        lo, hi = link_length_range(tb)
        lib().orc_observation(self._h, _ptr(pl), lo, hi, _ptr(obs), _ptr(mask))
        return obs, mask

    def slots(self) -> np.ndarray:
        out = np.zeros((self.E, self.S), np.uint8)
        lib().orc_get_slots(self._h, _ptr(out))
        return out

    def link_list(self, link: int) -> np.ndarray:
        out = np.zeros((self.S, 3), np.int32)
        c = lib().orc_get_link_list(self._h, int(link), _ptr(out))
        return out[:c].copy()

    def probe_gsnr(self, src, dst, p, start, n) -> float:
        return lib().orc_probe_gsnr(self._h, int(src), int(dst), int(p), int(start), int(n))

    def counters(self) -> dict:
        ci = np.zeros(20, np.int64)
        cd = np.zeros(3, np.float64)
        lib().orc_get_counters(self._h, _ptr(ci), _ptr(cd))
        d = {n: int(ci[i]) for i, n in enumerate(COUNTER_NAMES)}
        d["mod_hist"] = ci[12:20].copy()
        d["bit_rate_requested"], d["bit_rate_provisioned"], d["now"] = (float(x) for x in cd)
        return d

    @property
    def current_request(self) -> int:
        return lib().orc_current_request(self._h)


# ------------------------------------------------------------------------------------------------
# Request generation restated with CPython's own `random` (exact by construction): the reference
# draws, per request and in this order (envs/qrmsa.pyx:1079-1089, :1134-1148):
#   expovariate(1/mean_iat), expovariate(1/mean_holding), choices(nodes, w), choices(nodes, w'),
#   choices(bit_rates, probs, k=1)
# ------------------------------------------------------------------------------------------------
def generate_trace_python(n_nodes: int, n_rates: int, load: float, mean_holding: float, seed: int, n_requests: int,
                          start_time: float = 0.0, rng=None, randint_rates=None):
    import random

    rng = rng if rng is not None else random.Random(seed)
    mean_holding = float(np.float32(mean_holding))  # set_load takes a C float (qrmsa.pyx:1124)
    mean_iat = 1 / (load / mean_holding)
    nodes = list(range(n_nodes))
    w = np.full((n_nodes,), fill_value=1.0 / n_nodes, dtype=np.float64)
    probs = [1.0 / n_rates for _ in range(n_rates)]
    src = np.zeros(n_requests, np.uint8)
    dst = np.zeros(n_requests, np.uint8)
    rate = np.zeros(n_requests, np.uint8)
    arrival = np.zeros(n_requests, np.float32)
    holding = np.zeros(n_requests, np.float32)
    now = float(start_time)
    for i in range(n_requests):
        at = np.float32(now + rng.expovariate(1 / mean_iat))
        now = float(at)
        ht = np.float32(rng.expovariate(1.0 / mean_holding))
        s = rng.choices(nodes, weights=w)[0]
        w2 = np.copy(w)
        w2[s] = 0.0
        w2 /= np.sum(w2)
        d = rng.choices(nodes, weights=w2)[0]
        if randint_rates is not None:      # bit_rate_selection="continuous": rng.randint(lower, higher) (qrmsa.pyx:246-254)
            r = rng.randint(int(randint_rates[0]), int(randint_rates[1])) - int(randint_rates[0])
        else:
            r = rng.choices(list(range(n_rates)), probs, k=1)[0]
        src[i], dst[i], rate[i], arrival[i], holding[i] = s, d, r, at, ht
    return dict(src=src, dst=dst, rate=rate, arrival=arrival, holding=holding), rng, now


def philox4x32_10(c, k):
    """Philox4x32-10 (Salmon et al., SC'11; Random123 constants) on numpy uint32 arrays: c = 4 counter words,
    k = 2 key words.  Restates philox4x32_10 of csrc/qrmsa_kernels.cuh for the tests."""
    c = [np.asarray(x, np.uint64) for x in c]
    k = [np.uint64(k[0]), np.uint64(k[1])]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def generate_trace_philox(n_envs: int, n_requests: int, load, seed: int, src_cum, dst_cum, rate_cum,
                          mean_holding: float = 10800.0, env_offset: int = 0, pos0: int = 0):
    """numpy restatement of k_generate_trace (the on-device request generator): the draws of qrmsa.pyx:1079-1099,
    :1134-1148 from Philox streams keyed by (seed, env_offset + env, request index).  Arrays are [n_requests, n_envs]."""
    load = np.broadcast_to(np.asarray(load, np.float64), (n_envs,))
    mh = float(np.float32(mean_holding))
    lam = 1.0 / (1.0 / (load / mh))
    lam_hold = 1.0 / mh
    ge = np.arange(n_envs, dtype=np.uint64) + np.uint64(env_offset)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    N, R = len(src_cum), len(rate_cum)
    src = np.zeros((n_requests, n_envs), np.uint8); dst = np.zeros_like(src); rate = np.zeros_like(src)
    arrival = np.zeros((n_requests, n_envs), np.float32); holding = np.zeros_like(arrival)
    now = np.zeros(n_envs, np.float64)

    def u53(a, b):
        return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)

    for kk in range(n_requests):
        c = np.uint64(pos0 + kk)
        lo, hi = np.full(n_envs, c & np.uint64(0xFFFFFFFF)), np.full(n_envs, c >> np.uint64(32))
        gl, gh = ge & np.uint64(0xFFFFFFFF), (ge >> np.uint64(32)) * np.uint64(2)
        a = philox4x32_10((lo, hi, gl, gh), key)
        b = philox4x32_10((lo, hi, gl, gh + np.uint64(1)), key)
        at = (now + (-np.log(1.0 - u53(a[0], a[1])) / lam)).astype(np.float32)
        now = at.astype(np.float64)
        ht = (-np.log(1.0 - u53(a[2], a[3])) / lam_hold).astype(np.float32)
        s = np.minimum(np.searchsorted(src_cum[: N - 1], (b[0].astype(np.float64) / 4294967296.0) * src_cum[N - 1], side="right"), N - 1)
        xd = (b[1].astype(np.float64) / 4294967296.0) * dst_cum[s, N - 1]
        d = np.array([np.searchsorted(dst_cum[si, : N - 1], x, side="right") for si, x in zip(s, xd)])
        r = np.searchsorted(rate_cum[: R - 1], (b[2].astype(np.float64) / 4294967296.0) * rate_cum[R - 1], side="right")
        src[kk], dst[kk], rate[kk], arrival[kk], holding[kk] = s, d, r, at, ht
    return src, dst, rate, arrival, holding


def sample_masked_actions_numpy(logits, mask, seed: int, step: int):
    """numpy restatement of k_sample_masked (csrc/qrmsa_sampler.cuh): Gumbel-max over the masked logits with the
    Philox4x32-10 stream keyed by (seed, step) and counted by (env, action / 4).  Returns (actions int64 [n_envs],
    keys float32 [n_envs, n_actions] with -inf where masked)."""
    logits = np.asarray(logits, np.float32)
    mask = np.asarray(mask) != 0
    n_envs, n_actions = logits.shape
    seed &= (1 << 64) - 1
    k0 = (seed & 0xFFFFFFFF) ^ (((step * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)) >> 32)
    k1 = ((seed >> 32) & 0xFFFFFFFF) ^ (step & 0xFFFFFFFF)
    n_groups = (n_actions + 3) // 4
    g = np.broadcast_to(np.arange(n_groups, dtype=np.uint64)[None, :], (n_envs, n_groups)).ravel()
    e = np.broadcast_to(np.arange(n_envs, dtype=np.uint64)[:, None], (n_envs, n_groups)).ravel()
    z = np.zeros_like(g)
    r = philox4x32_10((g, z, e, z), (k0, k1))
    bits = np.stack([x.astype(np.uint32) for x in r], axis=1).reshape(n_envs, n_groups * 4)[:, :n_actions]
    u = ((bits >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    gum = -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)
    keys = np.where(mask, logits + gum, -np.inf).astype(np.float32)
    act = keys.argmax(axis=1).astype(np.int64)
    act[~mask.any(axis=1)] = n_actions - 1
    return act, keys
