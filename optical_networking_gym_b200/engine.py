"""Thin object wrapper over the C ABI (include/qrmsa_b200.h): one `Engine` = one qrmsa_ctx on one GPU.

PyTorch appears only as plumbing (device buffers, streams); tensors cross the boundary as raw
`data_ptr()` values and the CUDA stream as an integer handle.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import QRMSAError, check
from .tables import StaticTables


def _np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class Engine:
    def __init__(self, tables: StaticTables, n_envs: int, max_requests: int, device: int = 0):
        self.lib = _lib.load()
        self.tables = tables
        self.n_envs = int(n_envs)
        self.max_requests = int(max_requests)
        self.device = int(device)
        self.n_groups = 1
        self.n_loaded = 0
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
            path_length_km=np.ascontiguousarray(tables.path_length_km, np.float64),
            link_length_km=np.ascontiguousarray(tables.link_n_spans * tables.link_span_len_m / 1e3, np.float64),
        )
        t = _lib.StaticTablesC()
        for n in ("n_nodes", "n_links", "k_paths", "n_mods", "mods_to_consider", "n_rates", "n_slots", "max_hops"):
            setattr(t, n, int(getattr(tables, n)))
        for n, a in self._keep.items():
            setattr(t, n, a.ctypes.data)
        t.frequency_start = tables.frequency_start
        t.slot_bandwidth_hz = tables.slot_bandwidth_hz
        t.launch_power_w = tables.launch_power_w
        t.margin_db = tables.margin_db
        h = C.c_void_p()
        rc = self.lib.qrmsa_create(C.byref(t), self.n_envs, self.max_requests, self.device, C.byref(h))
        self._h = h
        if rc != 0:
            try:
                check(rc, h if h else None)
            finally:
                if h:
                    self.lib.qrmsa_destroy(h)
                    self._h = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self.lib.qrmsa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self, stream) -> Optional[int]:
        if stream is None:
            try:
                import torch

                return torch.cuda.current_stream(self.device).cuda_stream or None
            except Exception:
                return None
        return getattr(stream, "cuda_stream", stream) or None

    # ------------------------------------------------------------------ configuration
    def set_groups(self, n_groups: int):
        check(self.lib.qrmsa_set_groups(self._h, int(n_groups)), self._h)
        self.n_groups = int(n_groups)

    def set_staging(self, level: int):
        """0 / 1 / 2: how much per-env state the step kernel stages in shared memory (include/qrmsa_b200.h)."""
        check(self.lib.qrmsa_set_staging(self._h, int(level)), self._h)

    def set_features(self, measure_disruptions: bool = False, defragmentation: bool = False, n_defrag_services: int = 0):
        """The constructor switches of qrmsa.pyx:206-237 (see include/qrmsa_b200.h qrmsa_set_features)."""
        check(self.lib.qrmsa_set_features(self._h, int(bool(measure_disruptions)), int(bool(defragmentation)),
                                          int(n_defrag_services)), self._h)

    def step_disrupted(self, stream=None) -> np.ndarray:
        """Services found disrupted by the last decided request of every env (int32 [n_envs])."""
        out = np.zeros(self.n_envs, np.int32)
        check(self.lib.qrmsa_get_step_disrupted_host(self._h, _np_ptr(out), self._stream(stream)), self._h)
        return out

    def enable_gsnr_log(self, enable: bool = True):
        check(self.lib.qrmsa_enable_gsnr_log(self._h, int(enable)), self._h)

    # ------------------------------------------------------------------ episode control
    def reset(self, stream=None):
        check(self.lib.qrmsa_reset(self._h, self._stream(stream)), self._h)
        self.n_loaded = 0

    def cancel_pending_releases(self, stream=None):
        """The `self._events = []` of reset(options={"only_episode_counters": True}) (qrmsa.pyx:433)."""
        check(self.lib.qrmsa_cancel_pending_releases(self._h, self._stream(stream)), self._h)

    def load_trace_host(self, src, dst, rate, arrival, holding, stream=None):
        """Host arrays shaped [n_requests, n_envs] (request-major)."""
        arrs = [np.ascontiguousarray(src, np.uint8), np.ascontiguousarray(dst, np.uint8),
                np.ascontiguousarray(rate, np.uint8), np.ascontiguousarray(arrival, np.float32),
                np.ascontiguousarray(holding, np.float32)]
        n_req = arrs[0].shape[0]
        for a in arrs:
            if a.shape != (n_req, self.n_envs):
                raise ValueError(f"trace arrays must be [n_requests, n_envs], got {a.shape}")
        check(self.lib.qrmsa_load_trace_host(self._h, *[_np_ptr(a) for a in arrs], n_req, self._stream(stream)), self._h)
        self._sync(stream)  # the host arrays may be pageable: keep them alive until the copies are done
        self.n_loaded = n_req

    def generate_trace(self, n_requests: int, load, seed: int = 0, restart: bool = True, env_offset: int = 0,
                       mean_holding_time: float = 10800.0, node_request_probabilities=None,
                       bit_rate_probabilities=None, stream=None):
        """Draw the next n_requests requests of every env ON THE DEVICE (Philox streams keyed by seed and global env
        index; the draws of qrmsa.pyx:1079-1099, :1134-1148) instead of uploading a host trace."""
        from .tracegen import choice_tables

        ld = np.broadcast_to(np.asarray(load, np.float64), (self.n_envs,)).copy()
        src_cum, dst_cum, rate_cum = choice_tables(self.tables.n_nodes, self.tables.n_rates, node_request_probabilities,
                                                   bit_rate_probabilities)
        check(self.lib.qrmsa_generate_trace(self._h, int(seed), int(bool(restart)), int(env_offset), _np_ptr(ld),
                                            float(mean_holding_time), _np_ptr(src_cum), _np_ptr(dst_cum),
                                            _np_ptr(rate_cum), int(n_requests), self._stream(stream)), self._h)
        self.n_loaded = int(n_requests)

    def trace_host(self, first: int = 0, count: Optional[int] = None, stream=None):
        """The loaded / generated request stream as (src, dst, rate, arrival, holding), each [count, n_envs]."""
        count = self.n_loaded - first if count is None else int(count)
        shape = (count, self.n_envs)
        out = (np.empty(shape, np.uint8), np.empty(shape, np.uint8), np.empty(shape, np.uint8),
               np.empty(shape, np.float32), np.empty(shape, np.float32))
        check(self.lib.qrmsa_get_trace_host(self._h, int(first), count, *[_np_ptr(a) for a in out],
                                            self._stream(stream)), self._h)
        return out

    def load_trace_host_strided(self, ptrs, n_requests: int, row_stride: int, stream=None):
        """Asynchronous upload for a context that owns an env slice of a larger pinned [n_requests, row_stride]
        batch; `ptrs` = the five host addresses of this slice's first env (src, dst, rate, arrival, holding)."""
        check(self.lib.qrmsa_load_trace_host_strided(self._h, *[int(p) for p in ptrs], int(n_requests), int(row_stride),
                                                     self._stream(stream)), self._h)
        self.n_loaded = int(n_requests)

    def actions_host_strided(self, first: int, count: int, out_ptr: int, row_stride: int, stream=None):
        check(self.lib.qrmsa_get_actions_host_strided(self._h, first, count, int(out_ptr), int(row_stride),
                                                      self._stream(stream)), self._h)

    def load_trace_device(self, src, dst, rate, arrival, holding, stream=None):
        """torch CUDA tensors shaped [n_requests, n_envs] (uint8, uint8, uint8, float32, float32)."""
        n_req = src.shape[0]
        for a in (src, dst, rate, arrival, holding):
            if tuple(a.shape) != (n_req, self.n_envs) or not a.is_contiguous() or not a.is_cuda:
                raise ValueError("trace tensors must be contiguous CUDA [n_requests, n_envs]")
        check(self.lib.qrmsa_load_trace(self._h, src.data_ptr(), dst.data_ptr(), rate.data_ptr(), arrival.data_ptr(),
                                        holding.data_ptr(), n_req, self._stream(stream)), self._h)
        self.n_loaded = n_req

    def step_first_fit(self, n_steps: int, stream=None):
        check(self.lib.qrmsa_step_first_fit(self._h, int(n_steps), self._stream(stream)), self._h)

    def step_heuristic(self, policy, n_steps: int, stream=None):
        """policy: "first_fit" | "load_balancing" (or the integer id)."""
        pid = _lib.POLICIES[policy] if isinstance(policy, str) else int(policy)
        check(self.lib.qrmsa_step_heuristic(self._h, pid, int(n_steps), self._stream(stream)), self._h)

    def step_action(self, action, reward=None, status=None, gsnr=None, terminated=None, stream=None):
        """action: int64 CUDA tensor [n_envs]; outputs are optional preallocated CUDA tensors."""
        def p(x):
            return x.data_ptr() if x is not None else None
        check(self.lib.qrmsa_step_action(self._h, action.data_ptr(), p(reward), p(status), p(gsnr), p(terminated),
                                         self._stream(stream)), self._h)

    def observation_dims(self):
        a, b = C.c_int(0), C.c_int(0)
        check(self.lib.qrmsa_observation_dims(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def max_modulation_idx(self, stream=None) -> np.ndarray:
        """Every env's max_modulation_idx as of its last observation (uint8 [n_envs]; qrmsa.pyx:543-581)."""
        out = np.zeros(self.n_envs, np.uint8)
        check(self.lib.qrmsa_get_max_modulation_idx_host(self._h, _np_ptr(out), self._stream(stream)), self._h)
        return out

    def observation(self, obs, mask, stream=None):
        """obs: float32 CUDA [n_envs, obs_dim]; mask: uint8 CUDA [n_envs, n_actions] (both preallocated)."""
        check(self.lib.qrmsa_observation(self._h, obs.data_ptr(), mask.data_ptr(), self._stream(stream)), self._h)
        return obs, mask

    # ------------------------------------------------------------------ results
    def actions_host(self, first: int, count: int, stream=None) -> np.ndarray:
        out = np.empty((count, self.n_envs), np.int32)
        check(self.lib.qrmsa_get_actions_host(self._h, first, count, _np_ptr(out), self._stream(stream)), self._h)
        return out

    def actions_host_into(self, first: int, count: int, out_ptr: int, stream=None):
        check(self.lib.qrmsa_get_actions_host(self._h, first, count, out_ptr, self._stream(stream)), self._h)

    def actions_device(self, first: int, count: int, out, stream=None):
        check(self.lib.qrmsa_get_actions(self._h, first, count, out.data_ptr(), self._stream(stream)), self._h)
        return out

    def env_log(self, env: int, first: int = 0, count: Optional[int] = None):
        """One env's request stream and decision log: (src, dst, rate, arrival, holding, action words) for requests
        [first, first+count)."""
        count = self.n_loaded - first if count is None else int(count)
        rec = np.zeros((count, 4), np.uint32)
        check(self.lib.qrmsa_get_env_log_host(self._h, int(env), int(first), count, _np_ptr(rec)), self._h)
        z = rec[:, 2]
        return ((z & 0xff).astype(np.uint8), ((z >> 8) & 0xff).astype(np.uint8), ((z >> 16) & 0xff).astype(np.uint8),
                rec[:, 0].copy().view(np.float32), rec[:, 1].copy().view(np.float32), rec[:, 3].copy())

    def gsnr_host(self, first: int, count: int, stream=None) -> np.ndarray:
        out = np.empty((count, self.n_envs), np.float64)
        check(self.lib.qrmsa_get_gsnr_host(self._h, first, count, _np_ptr(out), self._stream(stream)), self._h)
        return out

    def ase_nli_host(self, first: int, count: int, stream=None):
        ase = np.empty((count, self.n_envs), np.float64)
        nli = np.empty((count, self.n_envs), np.float64)
        check(self.lib.qrmsa_get_ase_nli_host(self._h, first, count, _np_ptr(ase), _np_ptr(nli), self._stream(stream)), self._h)
        return ase, nli

    def counters(self, stream=None) -> np.ndarray:
        out = np.zeros((self.n_groups, _lib.N_COUNTERS), np.int64)
        check(self.lib.qrmsa_counters(self._h, _np_ptr(out), self._stream(stream)), self._h)
        return out

    def counters_dict(self, group: Optional[int] = None) -> dict:
        c = self.counters()
        row = c.sum(0) if group is None else c[group]
        d = {n: int(row[i]) for i, n in enumerate(_lib.COUNTER_NAMES)}
        d["mod_hist"] = row[16:24].copy()
        d["gn_pruned"] = int(row[24])
        d["disrupted_services"], d["defrag_cycles"], d["service_reallocations"] = int(row[25]), int(row[26]), int(row[27])
        return d

    def counters_tensor(self):
        """Zero-copy torch view (int64 CUDA [n_groups, N_COUNTERS]) of the device counters -- what the episode-end
        NCCL all-reduce takes (clone it first; the kernels keep adding into this memory)."""
        import torch

        class _View:
            pass

        v = _View()
        v.__cuda_array_interface__ = {"data": (self.counters_device_ptr(), False), "shape": (self.n_groups, _lib.N_COUNTERS),
                                      "typestr": "<i8", "version": 2}
        return torch.as_tensor(v, device=torch.device("cuda", self.device))

    def counters_device_ptr(self) -> int:
        p = C.c_void_p()
        check(self.lib.qrmsa_counters_device(self._h, C.byref(p)), self._h)
        return p.value

    def env_state(self, stream=None) -> np.ndarray:
        out = np.zeros((self.n_envs, 4), np.int32)
        check(self.lib.qrmsa_env_state_host(self._h, _np_ptr(out), self._stream(stream)), self._h)
        return out

    def export_slots(self, env: int) -> np.ndarray:
        out = np.zeros((self.tables.n_links, self.tables.n_slots), np.int32)
        check(self.lib.qrmsa_export_slots(self._h, int(env), _np_ptr(out)), self._h)
        return out

    def export_bitmaps(self, first: int, count: int) -> np.ndarray:
        W = (self.tables.n_slots + 31) // 32
        out = np.zeros((count, self.tables.n_links, W), np.uint32)
        check(self.lib.qrmsa_export_bitmaps(self._h, int(first), int(count), _np_ptr(out)), self._h)
        return out

    def export_link_list(self, env: int, link: int) -> np.ndarray:
        cap = self.tables.n_slots
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int(0)
        check(self.lib.qrmsa_export_link_list(self._h, int(env), int(link), _np_ptr(out), cap, C.byref(n)), self._h)
        return out[: n.value].copy()

    def probe_gsnr(self, env: int, src: int, dst: int, p: int, initial_slot: int, number_slots: int) -> float:
        g = C.c_double(0.0)
        check(self.lib.qrmsa_probe_gsnr(self._h, int(env), int(src), int(dst), int(p), int(initial_slot),
                                        int(number_slots), C.byref(g)), self._h)
        return g.value

    def probe_qot(self, env: int, src: int, dst: int, p: int, initial_slot: int, number_slots: int):
        """(GSNR, ASE-only, NLI-only) in dB for a hypothetical channel: the three values of core.osnr.calculate_osnr."""
        v = np.zeros(3, np.float64)
        check(self.lib.qrmsa_probe_qot(self._h, int(env), int(src), int(dst), int(p), int(initial_slot),
                                       int(number_slots), _np_ptr(v)), self._h)
        return float(v[0]), float(v[1]), float(v[2])

    def _sync(self, stream=None):
        import torch

        if stream is None:
            torch.cuda.current_stream(self.device).synchronize()
        else:
            torch.cuda.synchronize(self.device)


def unpack_bitmaps(words: np.ndarray, n_slots: int) -> np.ndarray:
    """uint32 [..., W] -> uint8 [..., S] (1 = free), slot s = bit (s & 31) of word s >> 5."""
    b = np.unpackbits(words.view(np.uint8).reshape(*words.shape[:-1], -1), axis=-1, bitorder="little")
    return b[..., :n_slots]


__all__ = ["Engine", "QRMSAError", "unpack_bitmaps"]
