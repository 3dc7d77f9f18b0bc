"""ctypes binding of libqrmsa_b200.so (include/qrmsa_b200.h) and its in-tree nvcc build.

The library is built IN-TREE (optical_networking_gym_b200/libqrmsa_b200.so) so that it travels with
the repository snapshot to the GPU box.  There is no CPU fallback: every compute entry point needs a
CUDA device and raises QRMSAError otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libqrmsa_b200.so")
SOURCES = ("qrmsa_b200.cu", "tracegen.cpp")
HEADERS = ("qrmsa_kernels.cuh", "qrmsa_sampler.cuh", os.path.join("..", "..", "include", "qrmsa_b200.h"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

N_COUNTERS = 32
COUNTER_NAMES = ("decided", "accepted", "rejected", "rate_requested_milli", "rate_provisioned_milli",
                 "hops_accepted", "links_read", "records_read", "gn_terms", "gn_evals", "releases",
                 "near_threshold", "blocked_resources", "blocked_osnr", "paths_tried", "errors")
ACTION_MASK = 0x00FFFFFF
FLAG_NEAR_THRESHOLD = 0x80000000
FLAG_DECIDED = 0x40000000
FLAG_ACCEPTED = 0x20000000
FLAG_BLOCKED_RESOURCES = 0x01000000   # rejected requests: the heuristic's blocked_due_to_resources
FLAG_BLOCKED_OSNR = 0x02000000        # rejected requests: blocked_due_to_osnr
FLAG_DISRUPTED = 0x10000000            # the service's GSNR fell below its modulation's minimum_osnr (measure_disruptions)
FLAG_RELEASE_CANCELLED = 0x08000000   # accepted, release event dropped by reset(options={"only_episode_counters": True})
FLAG_NEAR_TIE = 0x04000000            # highest-SNR policy: runner-up within 1e-6 dB of the chosen candidate
POLICY_FIRST_FIT, POLICY_LOAD_BALANCING, POLICY_HIGHEST_SNR, POLICY_LB_FIRST_FIT = 0, 1, 2, 3
POLICIES = {"first_fit": 0, "load_balancing": 1, "highest_snr": 2, "load_balancing_first_fit": 3}
STEP_ACCEPTED, STEP_REJECT_ACTION, STEP_NOT_FREE, STEP_LOW_GSNR, STEP_IDLE = range(5)


class QRMSAError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise QRMSAError("nvcc not found: the CUDA library cannot be built (no CPU fallback exists)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libqrmsa_b200.so (cross-compiles without a GPU)."""
    if force or needs_build():
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + list(SOURCES)
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise QRMSAError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    return LIB_PATH


class StaticTablesC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_nodes", "n_links", "k_paths", "n_mods", "mods_to_consider", "n_rates", "n_slots", "max_hops")] + [
        ("path_hops", C.c_void_p), ("path_links", C.c_void_p), ("link_n_spans", C.c_void_p),
        ("link_span_len_m", C.c_void_p), ("link_alpha", C.c_void_p), ("link_nf", C.c_void_p),
        ("mod_se", C.c_void_p), ("mod_min_osnr", C.c_void_p), ("bit_rates", C.c_void_p),
        ("slots_needed", C.c_void_p),
        ("frequency_start", C.c_double), ("slot_bandwidth_hz", C.c_double), ("launch_power_w", C.c_double),
        ("margin_db", C.c_double), ("path_length_km", C.c_void_p), ("link_length_km", C.c_void_p)]


# name -> (restype, argtypes); the symbol list is also what tests/test_abi.py checks against the header
_P, _I, _U64 = C.c_void_p, C.c_int, C.c_uint64
SIGNATURES = {
    "qrmsa_version": (C.c_char_p, []),
    "qrmsa_strerror": (C.c_char_p, [_I]),
    "qrmsa_last_error": (C.c_char_p, [_P]),
    "qrmsa_create": (_I, [C.POINTER(StaticTablesC), _I, _I, _I, C.POINTER(_P)]),
    "qrmsa_destroy": (None, [_P]),
    "qrmsa_set_groups": (_I, [_P, _I]),
    "qrmsa_set_staging": (_I, [_P, _I]),
    "qrmsa_set_features": (_I, [_P, _I, _I, _I]),
    "qrmsa_get_step_disrupted_host": (_I, [_P, _P, _P]),
    "qrmsa_enable_gsnr_log": (_I, [_P, _I]),
    "qrmsa_reset": (_I, [_P, _P]),
    "qrmsa_cancel_pending_releases": (_I, [_P, _P]),
    "qrmsa_load_trace": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "qrmsa_load_trace_host": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "qrmsa_load_trace_host_strided": (_I, [_P, _P, _P, _P, _P, _P, _I, C.c_int64, _P]),
    "qrmsa_generate_trace": (_I, [_P, _U64, _I, C.c_int64, _P, C.c_double, _P, _P, _P, _I, _P]),
    "qrmsa_get_trace_host": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "qrmsa_step_first_fit": (_I, [_P, _I, _P]),
    "qrmsa_step_heuristic": (_I, [_P, _I, _I, _P]),
    "qrmsa_step_action": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "qrmsa_observation": (_I, [_P, _P, _P, _P]),
    "qrmsa_observation_dims": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "qrmsa_get_max_modulation_idx_host": (_I, [_P, _P, _P]),
    "qrmsa_sample_masked_actions": (_I, [_P, _I, _P, _I, _I, C.c_int64, C.c_int64, _U64, _U64, _P, _I, _P]),
    "qrmsa_get_actions": (_I, [_P, _I, _I, _P, _P]),
    "qrmsa_get_actions_host": (_I, [_P, _I, _I, _P, _P]),
    "qrmsa_get_actions_host_strided": (_I, [_P, _I, _I, _P, C.c_int64, _P]),
    "qrmsa_get_env_log_host": (_I, [_P, _I, _I, _I, _P]),
    "qrmsa_get_gsnr_host": (_I, [_P, _I, _I, _P, _P]),
    "qrmsa_get_ase_nli_host": (_I, [_P, _I, _I, _P, _P, _P]),
    "qrmsa_counters": (_I, [_P, _P, _P]),
    "qrmsa_counters_device": (_I, [_P, C.POINTER(_P)]),
    "qrmsa_env_state_host": (_I, [_P, _P, _P]),
    "qrmsa_export_slots": (_I, [_P, _I, _P]),
    "qrmsa_export_bitmaps": (_I, [_P, _I, _I, _P]),
    "qrmsa_export_link_list": (_I, [_P, _I, _I, _P, _I, C.POINTER(_I)]),
    "qrmsa_probe_gsnr": (_I, [_P, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_double)]),
    "qrmsa_probe_qot": (_I, [_P, _I, _I, _I, _I, _I, _I, _P]),
    "qrmsa_tracegen_create": (_I, [_I, _U64, _I, _I, _P, C.c_double, _P, _P, _P, C.POINTER(_P)]),
    "qrmsa_tracegen_set_randint_rates": (_I, [_P, _I, _I]),
    "qrmsa_tracegen_next": (_I, [_P, _I, _P, _P, _P, _P, _P, _I]),
    "qrmsa_tracegen_destroy": (None, [_P]),
}

_lib = None


def load():
    """Load the in-tree library (building it first if sources are newer).  Fails loudly if it cannot."""
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, ctx=None) -> None:
    if rc != 0:
        lib = load()
        msg = lib.qrmsa_strerror(rc).decode()
        if ctx:
            detail = lib.qrmsa_last_error(ctx).decode()
            if detail:
                msg += ": " + detail
        raise QRMSAError(msg)
