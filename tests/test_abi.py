"""The C-ABI library loads on a CPU-only box and exports every symbol include/qrmsa_b200.h declares;
compute entry points fail loudly without a device (there is no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, load_tables
from optical_networking_gym_b200 import _lib


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "qrmsa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qrmsa_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding table and header disagree"


def test_version_and_strerror():
    lib = _lib.load()
    assert b"sm_100a" in lib.qrmsa_version()
    assert lib.qrmsa_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.qrmsa_strerror(5)


def test_sass_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from optical_networking_gym_b200.engine import Engine, QRMSAError

    with pytest.raises(QRMSAError, match="no CUDA device"):
        Engine(load_tables("ring4", 320), 4, 16)


def test_counter_enum_matches_python_names():
    text = open(os.path.join(ROOT, "include", "qrmsa_b200.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"(QRMSA_CNT_[A-Z_]+)\s*=\s*(\d+)", text))
    assert enum["QRMSA_CNT_DECIDED"] == 0 and enum["QRMSA_CNT_MOD_HIST"] == 16 and enum["QRMSA_CNT_GN_PRUNED"] == 24
    assert int(re.search(r"QRMSA_N_COUNTERS\s*=\s*(\d+)", text).group(1)) == _lib.N_COUNTERS
    assert len(_lib.COUNTER_NAMES) == 16
