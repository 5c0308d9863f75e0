"""CPU: the C-ABI library loads, exports every symbol include/b200reco.h declares, and its
host-side entry points work (no device compute here)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200reco.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from librecommender_b200 import _lib

    syms = _declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in b200reco.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) <= set(syms)
    assert _lib.lib.b200_version() == 100


def test_consumed_csr_host_matches_rust_semantics():
    from librecommender_b200 import ConsumedCSR
    from oracle.ranking import build_consumed_unique

    # reference tests/test_consumed.py:12-25
    u = [1, 1, 1, 2, 2, 1, 2, 3, 2, 3]
    i = [11, 11, 999, 0, 11, 11, 999, 11, 999, 0]
    csr = ConsumedCSR.from_interactions(u, i, n_users=4)
    assert csr.row(0).tolist() == []
    assert csr.row(1).tolist() == [11, 999, 11]
    assert csr.row(2).tolist() == [0, 11, 999]
    assert csr.row(3).tolist() == [11, 0]
    rng = np.random.default_rng(3)
    uu = rng.integers(0, 50, size=5000)
    ii = rng.integers(0, 7, size=5000)  # few items => many consecutive repeats
    csr = ConsumedCSR.from_interactions(uu, ii, n_users=50)
    uc, _ = build_consumed_unique(uu, ii)
    assert csr.to_dict() == uc
    assert ConsumedCSR.from_dict(uc, 50).to_dict() == uc


def test_error_reporting_without_gpu():
    from librecommender_b200 import _lib

    n = ctypes.c_size_t(0)
    rc = _lib.lib.b200_topk_rows_workspace_bytes(4, 0, 1, ctypes.byref(n))
    assert rc != 0
    assert b"bad shape" in _lib.lib.b200_last_error()
    rc = _lib.lib.b200_topk_rows_workspace_bytes(4, 1000, 10, ctypes.byref(n))
    assert rc == 0 and n.value > 0
