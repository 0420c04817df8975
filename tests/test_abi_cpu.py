"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/combblas_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import pytest

import cbb200_loader

cb = cbb200_loader.load_package()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "combblas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = cb.capi.lib()
    declared = header_functions()
    assert len(declared) >= 30
    missing = [f for f in declared if not hasattr(L, f)]
    assert not missing, missing
    assert sorted(cb.capi.SYMBOLS) == declared          # the binding covers the whole header
    assert L.cb_abi_version() == 1


def test_status_strings_and_semiring_identities():
    import ctypes
    import numpy as np
    L = cb.capi.lib()
    assert L.cb_status_string(3002) == b"DIMMISMATCH" and L.cb_status_string(3005) == b"MATRIXALIAS"
    for sr, dt, want in [(cb.PLUS_TIMES, cb.F64, 0.0), (cb.MIN_PLUS, cb.I32, 2**31 - 1), (cb.MIN_PLUS, cb.F32, np.finfo(np.float32).max),
                         (cb.MAX_SEL2ND, cb.I64, -1), (cb.OR_AND, cb.U8, 0)]:
        out = np.zeros(1, cb.capi.NP_OF[dt])
        assert L.cb_semiring_id(sr, dt, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert out[0] == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CBError) as e:
        cb.Context(0)
    assert e.value.status == 1            # CB_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "combblas-spmm-test_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(d, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "oracle/", "oracle.", "liboracle", "libcbref", "cbref_", "oracle_spmm"):
                    assert needle not in text, f"{os.path.join(d, f)} references the checker ({needle})"
