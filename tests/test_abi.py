"""The C-ABI library loads on a CPU-only box, exports every symbol include/adaflo_b200.h
declares, and refuses to compute without a CUDA device (no fallback)."""
import ctypes as C
import os
import re

import pytest

from adaflo_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "adaflo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adaflo_b200_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = header_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the python binding table covers the same set
    assert sorted(capi.SYMBOLS) == names


def test_version_and_error_strings():
    lib = capi.load()
    assert b"adaflo_b200" in lib.adaflo_b200_version()
    assert isinstance(lib.adaflo_b200_last_error(), bytes)


def test_struct_sizes_match_the_header_layout():
    # POD layouts are part of the ABI
    assert C.sizeof(capi.FlowParams) == 4 * 4 + 7 * 8
    assert C.sizeof(capi.TimeWeights) == 6 * 8
    assert C.sizeof(capi.Exchange) == 24
    assert C.sizeof(capi.QuadDesc) == 16 + 4 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    lib = capi.load()
    h = C.c_void_p()
    rc = lib.adaflo_b200_ctx_create(C.byref(h), 0, None)
    assert rc == 2  # ADAFLO_B200_ERR_CUDA
    assert b"no CUDA device" in lib.adaflo_b200_last_error()
    from adaflo_b200.operators import Context
    with pytest.raises(capi.AdafloB200Error):
        Context(0)


def test_product_code_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "adaflo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liboracle", r"libcpu_port", r"oracle/", r"#include\s+\"[^\"]*oracle"):
                    assert not re.search(pat, src, flags=re.M), f"{f} uses the oracle ({pat})"
