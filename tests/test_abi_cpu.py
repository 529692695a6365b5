"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/scrubby_gpu.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

from scrubby_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "scrubby_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported():
    L = _lib.load()
    declared = _declared()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/scrubby_gpu.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert L.sgpu_abi_version() == 2


def test_strerror_covers_every_status():
    L = _lib.load()
    for code, name in _lib.STATUS.items():
        msg = L.sgpu_strerror(code)
        assert msg and msg != b"unknown status", name
    hdr = open(os.path.join(ROOT, "include", "scrubby_gpu.h")).read()
    for code, name in _lib.STATUS.items():
        assert re.search(rf"{name}\s*=\s*{code}\b", hdr), name


def test_no_cpu_fallback_without_device():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.load()
    h = C.c_void_p()
    assert L.sgpu_ctx_create(0, C.byref(h)) == 16  # SGPU_ERR_CUDA, never a silent CPU path
    from scrubby_b200 import api

    with pytest.raises(api.ScrubbyGpuError):
        api.Context(0)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under scrubby_b200/ or include/ may reference it"""
    bad = []
    for base in ("scrubby_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|liboracle|scrubby_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
