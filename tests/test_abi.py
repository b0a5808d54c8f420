"""The C-ABI library loads and exports every symbol include/kv_b200.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

import pytest
import torch

from knightvision_b200 import _native as N
from knightvision_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    B.build_native()
    return N.lib()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "kv_b200.h")).read()
    return sorted(set(re.findall(r"KV_API\s+[\w\s\*]+?\b(kv_\w+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/kv_b200.h but not exported"
        assert s in N.SIGNATURES, f"{s} has no ctypes signature in knightvision_b200/_native.py"
    assert lib.kv_abi_version() >= 1


def test_no_cpu_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = lib.kv_create(0, ctypes.byref(ctx))
    assert rc != 0 and not ctx.value
    assert b"no CPU fallback" in lib.kv_last_error(None)
    from knightvision_b200.engine import Engine
    with pytest.raises(Exception):
        Engine(0)


def test_product_never_imports_oracle():
    """The oracle and the emulator are checkers: nothing under knightvision_b200/ includes, imports or loads them
    (comments may cite them)."""
    pkg = os.path.join(ROOT, "knightvision_b200")
    bad = re.compile(r'#\s*include\s*[<"][^">]*(oracle|simt_emu)|^\s*(from|import)\s+(oracle|simt_emu|tests)\b|'
                     r'libkv_oracle|libkv_emu|kvemu_|kvo_', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                m = bad.search(src)
                assert m is None or (f == "kv_warp.cuh" and "kvemu" in m.group(0)), (f, m.group(0))
