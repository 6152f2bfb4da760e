"""The C-ABI library loads and exports every symbol include/colorid_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import colorid_b200.lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "colorid_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cid_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_python_binding_covers_header():
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = L.load()
    h = ctypes.c_void_p()
    rc = lib.cid_ctx_create(0, ctypes.byref(h))
    assert rc == L.CID_E_CUDA
    assert b"no CPU fallback" in lib.cid_last_error()
