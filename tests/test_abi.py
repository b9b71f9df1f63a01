"""The C-ABI library loads on a CPU-only box and exports every symbol hanselx.h declares."""
import ctypes
import os
import re

from gretel_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hanselx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    build.build_lib()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == _declared()


def test_loads_and_reports_errors_without_gpu():
    lib = _lib.load()
    assert lib.hx_version() >= 100
    # argument errors are reported without touching CUDA
    assert lib.hx_create(-1, 1, 0, ctypes.byref(ctypes.c_void_p())) == _lib.HX_E_ARG
    assert b"argument check failed" in lib.hx_last_error()


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle."""
    pkg = os.path.join(ROOT, "gretel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "libhansel_oracle", "c_oracle", '#include "../oracle',
                            '#include "oracle'):
                    assert bad not in txt, (f, bad)
