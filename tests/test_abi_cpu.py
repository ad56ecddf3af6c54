"""The C-ABI library loads on a CPU-only box and exports every symbol include/acm_b200.h
declares (no compute calls without a GPU)."""
import os
import re

from helpers import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "acm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from acm_gnn_b200 import _lib
    from acm_gnn_b200 import build as B
    B.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in acm_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, "ctypes prototypes out of sync with the header"
    assert lib.acm_version() >= 100
    assert isinstance(lib.acm_last_error_string(), bytes)


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any launch, with an error string."""
    from acm_gnn_b200 import _lib
    lib = _lib.load()
    rc = lib.acm_spmm_mix_fwd(7, 256, 256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 3, 0, 0, 3.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    assert rc == 10001
    assert b"dtype" in lib.acm_last_error_string()
    rc = lib.acm_cast_pad(0, 1, 1, 1, 0, 1, 4, 0)
    assert rc == 10001
