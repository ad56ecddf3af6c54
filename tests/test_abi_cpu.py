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


def test_ctypes_prototypes_match_the_header_argument_by_argument():
    """Every ctypes prototype has the arity and the argument classes (pointer / 64-bit integer / int / float) of the
    declaration in include/acm_b200.h -- a stale prototype would otherwise only show up as garbage on the GPU."""
    import ctypes
    from acm_gnn_b200 import _lib
    src = open(os.path.join(ROOT, "include", "acm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = dict(re.findall(r"\b(acm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src))
    assert set(decls) == set(_lib._PROTOS)

    def cls(param):
        param = param.strip()
        if "*" in param:
            return ctypes.c_void_p
        if re.match(r"(const\s+)?(int64_t|uint64_t)\b", param):
            return ctypes.c_int64
        if re.match(r"(const\s+)?float\b", param):
            return ctypes.c_float
        assert re.match(r"(const\s+)?(int|int32_t)\b", param), param
        return ctypes.c_int

    for name, params in decls.items():
        want = [] if params.strip() in ("", "void") else [cls(p) for p in params.split(",")]
        assert _lib._PROTOS[name] == want, f"{name}: header {[w.__name__ for w in want]} vs ctypes {[a.__name__ for a in _lib._PROTOS[name]]}"
