"""ctypes binding of libacm_b200.so (the C ABI declared in include/acm_b200.h).

No torch types cross this boundary: raw device pointers, sizes and the CUDA stream handle.
The product path fails loudly when the library is missing -- there is no CPU fallback.
"""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libacm_b200.so")

ACM_F32, ACM_BF16 = 0, 1
GEMM_SIMT, GEMM_TCGEN05 = 0, 1

_c = ctypes
_vp, _i64, _i32, _f32 = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_float

# name -> argtypes  (restype int unless listed in _RESTYPES); mirrors include/acm_b200.h
_PROTOS = {
    "acm_version": [],
    "acm_last_error_string": [],
    "acm_launch_count": [],
    "acm_csr_rowptr": [_vp, _i64, _i64, _vp, _vp],
    "acm_degree_normalise": [_vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "acm_csr_transpose_values": [_vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "acm_pack_params": [_i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "acm_unpack_grads": [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "acm_cast_pad": [_vp, _i64, _i64, _i64, _vp, _i32, _i64, _vp],
    "acm_gemm_xw_fwd": [_i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp],
    "acm_gemm_bwd_dw": [_i32, _i32, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp],
    "acm_gemm_bwd_dx": [_i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp],
    "acm_spmm_mix_fwd": [_i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                         _i32, _i32, _i32, _f32, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "acm_spmm_long_rows": [_i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "acm_mix_bwd": [_i32, _i32, _i32, _i64, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                    _i32, _i32, _i32, _f32, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp],
    "acm_spmm_t_bwd_rank1": [_i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp],
    "acm_gemm_xw_fwd_push": [_vp, _i64, _vp, _vp, _i32, _i64, _vp, _vp, _i64, _i64, _i64, _i32, _vp],
    "acm_spmm_t_bwd": [_i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "acm_nll_log_softmax": [_vp, _i64, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _i64, _vp],
    "acm_glue_fwd": [_i32, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _vp],
    "acm_glue_bwd": [_i32, _i32, _vp, _vp, _vp, _i64, _f32, _vp],
    "acm_set_l2_fetch_granularity": [_i32],
    "acm_set_gather_mode": [_i32],
    "acm_set_narrow_row_hint": [_i32],
    "acm_set_mix_bwd_occupancy": [_i32],
    "acm_set_mix_bwd_ring": [_i32],
    "acm_set_gemm_direct_store": [_i32],
    "acm_gemm_ab": [_i32, _i32, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "acm_linear_fwd": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp],
    "acm_gemm_atb": [_i32, _i32, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp],
    "acm_spmm_agg_first": [_i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    "acm_fused_agg_fwd": [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _i32, _i64, _vp, _vp, _vp, _vp, _vp],
    "acm_spmm_plain": [_i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp],
}
_RESTYPES = {"acm_last_error_string": _c.c_char_p, "acm_launch_count": _i64}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None
_lock = threading.Lock()


class AcmLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise AcmLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m acm_gnn_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback for the ACM layer.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in _PROTOS.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, _c.c_int)
        g = os.environ.get("ACMB200_GATHER")
        if g is not None:
            lib.acm_set_gather_mode({"1": 1, "async": 1, "cp.async": 1, "2": 2, "bulk": 2, "3": 3, "tma": 3, "gather4": 3, "4": 4, "tma-all": 4}.get(g.lower(), 0))
        h = os.environ.get("ACMB200_NARROW_HINT")
        if h is not None:
            lib.acm_set_narrow_row_hint(int(h))
        r = os.environ.get("ACMB200_MIXBWD_RING")
        if r is not None:
            lib.acm_set_mix_bwd_ring(int(r))
        ds = os.environ.get("ACMB200_GEMM_DIRECT")
        if ds is not None:
            lib.acm_set_gemm_direct_store(int(ds))
        o = os.environ.get("ACMB200_MIXBWD_OCC")
        if o is not None:
            lib.acm_set_mix_bwd_occupancy(int(o))
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().acm_last_error_string().decode("utf-8", "replace")
        raise AcmLibraryError(f"{what} failed (code {rc}): {msg}")


class KernelTimer:
    """CUDA-event timing of individual library launches on the launching stream (bench.py's
    live roofline measurement).  Active while installed via ``set_timer``."""

    def __init__(self):
        self.spans = {}

    def span(self, key):
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.spans.setdefault(key, []).append((a, b))
        return a, b

    def summary(self):
        """{key: (launches, total_ms)} -- call after a device synchronize."""
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.spans.items()}


_TIMER = None


def set_timer(t):
    global _TIMER
    _TIMER = t


def call(name, *args, tag=None):
    if _TIMER is None:
        check(getattr(load(), name)(*args), name)
        return
    import torch
    a, b = _TIMER.span(name if tag is None else f"{name}:{tag}")
    st = torch.cuda.current_stream()
    a.record(st)
    check(getattr(load(), name)(*args), name)
    b.record(st)


def launch_count():
    return int(load().acm_launch_count())


def ptr(t):
    """Device pointer of a tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()
