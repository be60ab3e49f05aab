"""ctypes binding of libr4d.so (the C ABI declared in include/r4d.h).

There is deliberately NO fallback: if the shared object is missing, or a call fails, an exception is raised.
torch is only used by callers for device memory and streams; no torch type crosses this boundary.
"""
import ctypes
import os
import re

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libr4d.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "r4d.h")

R4D_IDX_NONE = 0x7FFFFFFF
R4D_TOPK_MAX = 32
R4D_E_ARG, R4D_E_CUDA, R4D_E_WORKSPACE = -1, -2, -3

DENSE_HALF_COS, DENSE_COS_DECAY, DENSE_HALF_COS_DECAY = 0, 1, 2
PREC_BF16, PREC_BF16X3 = 0, 1


class R4DError(RuntimeError):
    pass


_c = ctypes
_vp, _i32, _i64, _sz, _f32, _f64 = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_size_t, _c.c_float, _c.c_double

# name -> (restype, argtypes); must list every function include/r4d.h declares (tests/test_abi.py checks this)
PROTOTYPES = {
    "r4d_version": (_c.c_int, []),
    "r4d_last_error": (_c.c_char_p, []),
    "r4d_device_ok": (_c.c_int, []),
    "r4d_set_option": (_c.c_int, [_c.c_char_p, _c.c_int]),
    "r4d_bitset_words": (_i32, [_i32]),
    "r4d_bitset_pitch_words": (_i32, [_i32]),
    "r4d_bitset_encode": (_c.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "r4d_jaccard_full": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i64, _i64, _vp, _i64, _vp,
                                    _i64, _vp]),
    "r4d_jaccard_topk_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "r4d_jaccard_topk": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp,
                                    _vp, _sz, _vp]),
    "r4d_jaccard_topk_scatter": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _i32, _i32,
                                            _vp, _sz, _vp]),
    "r4d_jaccard_topk_merge": (_c.c_int, [_vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "r4d_postings_index_bytes": (_sz, [_i64, _i32, _i64]),
    "r4d_postings_build_workspace_bytes": (_sz, [_i64, _i32]),
    "r4d_postings_build": (_c.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _sz, _vp, _sz, _vp]),
    "r4d_jaccard_topk_postings_workspace_bytes": (_sz, [_i64]),
    "r4d_jaccard_topk_postings": (_c.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i32, _i64, _i32, _i32, _i64, _i64, _vp,
                                             _vp, _vp, _vp, _sz, _vp]),
    "r4d_jaccard_topk_postings_relay_bytes": (_sz, [_i64, _i32]),
    "r4d_jaccard_topk_postings_packed": (_c.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i32, _i64, _i32, _i32, _i64, _i64, _vp,
                                                    _vp, _vp, _vp, _sz, _vp]),
    "r4d_jaccard_topk_postings_scatter": (_c.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i32, _i64, _i32, _i32, _i64, _i64,
                                                     _vp, _i32, _i32, _vp, _sz, _vp]),
    "r4d_rank_rows_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "r4d_rank_rows_f64": (_c.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "r4d_rank_rows_f32": (_c.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "r4d_topk_rows_f64": (_c.c_int, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp]),
    "r4d_triplet_mine_f64": (_c.c_int, [_vp, _vp, _i64, _i64, _f64, _i32, _vp, _vp, _vp, _vp]),
    "r4d_triplet_mine": (_c.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _i64, _f64, _i32, _i32, _vp, _vp, _vp, _i64, _vp,
                                    _vp, _vp, _vp, _vp, _vp]),
    "r4d_mt19937_choice_replay": (_c.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "r4d_triplet_sample": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _c.c_uint64, _vp, _vp]),
    "r4d_dense_dpad": (_i32, [_i32]),
    "r4d_dense_prepare": (_c.c_int, [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp]),
    "r4d_dense_topk_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "r4d_dense_topk": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _i32, _i32, _i64, _vp,
                                  _vp, _vp, _sz, _vp]),
    "r4d_dense_topk_scatter": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _i32, _i32, _i64, _vp,
                                          _i32, _i32, _vp, _sz, _vp]),
    "r4d_dense_full": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _f32, _i32, _vp, _i64, _vp]),
    "r4d_dense_topk_merge": (_c.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "r4d_profile_read": (_c.c_int, [_c.c_char_p, _vp, _vp]),
    "r4d_kernel_launches": (_i64, []),
    "r4d_meanpool_workspace_bytes": (_sz, [_i64, _i32]),
    "r4d_meanpool_prepare": (_c.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r4d_format_int_rows_bound": (_sz, [_i64, _i64]),
    "r4d_format_int_rows": (_i64, [_vp, _i64, _i64, _i64, _vp, _sz]),
    "r4d_format_lut_rows": (_i64, [_vp, _i64, _i64, _i64, _vp, _vp, _i32, _vp, _sz]),
    "r4d_format_rows_device_sizes": (_c.c_int, [_vp, _i64, _i64, _i64, _vp, _i32, _vp, _vp, _vp]),
    "r4d_format_rows_device": (_c.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "r4d_parse_rows_count": (_i64, [_vp, _sz, _vp]),
    "r4d_parse_int_rows": (_i64, [_vp, _sz, _vp, _vp, _i64, _i64]),
    "r4d_parse_float_rows": (_i64, [_vp, _sz, _vp, _vp, _i64, _i64]),
}

_lib = None


def header_functions():
    """Names of all functions declared in include/r4d.h (used by the ABI test)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r4d_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load libr4d.so once and set prototypes.  Raises R4DError when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise R4DError(
            f"{LIB_PATH} is missing: build it with `python -m rag4dyg_b200.build` (or __graft_entry__.build()). "
            "There is no CPU fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().r4d_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise R4DError(f"{what} failed (rc={rc}): {last_error()}")


def set_option(key, value):
    """r4d_set_option: returns the previous value."""
    prev = load().r4d_set_option(key.encode(), int(value))
    if prev == R4D_E_ARG and key not in ("dense_pair_qres",):
        raise R4DError(last_error())
    return prev


def profile_read(kernel):
    """r4d_profile_read: (summed ms, launches) of the events recorded since the last read (option "kernel_timing")."""
    ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
    check(load().r4d_profile_read(kernel.encode(), ctypes.byref(ms), ctypes.byref(n)), "r4d_profile_read")
    return ms.value, n.value


def require_device():
    """Fail loudly unless an sm_100 device is usable."""
    if not load().r4d_device_ok():
        raise R4DError("libr4d: no usable sm_100 CUDA device: " + last_error())
