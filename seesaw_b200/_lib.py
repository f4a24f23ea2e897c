"""ctypes binding of libseesaw_b200.so (the C ABI in include/seesaw_b200.h).

There is no fallback: if the shared library is missing this module raises at import, and every
compute call raises :class:`SeesawB200Error` when no sm_100 device is present."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libseesaw_b200.so")

SSW_F32, SSW_F16 = 0, 1
SSW_BOX_I32, SSW_BOX_F32, SSW_BOX_F64 = 0, 1, 2
SSW_MAX_TOPK = 2048
SSW_MAX_BATCH = 64
SSW_MAX_KNN_K1 = 64
SSW_MAX_WORLD = 8


class SeesawB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libseesaw_b200 error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C seesaw_b200/csrc` (nvcc, sm_100a). seesaw_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)

# name -> (restype, argtypes); kept in one table so tests can check every header symbol is bound
SIGNATURES = {
    "ssw_last_error": (C.c_char_p, []),
    "ssw_version": (C.c_int, []),
    "ssw_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ssw_db_create": (C.c_int, [C.POINTER(_p), C.c_int, _p, C.c_int, C.c_int, C.c_int64, C.c_int, _p, C.c_int64]),
    "ssw_db_create_device": (C.c_int, [C.POINTER(_p), C.c_int, _p, C.c_int, C.c_int, C.c_int64, C.c_int, _p, C.c_int64]),
    "ssw_db_create_synthetic": (C.c_int, [C.POINTER(_p), C.c_int, C.c_int, C.c_int64, C.c_int, _p, C.c_int64,
                                          C.c_uint64, C.c_int]),
    "ssw_db_destroy": (C.c_int, [_p]),
    "ssw_db_info": (C.c_int, [_p, _i64p, _i64p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ssw_db_vectors_device": (C.c_int, [_p, C.POINTER(_p)]),
    "ssw_scan_topk": (C.c_int, [_p, _p, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p]),
    "ssw_scan_topk_device": (C.c_int, [_p, _p, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p]),
    "ssw_exclude_words": (C.c_int, [_p, _i64p]),
    "ssw_exclude_build_device": (C.c_int, [_p, _p, _p, C.c_int, C.c_int64, _p, _p]),
    "ssw_merge_topk_device": (C.c_int, [C.c_int, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p]),
    "ssw_xchg_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_p), _p, _i64p]),
    "ssw_xchg_open": (C.c_int, [C.c_int, _p, C.POINTER(_p)]),
    "ssw_xchg_close": (C.c_int, [C.c_int, _p]),
    "ssw_xchg_destroy": (C.c_int, [C.c_int, _p]),
    "ssw_scan_topk_sharded_device": (C.c_int, [_p, _p, C.c_int, C.c_int, _p, C.POINTER(_p), C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_uint32, _p, _p, _p, _p, _p, _p]),
    "ssw_scan_topk_sharded_pipelined_device": (C.c_int, [_p, _p, C.c_int, C.c_int, _p, C.POINTER(_p), C.c_int, C.c_int, C.c_int,
                                                         C.c_int, C.c_uint32, _p, _p, _p, _p, _p, _p]),
    "ssw_scan_pipeline_drain": (C.c_int, [_p, _p]),
    "ssw_scan_pipeline_side_sms": (C.c_int, [_p, C.c_int]),
    "ssw_scan_pipeline_side_sms_in_use": (C.c_int, [_p, C.POINTER(C.c_int)]),
    "ssw_scan_topk_sharded": (C.c_int, [_p, _p, C.c_int, C.c_int, _p, _p, C.POINTER(_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_uint32, _p, _p, _p, _p]),
    "ssw_set_scan_mode": (C.c_int, [_p, C.c_int]),
    "ssw_db_set_boxes": (C.c_int, [_p, _p, _p, _p, _p, _p]),
    "ssw_db_set_boxes_typed": (C.c_int, [_p, C.c_int, _p, _p, _p, _p, _p]),
    "ssw_db_attach_exact": (C.c_int, [_p, _p]),
    "ssw_db_exact_info": (C.c_int, [_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), _i64p, _i64p]),
    "ssw_rescore": (C.c_int, [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "ssw_topk_from_scores": (C.c_int, [_p, _p, _p, C.c_int, _p, C.c_int64, _p, _p, _p, _p]),
    "ssw_topk_from_order": (C.c_int, [_p, _p, C.c_int64, C.c_int, _p, C.c_int64, _p, _p, _p, _p]),
    "ssw_score_all": (C.c_int, [_p, _p, _p]),
    "ssw_score_all_device": (C.c_int, [_p, _p, _p, _p]),
    "ssw_knn_build": (C.c_int, [C.c_int, _p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _p, _p]),
    "ssw_knn_exact_stats": (C.c_int, [_i64p, _i64p, C.POINTER(C.c_double)]),
    "ssw_knn_build_device": (C.c_int, [C.c_int, _p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, _p, _p, _p]),
    "ssw_knn_graph": (C.c_int, [C.c_int, _p, C.c_int, C.c_int64, C.c_int, C.c_int, _p, _p, _p, _p, C.c_int64, _i64p]),
    "ssw_knn_edges_workspace_bytes": (C.c_int, [C.c_int64, _i64p]),
    "ssw_knn_edges_device": (C.c_int, [C.c_int, _p, _p, C.c_int64, C.c_int, C.c_int64, _p, _p, _p, _p, _p, _p, _p]),
    "ssw_weight_matrix": (C.c_int, [C.c_int, _p, _p, _p, C.c_int64, C.c_int64, _p, _p, _p, C.c_int64, _i64p, _p]),
    "ssw_lp_create": (C.c_int, [C.POINTER(_p), C.c_int, C.c_int64, _p, _p, _p, _p, C.c_double]),
    "ssw_lp_destroy": (C.c_int, [_p]),
    "ssw_lp_fit": (C.c_int, [_p, _p, _p, C.c_int64, _p, _p, C.c_int, C.c_double, _p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ssw_scan_stats": (C.c_int, [_p, C.c_int, _i64p, _i64p]),
    "ssw_lp_fit_scaled": (C.c_int, [_p, _p, _p, C.c_int64, _p, _p, C.c_int, C.c_double, _p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ssw_kernel_launch_count": (C.c_int64, []),
    "ssw_profile_enable": (C.c_int, [_p, C.c_int]),
    "ssw_profile_read": (C.c_int, [_p, C.POINTER(C.c_double), _i64p]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)
    _f.restype, _f.argtypes = _res, _args


def check(code):
    if code != 0:
        raise SeesawB200Error(code, lib.ssw_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    check(lib.ssw_device_count(C.byref(n)))
    return n.value


def kernel_launch_count() -> int:
    return int(lib.ssw_kernel_launch_count())


def ptr(a):
    """Pointer of a C-contiguous numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
