"""ctypes binding of include/rvcr.h (librvcr.so).  No CPU fallback: if the library is missing it is
built with nvcc; if that fails, or no B200 is visible at call time, the error is raised loudly."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librvcr.so")

RVCR_OK = 0
ERR_NAMES = {-1: "RVCR_ERR_INVALID", -2: "RVCR_ERR_IO", -3: "RVCR_ERR_FORMAT", -4: "RVCR_ERR_CUDA",
             -5: "RVCR_ERR_UNSUPPORTED", -6: "RVCR_ERR_STATE", -7: "RVCR_ERR_OOM"}


class RvcrError(RuntimeError):
    """Raised for every non-zero return code of the C ABI (the reference's own try/except at
    pipeline.py:554-559 / extract_index.py:74-78 turns it into 'no retrieval' / a printed message)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Info(C.Structure):
    _fields_ = [("d", C.c_int32), ("is_trained", C.c_int32), ("ntotal", C.c_int64), ("nlist", C.c_int64),
                ("nprobe", C.c_int64), ("nstored", C.c_int64), ("owned_begin", C.c_int64),
                ("owned_end", C.c_int64), ("device", C.c_int32), ("metric", C.c_int32)]


class FileInfo(C.Structure):
    _fields_ = [("d", C.c_int32), ("is_trained", C.c_int32), ("ntotal", C.c_int64), ("nlist", C.c_int64),
                ("nprobe", C.c_int64), ("metric", C.c_int32), ("layout", C.c_int32), ("direct_map_type", C.c_int32),
                ("list_encoding", C.c_int32), ("nvectors", C.c_int64), ("max_list_size", C.c_int64),
                ("nonempty_lists", C.c_int64), ("file_bytes", C.c_int64), ("payload_offset", C.c_int64),
                ("searchable", C.c_int32), ("fourcc", C.c_char * 8), ("quantizer_fourcc", C.c_char * 8),
                ("reserved32", C.c_int32), ("reserved", C.c_int64 * 4)]


class MbkStats(C.Structure):
    _fields_ = [("n_steps", C.c_int64), ("max_steps", C.c_int64), ("batch_size", C.c_int64), ("init_size", C.c_int64),
                ("n_reassigned", C.c_int64), ("kernels_launched", C.c_int64), ("converged", C.c_int32), ("rows_resident", C.c_int32),
                ("ewa_inertia", C.c_double), ("init_inertia", C.c_double), ("reserved", C.c_int64 * 4)]


class Stats(C.Structure):
    _fields_ = [("calls_search", C.c_int64), ("queries", C.c_int64), ("kernels_launched", C.c_int64),
                ("coarse_tc_queries", C.c_int64), ("coarse_candidates", C.c_int64),
                ("coarse_fallback", C.c_int64), ("scan_items", C.c_int64), ("scan_tc_candidates", C.c_int64),
                ("scan_tc_fallback", C.c_int64), ("reserved", C.c_int64 * 6)]


_H = C.c_void_p          # rvcr_index*
_P = C.c_void_p          # any device / host pointer
_I64 = C.c_int64
_I32 = C.c_int32

# name -> (restype, argtypes); must list every symbol include/rvcr.h declares (tests check this)
SIGNATURES = {
    "rvcr_last_error": (C.c_char_p, []),
    "rvcr_version": (C.c_int, []),
    "rvcr_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rvcr_index_create": (C.c_int, [_I32, _I64, C.c_int, C.POINTER(_H)]),
    "rvcr_index_read": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_H)]),
    "rvcr_index_probe": (C.c_int, [C.c_char_p, C.POINTER(FileInfo), _P, _I64]),
    "rvcr_shard_range": (C.c_int, [_P, _I64, C.c_int, C.c_int, C.POINTER(_I64), C.POINTER(_I64)]),
    "rvcr_index_read_shard": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(_H)]),
    "rvcr_index_set_shard": (C.c_int, [_H, _I64, _I64]),
    "rvcr_index_write": (C.c_int, [_H, C.c_char_p]),
    "rvcr_index_free": (None, [_H]),
    "rvcr_index_info": (C.c_int, [_H, C.POINTER(Info)]),
    "rvcr_index_set_nprobe": (C.c_int, [_H, _I64]),
    "rvcr_index_set_centroids": (C.c_int, [_H, _P, _I64]),
    "rvcr_index_get_centroids": (C.c_int, [_H, _P]),
    "rvcr_index_list_sizes": (C.c_int, [_H, _P]),
    "rvcr_index_train": (C.c_int, [_H, _P, _I64]),
    "rvcr_index_add": (C.c_int, [_H, _P, _I64, C.c_int, _P]),
    "rvcr_index_add_preassigned": (C.c_int, [_H, _P, _I64, _P, _P]),
    "rvcr_index_alloc_lists": (C.c_int, [_H, _P, _I64]),
    "rvcr_index_device_ptrs": (C.c_int, [_H, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "rvcr_index_commit": (C.c_int, [_H, _P]),
    "rvcr_assign": (C.c_int, [_H, _P, _I64, _P, _P, _P]),
    "rvcr_search": (C.c_int, [_H, _P, _I64, _I32, _P, _P, _P]),
    "rvcr_search_host": (C.c_int, [_H, _P, _I64, _I32, _P, _P]),
    "rvcr_retrieve_blend": (C.c_int, [_H, _P, _I64, C.c_double, _P, _P, _P]),
    "rvcr_retrieve_blend_upsample": (C.c_int, [_H, _P, _I64, C.c_double, _P, C.c_double, _P, _P, _P, _P]),
    "rvcr_upsample_protect": (C.c_int, [_P, _P, _P, _I64, _I32, C.c_double, _P, C.c_int, _P]),
    "rvcr_final_proj": (C.c_int, [_P, _I64, _I32, _P, _P, _I32, _P, C.c_int, _P]),
    "rvcr_final_proj_retrieve_blend": (C.c_int, [_H, _P, _I64, _I32, _P, _P, _P, C.c_double, _P, _P, _P]),
    "rvcr_retrieve_blend_host": (C.c_int, [_H, _P, _I64, C.c_double]),
    "rvcr_retrieve_blend_assigned": (C.c_int, [_H, _P, _I64, _P, C.c_double, _P, _P, _P]),
    "rvcr_retrieve_blend_assigned_push": (C.c_int, [_H, _P, _I64, _P, C.c_double, _P, _P, _P, _P]),
    "rvcr_peer_alloc": (C.c_int, [C.c_int, _I64, C.POINTER(_P), _P]),
    "rvcr_peer_open": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "rvcr_peer_close": (C.c_int, [C.c_int, _P]),
    "rvcr_peer_free": (C.c_int, [C.c_int, _P]),
    "rvcr_peer_pull_rows": (C.c_int, [_P, _P, _P, _I64, _I32, _P, C.c_int, _P]),
    "rvcr_peer_dispatch": (C.c_int, [_P, _I64, _I64, _I32, _I32, _P, _P, _I64, C.c_int, _P]),
    "rvcr_peer_barrier": (C.c_int, [_P, _I32, _I32, _I32, C.c_int, _P]),
    "rvcr_retrieve_blend_inbox_push": (C.c_int, [_H, _P, _I64, _P, C.c_double, _P, _P]),
    "rvcr_search_assigned": (C.c_int, [_H, _P, _I64, _P, _P, _P, _P]),
    "rvcr_reconstruct_n": (C.c_int, [_H, _I64, _I64, _P, C.c_int]),
    "rvcr_merge_coarse": (C.c_int, [_P, _P, _I32, _I64, _P, C.c_int, _P]),
    "rvcr_merge_topk": (C.c_int, [_P, _P, _I32, _I64, _I32, _P, _P, C.c_int, _P]),
    "rvcr_set_option": (C.c_int, [_H, C.c_char_p, _I64]),
    "rvcr_get_option": (C.c_int, [_H, C.c_char_p, C.POINTER(_I64)]),
    "rvcr_get_stats": (C.c_int, [_H, C.POINTER(Stats)]),
    "rvcr_get_phase_ms": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "rvcr_minibatch_kmeans": (C.c_int, [_P, _I64, _I32, _I64, _I64, _I64, _I64, C.c_double, _I64, C.c_uint64, C.c_int, C.c_int, _P,
                                        C.POINTER(MbkStats)]),
    "rvcr_debug_rand_perm": (C.c_int, [_I64, _I64, _P]),
    "rvcr_debug_split_clusters": (C.c_int, [_I32, _I64, _I64, _P, _P, C.POINTER(_I32)]),
    "rvcr_debug_coarse_tc": (C.c_int, [_H, _P, _I64, _I32, _I32, _I32, _P, _P, C.POINTER(_I32), _P,
                                       C.POINTER(C.c_float), _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load (building first if needed) librvcr.so.  Raises if it cannot be produced."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                from . import build as _build
                _build.build()
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)  # AttributeError == header / library mismatch: fail loudly
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(code: int) -> None:
    if code != RVCR_OK:
        msg = lib().rvcr_last_error()
        raise RvcrError(code, msg.decode("utf-8", "replace") if msg else "")


def device_count() -> int:
    n = C.c_int(0)
    check(lib().rvcr_device_count(C.byref(n)))
    return n.value
