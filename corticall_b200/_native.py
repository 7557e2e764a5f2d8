"""ctypes binding of libcorticall_cuda.so -- one Python callable per entry point of include/corticall_cuda.h.

The library is the product; this module only marshals pointers and sizes.  There is NO fallback: if the
shared object is missing, importing `lib()` raises, and every compute call on a machine without a CUDA
device returns CC_ERR_CUDA, which surfaces as CortexJDKException.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcorticall_cuda.so")

CC_OK = 0
CC_ERR_NOT_CORTEX, CC_ERR_BAD_VERSION, CC_ERR_BAD_TRAILER, CC_ERR_IO, CC_ERR_UNSORTED = 1, 2, 3, 4, 5
CC_ERR_RANGE, CC_ERR_CUDA, CC_ERR_NCCL, CC_ERR_ARG, CC_ERR_UNSUPPORTED = 6, 7, 8, 9, 10
CC_ALGO_AUTO, CC_ALGO_BSEARCH, CC_ALGO_MERGE = 0, 1, 2


class CortexJDKException(RuntimeError):
    """uk.ac.ox.well.cortexjdk.utils.exceptions.CortexJDKException (S/utils/exceptions/CortexJDKException.java:3-10)."""

    def __init__(self, message: str, status: int = -1):
        super().__init__(message)
        self.status = status


class ColorInfo(C.Structure):
    _fields_ = [("mean_read_length", C.c_uint32), ("total_sequence", C.c_uint64),
                ("tip_clipping", C.c_uint8), ("low_covg_supernodes_removed", C.c_uint8),
                ("low_covg_kmers_removed", C.c_uint8), ("cleaned_against_graph", C.c_uint8),
                ("low_cov_supernodes_threshold", C.c_uint32), ("low_cov_kmer_threshold", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_float), ("total_ms", C.c_float), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("launches", C.c_uint32)]


class ShardedStats(C.Structure):
    _fields_ = [("route_ms", C.c_float), ("search_ms", C.c_float), ("gather_ms", C.c_float), ("chunk_ms", C.c_float),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("launches", C.c_uint32), ("overflow_retries", C.c_uint32)]


_P = C.c_void_p
_U32P = C.POINTER(C.c_uint32)
_U64P = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); must list every CC_API symbol of include/corticall_cuda.h (tests check this)
SIGNATURES = {
    "cc_last_error": (C.c_char_p, []),
    "cc_version": (C.c_char_p, []),
    "cc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cc_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "cc_open_memory": (C.c_int, [_P, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "cc_open_device": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "cc_dispose": (None, [_P]),
    "cc_header": (C.c_int, [_P, _U32P, _U32P, _U32P, _U32P, _U64P, _U64P, _U64P]),
    "cc_color_name": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_size_t]),
    "cc_color_graph_name": (C.c_int, [_P, C.c_uint32, C.c_char_p, C.c_size_t]),
    "cc_color_info_get": (C.c_int, [_P, C.c_uint32, C.POINTER(ColorInfo)]),
    "cc_color_for_sample_name": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int32)]),
    "cc_get_records": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    "cc_decode_records": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P, _P, _P]),
    "cc_decode_records_dev": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P, _P, _P, _P]),
    "cc_find_novel": (C.c_int, [_P, C.c_int32, _P, C.c_int, _P, _P, C.c_uint64, _U64P]),
    "cc_find_novel_dev": (C.c_int, [_P, C.c_int32, _P, C.c_int, _P, _P, C.c_uint64, _P, _P]),
    "cc_find_novel_host": (C.c_int, [C.c_int, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int32, _P, C.c_int,
                                     _P, _P, C.c_uint64, _U64P, C.POINTER(Stats)]),
    "cc_write_roi_file": (C.c_int, [_P, C.c_int32, _P, C.c_int, C.c_char_p, _U64P]),
    "cc_pack_canonical": (C.c_int, [C.c_int, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "cc_pack_canonical_dev": (C.c_int, [C.c_int, _P, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "cc_pack_kmers_dev": (C.c_int, [C.c_int, _P, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "cc_build_index": (C.c_int, [_P, C.c_int]),
    "cc_find_ascii": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_int]),
    "cc_find_ascii_dev": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_int, _P]),
    "cc_find_windows": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_int]),
    "cc_find_windows_dev": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_int, _P]),
    "cc_find_packed": (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_int]),
    "cc_find_packed_dev": (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_int, _P]),
    "cc_find_records": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "cc_contains_windows": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "cc_bucket_by_owner_dev": (C.c_int, [C.c_int, _P, _P, C.c_uint64, C.c_uint32, _P, C.c_int, _P, _P, _P, _P]),
    "cc_scatter_results_dev": (C.c_int, [C.c_int, _P, _P, C.c_uint64, _P, _P]),
    "cc_route_state_bytes": (C.c_int, [C.c_uint64, C.c_int, _U64P]),
    "cc_route_queries_dev": (C.c_int, [C.c_int, _P, _P, C.c_uint64, C.c_uint32, _P, C.c_int, C.c_int, C.c_uint64, _P, _P, _P, C.c_uint64, _P, _P]),
    "cc_publish_counts_dev": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_uint64, _P, _P]),
    "cc_find_routed_dev": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_uint64, _P, _P]),
    "cc_gather_routed_dev": (C.c_int, [C.c_int, _P, _P, C.c_uint64, C.c_uint64, _P, C.c_int, C.c_uint64, _P, _P]),
    "cc_open_sharded": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "cc_open_sharded_memory": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "cc_open_sharded_placed": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(_P)]),
    "cc_open_sharded_memory_placed": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(_P)]),
    "cc_sharded_placement": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "cc_open_sharded_device": (C.c_int, [C.POINTER(_P), _U64P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "cc_dispose_sharded": (None, [_P]),
    "cc_sharded_info": (C.c_int, [_P, C.POINTER(C.c_int), _U64P, _U32P, _U32P]),
    "cc_sharded_shard": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_int), _U64P]),
    "cc_sharded_last_stats": (C.c_int, [_P, C.POINTER(ShardedStats)]),
    "cc_find_packed_sharded": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "cc_find_ascii_sharded": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "cc_find_windows_sharded": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "cc_find_packed_sharded_dev": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _U64P, C.POINTER(_P)]),
    "cc_find_novel_sharded": (C.c_int, [_P, C.c_int32, _P, C.c_int, _P, _P, C.c_uint64, _U64P]),
    "cc_write_roi_file_sharded": (C.c_int, [_P, C.c_int32, _P, C.c_int, C.c_char_p, _U64P]),
    "cc_join": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "cc_remove": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.POINTER(_P), _U64P]),
    "cc_sort": (C.c_int, [_P, C.POINTER(_P)]),
    "cc_write_graph": (C.c_int, [_P, C.c_char_p]),
    "cc_find_low_coverage": (C.c_int, [_P, C.c_int32, C.POINTER(_P)]),
    "cc_find_shared": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int, _P, C.c_int, C.POINTER(_P)]),
    "cc_recover_excluded_kmers": (C.c_int, [_P, _P, C.c_int32, C.POINTER(_P), _U64P]),
    "cc_cov_stats": (C.c_int, [_P, C.c_int32, _P, C.c_int, _P, _P, C.c_uint64, _U64P]),
    "cc_last_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "cc_launch_count": (C.c_uint64, []),
    "cc_device_body": (C.c_int, [_P, C.POINTER(_P), _U64P]),
    "cc_device_keys": (C.c_int, [_P, C.POINTER(_P), _U64P]),
    "cc_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
}

_LIB = None


def lib() -> C.CDLL:
    """Loads libcorticall_cuda.so (built by `__graft_entry__.build()` / `make -C corticall_b200/csrc`)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the k-mer hot path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error() -> str:
    return lib().cc_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status != CC_OK:
        raise CortexJDKException(last_error(), status)


def launch_count() -> int:
    return int(lib().cc_launch_count())


def set_option(name: str, value: int) -> None:
    check(lib().cc_set_option(name.encode(), int(value)))


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().cc_device_count(C.byref(n))
    return n.value if rc == CC_OK else 0
