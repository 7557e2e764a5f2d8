"""ctypes binding of oracle/liborc.so (the C restatement).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcHeader(C.Structure):
    _fields_ = [("version", C.c_uint32), ("kmer_size", C.c_uint32), ("kmer_bits", C.c_uint32), ("num_colors", C.c_uint32),
                ("data_offset", C.c_uint64), ("record_size", C.c_uint64), ("num_records", C.c_uint64)]


class OrcGraph(C.Structure):
    _fields_ = [("file", C.c_void_p), ("file_size", C.c_uint64), ("h", OrcHeader), ("cached_small", C.c_uint32)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "ctx_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liborc.so"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, i64p, u64p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_open.argtypes = [C.POINTER(OrcGraph), u8p, C.c_uint64]; L.orc_open.restype = C.c_int
        L.orc_color_name.argtypes = [C.POINTER(OrcGraph), C.c_uint32, C.c_char_p, C.c_size_t]; L.orc_color_name.restype = C.c_int
        L.orc_color_for_sample_name.argtypes = [C.POINTER(OrcGraph), C.c_char_p]; L.orc_color_for_sample_name.restype = C.c_int
        L.orc_get_record.argtypes = [C.POINTER(OrcGraph), C.c_uint64, i64p, i32p, u8p]; L.orc_get_record.restype = C.c_int
        L.orc_decode_binary_kmer.argtypes = [i64p, C.c_uint32, C.c_uint32, u8p]; L.orc_decode_binary_kmer.restype = None
        L.orc_encode_binary_kmer.argtypes = [u8p, C.c_uint32, i64p]; L.orc_encode_binary_kmer.restype = C.c_int
        L.orc_edges_to_string.argtypes = [C.c_uint8, C.c_char_p]; L.orc_edges_to_string.restype = None
        L.orc_complement.argtypes = [C.c_uint8]; L.orc_complement.restype = C.c_uint8
        L.orc_reverse_complement.argtypes = [u8p, C.c_size_t, u8p]; L.orc_reverse_complement.restype = None
        L.orc_lowest_orientation.argtypes = [u8p, C.c_size_t, u8p]; L.orc_lowest_orientation.restype = C.c_int
        L.orc_byte_kmer_compare.argtypes = [u8p, u8p, C.c_size_t]; L.orc_byte_kmer_compare.restype = C.c_int
        L.orc_find_record.argtypes = [C.POINTER(OrcGraph), u8p]; L.orc_find_record.restype = C.c_int64
        L.orc_is_novel.argtypes = [i32p, i32p, C.c_int, C.c_int32]; L.orc_is_novel.restype = C.c_int
        L.orc_find_rois.argtypes = [C.POINTER(OrcGraph), C.c_int32, i32p, C.c_int, u8p, u64p, C.c_uint64, C.c_int]
        L.orc_find_rois.restype = C.c_uint64
        L.orc_find_rois_body.argtypes = [u8p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, i32p, C.c_int,
                                         u8p, u64p, C.c_uint64, C.c_int]
        L.orc_find_rois_body.restype = C.c_uint64
        L.orc_write_roi_header.argtypes = [C.c_uint32, C.c_uint32, C.c_char_p, u8p, C.c_size_t]; L.orc_write_roi_header.restype = C.c_size_t
        L.orc_find_windows.argtypes = [C.POINTER(OrcGraph), u8p, C.c_uint64, i64p]; L.orc_find_windows.restype = None
        L.orc_find_batch.argtypes = [C.POINTER(OrcGraph), u8p, C.c_uint64, i64p]; L.orc_find_batch.restype = None
        L.orc_pack_windows.argtypes = [u8p, C.c_uint64, C.c_uint32, u64p, u8p]; L.orc_pack_windows.restype = None
        G = C.POINTER(OrcGraph)
        L.orc_find_low_coverage.argtypes = [G, C.c_int32, u8p]; L.orc_find_low_coverage.restype = None
        L.orc_find_shared.argtypes = [G, G, C.c_int32, i32p, C.c_int, i32p, C.c_int, u8p]; L.orc_find_shared.restype = C.c_int
        L.orc_recover_excluded_kmers.argtypes = [G, G, C.c_int32, u8p, i32p]; L.orc_recover_excluded_kmers.restype = C.c_uint64
        L.orc_cov_stats_pairs.argtypes = [G, C.c_int32, i32p, C.c_int, i32p, i32p]; L.orc_cov_stats_pairs.restype = None
        L.orc_remove.argtypes = [C.POINTER(G), C.c_int, u8p, C.c_uint64, u64p]; L.orc_remove.restype = C.c_uint64
        _LIB = L
    return _LIB


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


class Graph:
    """orc_graph over an in-memory .ctx image."""

    def __init__(self, data: bytes | np.ndarray):
        self._buf = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data, dtype=np.uint8)
        self.g = OrcGraph()
        self.rc = lib().orc_open(C.byref(self.g), _p(self._buf), len(self._buf))

    @property
    def ok(self) -> bool:
        return self.rc == 0

    @property
    def h(self) -> OrcHeader:
        return self.g.h

    def color_name(self, c: int) -> str:
        b = C.create_string_buffer(256)
        n = lib().orc_color_name(C.byref(self.g), c, b, 256)
        assert n >= 0
        return b.value.decode("latin-1")

    def color_for_sample_name(self, name: str) -> int:
        return lib().orc_color_for_sample_name(C.byref(self.g), name.encode())

    def get_record(self, i: int):
        s, c = self.h.kmer_bits, self.h.num_colors
        bk = np.zeros(s, dtype=np.int64); cov = np.zeros(c, dtype=np.int32); ed = np.zeros(c, dtype=np.uint8)
        rc = lib().orc_get_record(C.byref(self.g), i, _p(bk), _p(cov), _p(ed))
        return None if rc else (bk, cov, ed)

    def kmer_string(self, bk: np.ndarray) -> bytes:
        out = np.zeros(self.h.kmer_size, dtype=np.uint8)
        lib().orc_decode_binary_kmer(_p(bk), self.h.kmer_size, self.h.kmer_bits, _p(out))
        return out.tobytes()

    def find_record(self, kmer: bytes) -> int:
        q = np.frombuffer(kmer, dtype=np.uint8)
        assert len(q) == self.h.kmer_size
        return int(lib().orc_find_record(C.byref(self.g), _p(q)))

    def find_batch(self, kmers: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(kmers, dtype=np.uint8).reshape(-1, self.h.kmer_size)
        out = np.empty(len(q), dtype=np.int64)
        lib().orc_find_batch(C.byref(self.g), _p(q), len(q), _p(out))
        return out

    def find_windows(self, seq: bytes | np.ndarray) -> np.ndarray:
        a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
        nw = max(len(a) - self.h.kmer_size + 1, 0)
        out = np.empty(nw, dtype=np.int64)
        lib().orc_find_windows(C.byref(self.g), _p(a), len(a), _p(out))
        return out

    def find_rois(self, child: int, parents: list[int], faithful: bool = False):
        par = np.asarray(parents, dtype=np.int32)
        n = self.h.num_records
        osz = 8 * self.h.kmer_bits + 5
        out = np.empty(n * osz, dtype=np.uint8); idx = np.empty(n, dtype=np.uint64)
        cnt = lib().orc_find_rois(C.byref(self.g), child, _p(par), len(par), _p(out), _p(idx), n, int(faithful))
        return out[:cnt * osz].tobytes(), idx[:cnt].copy()


def _i32(a):
    return np.ascontiguousarray(list(a) if not isinstance(a, np.ndarray) else a, dtype=np.int32)


def find_low_coverage_mask(roi: "Graph", min_coverage: int) -> np.ndarray:
    out = np.zeros(max(roi.h.num_records, 1), dtype=np.uint8)
    lib().orc_find_low_coverage(C.byref(roi.g), int(min_coverage), _p(out))
    return out[:roi.h.num_records].astype(bool)


def find_shared_mask(graph: "Graph", roi: "Graph", child: int, parents, ignore):
    """mask of the shared ROI records, or None when the reference would hit its NullPointerException."""
    out = np.zeros(max(roi.h.num_records, 1), dtype=np.uint8)
    pa, ig = _i32(parents), _i32(ignore)
    rc = lib().orc_find_shared(C.byref(graph.g), C.byref(roi.g), int(child), _p(pa) if pa.size else None, pa.size,
                               _p(ig) if ig.size else None, ig.size, _p(out))
    return None if rc else out[:roi.h.num_records].astype(bool)


def recover_excluded_kmers_decisions(graph: "Graph", dirty: "Graph", child: int):
    n = graph.h.num_records
    written = np.zeros(max(n, 1), dtype=np.uint8)
    cov0 = np.zeros(max(n, 1), dtype=np.int32)
    rec = lib().orc_recover_excluded_kmers(C.byref(graph.g), C.byref(dirty.g), int(child), _p(written), _p(cov0))
    return written[:n], cov0[:n], int(rec)


def cov_stats_pairs(graph: "Graph", child: int, parents):
    n = graph.h.num_records
    key = np.zeros(max(n, 1), dtype=np.int32)
    weight = np.zeros(max(n, 1), dtype=np.int32)
    pa = _i32(parents)
    lib().orc_cov_stats_pairs(C.byref(graph.g), int(child), _p(pa) if pa.size else None, pa.size, _p(key), _p(weight))
    return key[:n], weight[:n]


def remove_records(primary: "Graph", secondaries: list["Graph"]):
    """Remove.java over the merged collection, record by record -> (kept records as bytes in the primary's layout, removed)."""
    gs = [primary] + list(secondaries)
    arr = (C.POINTER(OrcGraph) * len(gs))(*[C.pointer(g.g) for g in gs])
    cap = sum(g.h.num_records for g in gs)
    osz = 8 * primary.h.kmer_bits + 5 * primary.h.num_colors
    out = np.empty(max(cap, 1) * osz, dtype=np.uint8)
    removed = C.c_uint64(0)
    kept = int(lib().orc_remove(arr, len(gs), _p(out), cap, C.byref(removed)))
    return out[:kept * osz].tobytes(), int(removed.value)


def find_rois_body(body: np.ndarray, n: int, k: int, s: int, c: int, child: int, parents, faithful=False, cap=None):
    par = np.asarray(parents, dtype=np.int32)
    cap = n if cap is None else cap
    osz = 8 * s + 5
    out = np.empty(max(cap, 1) * osz, dtype=np.uint8); idx = np.empty(max(cap, 1), dtype=np.uint64)
    cnt = int(lib().orc_find_rois_body(_p(body), n, k, s, c, child, _p(par), len(par), _p(out), _p(idx), cap, int(faithful)))
    m = min(cnt, cap)
    return cnt, out[:m * osz], idx[:m]


def pack_windows(seq: bytes | np.ndarray, k: int):
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, dtype=np.uint8)
    s = (k + 31) // 32
    nw = max(len(a) - k + 1, 0)
    words = np.zeros((nw, s), dtype=np.uint64); flags = np.zeros(nw, dtype=np.uint8)
    lib().orc_pack_windows(_p(a), len(a), k, _p(words), _p(flags))
    return words, flags


def lowest_orientation(kmer: bytes):
    a = np.frombuffer(kmer, dtype=np.uint8); out = np.empty_like(a)
    fl = lib().orc_lowest_orientation(_p(a), len(a), _p(out))
    return out.tobytes(), bool(fl)


def roi_header(k: int, s: int, name: str) -> bytes:
    out = np.zeros(76 + len(name) + 8, dtype=np.uint8)
    n = lib().orc_write_roi_header(k, s, name.encode("latin-1"), _p(out), len(out))
    return out[:n].tobytes()
