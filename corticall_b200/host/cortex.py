"""CortexGraph / CortexRecord / CortexHeader / CortexColor -- the reference's graph data model, served by
libcorticall_cuda.  Same class and method names as the Java classes so code (and tests) written against
the reference read the same; batch entry points (findRecordIndices, findWindows, findNovel, ...) are additive.

Reference (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
  S/utils/io/graph/DeBruijnGraph.java:16-53          the interface
  S/utils/io/graph/cortex/CortexGraph.java:40-415    reader, iterator, getRecord, findRecord
  S/utils/io/graph/cortex/CortexRecord.java:13-409   record value type
  S/utils/io/graph/cortex/CortexHeader.java, CortexColor.java

Every record decode, lookup and scan below runs on the GPU through the C ABI (include/corticall_cuda.h).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .. import _native as N
from .._native import CortexJDKException
from .kmer import CanonicalKmer, CortexBinaryKmer, CortexByteKmer, decodeBinaryKmer, encodeBinaryKmer, getKmerBits


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data


class CortexColor:
    """S/utils/io/graph/cortex/CortexColor.java"""

    def __init__(self):
        self.meanReadLength = 0
        self.totalSequence = 0
        self.sampleName = ""
        self.errorRate = 0.0
        self.tipClippingApplied = False
        self.lowCovgSupernodesRemoved = False
        self.lowCovgKmersRemoved = False
        self.cleanedAgainstGraph = False
        self.lowCovSupernodesThreshold = 0
        self.lowCovKmerThreshold = 0
        self.cleanedAgainstGraphName = ""

    def getMeanReadLength(self): return self.meanReadLength
    def getTotalSequence(self): return self.totalSequence
    def getSampleName(self): return self.sampleName
    def getErrorRate(self): return self.errorRate
    def isTipClippingApplied(self): return self.tipClippingApplied
    def isLowCovgSupernodesRemoved(self): return self.lowCovgSupernodesRemoved
    def isLowCovgKmersRemoved(self): return self.lowCovgKmersRemoved
    def isCleanedAgainstGraph(self): return self.cleanedAgainstGraph
    def getLowCovSupernodesThreshold(self): return self.lowCovSupernodesThreshold
    def getLowCovKmerThreshold(self): return self.lowCovKmerThreshold
    def getCleanedAgainstGraphName(self): return self.cleanedAgainstGraphName


class CortexHeader:
    """S/utils/io/graph/cortex/CortexHeader.java"""

    def __init__(self):
        self.version = 0
        self.kmerSize = 0
        self.kmerBits = 0
        self.numColors = 0
        self.colors: list[CortexColor] = []

    def getVersion(self): return self.version
    def getKmerSize(self): return self.kmerSize
    def getKmerBits(self): return self.kmerBits
    def getNumColors(self): return self.numColors
    def getColors(self): return self.colors
    def getColor(self, c: int): return self.colors[c]
    def hasColor(self, color: int): return color < len(self.colors)
    def addColor(self, color: CortexColor): self.colors.append(color)


_EDGE_STR = b"acgtACGT"


class CortexRecord:
    """One record: binaryKmer (Java long[] convention: byte-swapped on-disk words), int coverages, edge bytes."""

    __slots__ = ("binaryKmer", "coverages", "edges", "kmerSize", "kmerBits")

    def __init__(self, binaryKmer, coverages, edges, kmerSize: int, kmerBits: int):
        self.binaryKmer = [int(x) for x in binaryKmer]
        self.coverages = [int(x) for x in coverages]
        self.edges = bytes(int(e) & 0xFF for e in edges)
        self.kmerSize = int(kmerSize)
        self.kmerBits = int(kmerBits)

    @classmethod
    def fromString(cls, recordString: str) -> "CortexRecord":
        """CortexRecord(String) :43-102 -- "KMER cov.. edges.." (used by the reference's tests)."""
        pieces = recordString.split()
        kmer = pieces[0].encode()
        ncol = (len(pieces) - 1) // 2
        covs = [int(x) for x in pieces[1:1 + ncol]]
        edges = []
        for es in pieces[1 + ncol:1 + 2 * ncol]:
            e = 0
            for i in range(4):
                if es[i] != ".":
                    e |= 1 << (7 - i)
                if es[i + 4] != ".":
                    e |= 1 << i
            edges.append(e)
        return cls(encodeBinaryKmer(kmer), covs, edges, len(kmer), getKmerBits(len(kmer)))

    def getKmerSize(self): return self.kmerSize
    def getKmerBits(self): return self.kmerBits
    def getNumColors(self): return len(self.coverages)
    def getBinaryKmer(self): return self.binaryKmer
    def getKmerAsBytes(self) -> bytes: return decodeBinaryKmer(self.binaryKmer, self.kmerSize, self.kmerBits)
    def getCortexBinaryKmer(self): return CortexBinaryKmer(self.binaryKmer)
    def getCanonicalKmer(self): return CanonicalKmer(self.getKmerAsBytes(), True)
    def getKmerAsString(self) -> str: return self.getKmerAsBytes().decode()
    def getKmerAsByteKmer(self): return CortexByteKmer(self.getKmerAsBytes())
    def getEdges(self): return self.edges
    def getCoverages(self): return self.coverages
    def getCoverage(self, color: int): return self.coverages[color]

    def getEdgesAsBytes(self, color: int | None = None):   # :117-140
        table = []
        for e in self.edges:
            se = e - 256 if e > 127 else e
            left, right = se >> 4, e & 0xF
            row = bytearray(8)
            for i in range(4):
                row[i] = _EDGE_STR[i] if left & (1 << (3 - i)) else 0x2E
                row[i + 4] = _EDGE_STR[i + 4] if right & (1 << i) else 0x2E
            table.append(bytes(row))
        return table if color is None else table[color]

    def getEdgeAsStrings(self): return [r.decode() for r in self.getEdgesAsBytes()]
    def getEdgesAsString(self, color: int): return self.getEdgesAsBytes(color).decode()

    def getInEdgesAsBytes(self, color: int, complement: bool = False):     # :214-237
        e = self.edges[color]
        left = (e - 256 if e > 127 else e) >> 4
        names = b"TGCA" if complement else b"ACGT"
        return [names[i] for i in range(4) if left & (1 << (3 - i))]

    def getOutEdgesAsBytes(self, color: int, complement: bool = False):    # :251-273
        right = self.edges[color] & 0xF
        names = b"TGCA" if complement else b"ACGT"
        return [names[i] for i in range(4) if right & (1 << i)]

    def getInEdgesAsStrings(self, color: int, complement: bool = False): return [chr(b) for b in self.getInEdgesAsBytes(color, complement)]
    def getOutEdgesAsStrings(self, color: int, complement: bool = False): return [chr(b) for b in self.getOutEdgesAsBytes(color, complement)]
    def getInDegree(self, color: int): return len(self.getInEdgesAsBytes(color, False))
    def getOutDegree(self, color: int): return len(self.getOutEdgesAsBytes(color, False))

    def toString(self, *colors: int) -> str:               # :166-194
        cols = colors if colors else range(len(self.coverages))
        return " ".join([self.getKmerAsString()] + [str(self.coverages[c]) for c in cols] + [self.getEdgesAsString(c) for c in cols])

    __str__ = toString

    def __eq__(self, o):                                   # :200-208
        return (isinstance(o, CortexRecord) and self.binaryKmer == o.binaryKmer and self.coverages == o.coverages
                and self.edges == o.edges)

    def __hash__(self):
        return hash((tuple(self.binaryKmer), tuple(self.coverages), self.edges))

    def compareTo(self, o: "CortexRecord") -> int:         # :210-212
        a, b = self.getKmerAsString(), o.getKmerAsString()
        return (a > b) - (a < b)

    decodeBinaryKmer = staticmethod(decodeBinaryKmer)
    encodeBinaryKmer = staticmethod(encodeBinaryKmer)
    getKmerBitsFor = staticmethod(getKmerBits)


class CortexGraph:
    """Drop-in for uk.ac.ox.well.cortexjdk.utils.io.graph.cortex.CortexGraph over a device-resident record array.

    `CortexGraph(path)` / `CortexGraph(bytes)` parse the header on the host and upload the record body to
    `device`.  The per-record API (iteration, getRecord, findRecord) keeps the reference's semantics; the
    batch API is what FindROIs / Call should use.
    """

    ITER_BLOCK = 1 << 16     # records decoded per GPU call while iterating

    def __init__(self, source, device: int = 0):
        L = N.lib()
        h = N._P()
        self._keep = None
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            buf = np.frombuffer(bytes(source), dtype=np.uint8) if not isinstance(source, np.ndarray) else np.ascontiguousarray(source, dtype=np.uint8)
            N.check(L.cc_open_memory(_ptr(buf), buf.size, device, C.byref(h)))
            self.cortexFile = None
        else:
            self.cortexFile = os.path.abspath(os.fspath(source))
            N.check(L.cc_open(self.cortexFile.encode(), device, C.byref(h)))
        self._h = h
        self._device = device
        self.firstIndex = 0
        self._load_header()
        self.recordsSeen = 0
        self._block = None        # (first, words, cov, edges) of the decoded block the iterator is in
        self._nextRecord = self._record_at(0)

    @classmethod
    def fromDevice(cls, dev_ptr: int, kmerSize: int, numColors: int, numRecords: int, firstIndex: int = 0, device: int = 0,
                   keepalive=None) -> "CortexGraph":
        """Wrap a device-resident record array (a k-mer-range shard, or a generated benchmark body)."""
        self = cls.__new__(cls)
        h = N._P()
        N.check(N.lib().cc_open_device(dev_ptr, kmerSize, getKmerBits(kmerSize), numColors, numRecords, firstIndex, device, C.byref(h)))
        self._h, self._device, self._keep, self.cortexFile = h, device, keepalive, None
        self.firstIndex = firstIndex
        self._load_header()
        self.recordsSeen = 0
        self._block = None
        self._nextRecord = None
        return self

    # ------------------------------------------------------------------ header
    def _load_header(self):
        L = N.lib()
        v, k, s, c = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        n, off, rs = C.c_uint64(), C.c_uint64(), C.c_uint64()
        N.check(L.cc_header(self._h, C.byref(v), C.byref(k), C.byref(s), C.byref(c), C.byref(n), C.byref(off), C.byref(rs)))
        hd = CortexHeader()
        hd.version, hd.kmerSize, hd.kmerBits, hd.numColors = v.value, k.value, s.value, c.value
        buf = C.create_string_buffer(1 << 16)
        for i in range(c.value):
            col = CortexColor()
            N.check(L.cc_color_name(self._h, i, buf, len(buf)))
            col.sampleName = buf.value.decode("latin-1")
            N.check(L.cc_color_graph_name(self._h, i, buf, len(buf)))
            col.cleanedAgainstGraphName = buf.value.decode("latin-1")
            ci = N.ColorInfo()
            N.check(L.cc_color_info_get(self._h, i, C.byref(ci)))
            col.meanReadLength = ci.mean_read_length
            col.totalSequence = ci.total_sequence
            col.tipClippingApplied = bool(ci.tip_clipping)
            col.lowCovgSupernodesRemoved = bool(ci.low_covg_supernodes_removed)
            col.lowCovgKmersRemoved = bool(ci.low_covg_kmers_removed)
            col.cleanedAgainstGraph = bool(ci.cleaned_against_graph)
            col.lowCovSupernodesThreshold = ci.low_cov_supernodes_threshold
            col.lowCovKmerThreshold = ci.low_cov_kmer_threshold
            hd.addColor(col)
        self.header = hd
        self.numRecords, self.dataOffset, self.recordSize = n.value, off.value, rs.value

    def getFile(self): return self.cortexFile
    def getHeader(self): return self.header
    def getVersion(self): return self.header.version
    def getKmerSize(self): return self.header.kmerSize
    def getKmerBits(self): return self.header.kmerBits
    def getNumColors(self): return self.header.numColors
    def getNumRecords(self): return self.numRecords
    def getColors(self): return self.header.colors
    def hasColor(self, color: int): return self.header.hasColor(color)
    def getColor(self, color: int): return self.header.getColor(color)
    def getSampleName(self, color: int): return self.getColor(color).getSampleName()

    def getColorForSampleName(self, sampleName: str) -> int:           # CortexGraph.java:335-354
        out = C.c_int32(-1)
        N.check(N.lib().cc_color_for_sample_name(self._h, sampleName.encode("latin-1"), C.byref(out)))
        return out.value

    def getColorsForSampleNames(self, sampleNames) -> list[int]:       # :356-366
        return [self.getColorForSampleName(s) for s in sampleNames] if sampleNames else []

    # ------------------------------------------------------------------ K1: decode
    def decodeRecords(self, first: int, count: int):
        """Device decode of records [first, first+count) -> (words uint64 [count,s] native order, coverage int32 [count,c], edges uint8 [count,c])."""
        s, c = self.header.kmerBits, self.header.numColors
        words = np.empty((count, s), dtype=np.uint64)
        cov = np.empty((count, c), dtype=np.int32)
        edges = np.empty((count, c), dtype=np.uint8)
        N.check(N.lib().cc_decode_records(self._h, first, count, _ptr(words), _ptr(cov), _ptr(edges)))
        return words, cov, edges

    def getRawRecords(self, first: int, count: int) -> np.ndarray:
        out = np.empty((count, self.recordSize), dtype=np.uint8)
        N.check(N.lib().cc_get_records(self._h, first, count, _ptr(out)))
        return out

    def _make_record(self, words, cov, edges) -> CortexRecord:
        # Java long[] = Long.reverseBytes(native word) (CortexGraph.java:208-209)
        bk = words.byteswap().view(np.int64)
        return CortexRecord(bk, cov, edges, self.header.kmerSize, self.header.kmerBits)

    def _record_at(self, i: int):
        if i >= self.numRecords:                                        # :190,:236 -> null
            return None
        b = self._block
        if b is None or not (b[0] <= i < b[0] + len(b[1])):
            cnt = min(self.ITER_BLOCK, self.numRecords - i)
            b = self._block = (i,) + self.decodeRecords(i, cnt)
        j = i - b[0]
        return self._make_record(b[1][j], b[2][j], b[3][j])

    # ------------------------------------------------------------------ DeBruijnGraph: seek / iterate
    def position(self, i: int | None = None):
        if i is None:
            return self.recordsSeen
        if i < 0:                                                       # :173-175
            raise CortexJDKException("Record index is prefix of range (%d vs 0-%d)" % (i, self.numRecords - 1), N.CC_ERR_RANGE)
        self.recordsSeen = i
        self._nextRecord = self._get_next_record()

    def _get_next_record(self):                                         # getNextRecord :189-237
        if self.recordsSeen < self.numRecords:
            r = self._record_at(self.recordsSeen)
            self.recordsSeen += 1
            return r
        return None

    def getRecord(self, i: int):                                        # :183-187
        self.position(i)
        return self._nextRecord

    def iterator(self):
        self.position(0)
        return self

    def __iter__(self):
        return self.iterator()

    def hasNext(self) -> bool:
        return self._nextRecord is not None

    def next(self):
        cur = self._nextRecord
        self._nextRecord = self._get_next_record()
        if self._nextRecord is None:
            self.close()
        return cur

    def __next__(self):
        if self._nextRecord is None:
            raise StopIteration
        return self.next()

    def remove(self):
        raise NotImplementedError("UnsupportedOperationException")

    def close(self):
        """Like the reference's close() (:264-270) this does NOT invalidate the graph; see dispose()."""

    def dispose(self):
        if getattr(self, "_h", None):
            N.lib().cc_dispose(self._h)
            self._h = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    # ------------------------------------------------------------------ K4: findRecord (single) and batch lookups
    def _as_kmer_bytes(self, kmer) -> bytes:
        if isinstance(kmer, str):
            return kmer.encode("latin-1")
        if isinstance(kmer, CortexByteKmer):
            return kmer.getKmer()
        if isinstance(kmer, CanonicalKmer):
            return kmer.getKmerAsBytes()
        return bytes(kmer)

    def findRecord(self, kmer):                                         # :272-321
        b = self._as_kmer_bytes(kmer)
        if len(b) != self.header.kmerSize:
            # CortexByteKmer.compareTo walks the QUERY length (:43): longer throws, shorter can never equal
            if len(b) > self.header.kmerSize:
                raise IndexError("ArrayIndexOutOfBoundsException: query longer than the graph's k-mer size")
            return None
        idx, raw = self.findRecords(np.frombuffer(b, dtype=np.uint8).reshape(1, -1))
        return self._record_from_raw(raw[0]) if idx[0] >= 0 else None

    def findRecords(self, kmers):
        """cc_find_records: a handful of k-mers (a vertex and its neighbours) in one call and one kernel launch ->
        (int64 indices, uint8 [nq, recordSize] raw records, zeros for misses)."""
        q = np.ascontiguousarray(kmers, dtype=np.uint8)
        if q.ndim != 2 or q.shape[1] != self.header.kmerSize:
            raise ValueError("queries must be [nq, %d] ASCII bytes" % self.header.kmerSize)
        idx = np.empty(q.shape[0], dtype=np.int64)
        raw = np.empty((q.shape[0], self.recordSize), dtype=np.uint8)
        N.check(N.lib().cc_find_records(self._h, _ptr(q), q.shape[0], _ptr(idx), _ptr(raw)))
        return idx, raw

    def _record_from_raw(self, raw: np.ndarray) -> CortexRecord:
        s, c = self.header.kmerBits, self.header.numColors
        words = raw[:8 * s].copy().view("<u8")
        cov = raw[8 * s:8 * s + 4 * c].copy().view("<i4")
        edges = raw[8 * s + 4 * c:8 * s + 5 * c].copy()
        return self._make_record(words, cov, edges)

    def findRecordIndices(self, kmers, algo: int = N.CC_ALGO_AUTO) -> np.ndarray:
        """kmers: uint8 [nq, k] ASCII (any orientation) -> int64 record index per query, -1 = the reference's null."""
        q = np.ascontiguousarray(kmers, dtype=np.uint8)
        if q.ndim != 2 or q.shape[1] != self.header.kmerSize:
            raise ValueError("queries must be [nq, %d] ASCII bytes" % self.header.kmerSize)
        out = np.empty(q.shape[0], dtype=np.int64)
        N.check(N.lib().cc_find_ascii(self._h, _ptr(q), q.shape[0], _ptr(out), algo))
        return out

    def findWindows(self, seq, algo: int = N.CC_ALGO_AUTO) -> np.ndarray:
        """Every k-window of seq looked up (Call.loadChildWalk :2358-2381) -> int64 [len-k+1]."""
        a = np.frombuffer(seq.encode("latin-1") if isinstance(seq, str) else bytes(seq), dtype=np.uint8) \
            if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, dtype=np.uint8)
        nw = max(a.size - self.header.kmerSize + 1, 0)
        out = np.empty(nw, dtype=np.int64)
        if nw:
            N.check(N.lib().cc_find_windows(self._h, _ptr(a), a.size, _ptr(out), algo))
        return out

    def findPacked(self, words, flags=None, algo: int = N.CC_ALGO_AUTO) -> np.ndarray:
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, self.header.kmerBits)
        f = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        out = np.empty(w.shape[0], dtype=np.int64)
        N.check(N.lib().cc_find_packed(self._h, _ptr(w), _ptr(f), w.shape[0], _ptr(out), algo))
        return out

    def containsWindows(self, seq) -> np.ndarray:
        """ROI membership of every window: rois.contains(new CanonicalKmer(window)) (Call.java:191-197,2425-2451)."""
        a = np.frombuffer(seq.encode("latin-1") if isinstance(seq, str) else bytes(seq), dtype=np.uint8)
        nw = max(a.size - self.header.kmerSize + 1, 0)
        out = np.zeros(nw, dtype=np.uint8)
        if nw:
            N.check(N.lib().cc_contains_windows(self._h, _ptr(a), a.size, _ptr(out)))
        return out.astype(bool)

    def buildIndex(self, bits: int = 0):
        N.check(N.lib().cc_build_index(self._h, bits))

    # ------------------------------------------------------------------ K1+K2: the novelty scan
    def findNovel(self, child: int, parents, cap: int | None = None, want_index: bool = True):
        """FindROIs over the whole graph: (count, records uint8 [m, 8s+5] in the writer's 1-colour layout, indices uint64 [m])."""
        par = np.asarray(list(parents), dtype=np.int32)
        O = 8 * self.header.kmerBits + 5
        cap = self.numRecords if cap is None else cap
        # first call with the library's default staging; grows only when the novel set is large
        guess = min(cap, max(65536, self.numRecords // 32))
        for _ in range(2):
            out = np.empty((max(guess, 1), O), dtype=np.uint8)
            idx = np.empty(max(guess, 1), dtype=np.uint64) if want_index else None
            cnt = C.c_uint64(0)
            N.check(N.lib().cc_find_novel(self._h, child, _ptr(par), par.size, _ptr(out), _ptr(idx), guess, C.byref(cnt)))
            if min(cnt.value, cap) <= guess:
                break
            guess = min(cnt.value, cap)
        m = min(cnt.value, guess)
        return cnt.value, out[:m], (idx[:m] if want_index else None)

    def writeRois(self, child: int, parents, out_path) -> int:
        par = np.asarray(list(parents), dtype=np.int32)
        cnt = C.c_uint64(0)
        N.check(N.lib().cc_write_roi_file(self._h, child, _ptr(par), par.size, os.fspath(out_path).encode(), C.byref(cnt)))
        return cnt.value

    @classmethod
    def join(cls, graphs) -> "CortexGraph":
        """The merged view CortexCollection iterates and Join writes (CortexCollection.java:245-293): sorted union of
        the graphs' k-mers, colours concatenated in argument order -- as a new device-resident graph."""
        graphs = list(graphs)
        arr = (N._P * len(graphs))(*[g._h for g in graphs])
        h = N._P()
        N.check(N.lib().cc_join(arr, len(graphs), C.byref(h)))
        self = cls.__new__(cls)
        self._h, self._device, self._keep, self.cortexFile = h, graphs[0]._device, None, None
        self.firstIndex = 0
        self._load_header()
        self.recordsSeen = 0
        self._block = None
        self._nextRecord = self._record_at(0)
        return self

    def remove(self, secondaries):
        """Remove.java:30-88 with this = the primary graph -> (new graph of the kept records, number removed)."""
        sec = list(secondaries)
        arr = (N._P * max(len(sec), 1))(*[g._h for g in sec])
        h = N._P()
        removed = C.c_uint64(0)
        N.check(N.lib().cc_remove(self._h, arr, len(sec), C.byref(h), C.byref(removed)))
        return CortexGraph._adopt(h, self._device), int(removed.value)

    def sorted(self) -> "CortexGraph":
        """Sort (S/commands/utils/Sort.java:19-50): the same records in ascending k-mer order, as a new graph."""
        h = N._P()
        N.check(N.lib().cc_sort(self._h, C.byref(h)))
        other = CortexGraph.__new__(CortexGraph)
        other._h, other._device, other._keep, other.cortexFile = h, self._device, None, None
        other.firstIndex = 0
        other._load_header()
        other.recordsSeen = 0
        other._block = None
        other._nextRecord = other._record_at(0)
        return other

    @classmethod
    def _adopt(cls, handle, device) -> "CortexGraph":
        """Wraps a handle returned by the library (a new device-resident graph)."""
        other = cls.__new__(cls)
        other._h, other._device, other._keep, other.cortexFile = handle, device, None, None
        other.firstIndex = 0
        other._load_header()
        other.recordsSeen = 0
        other._block = None
        other._nextRecord = other._record_at(0)
        return other

    # ---- scan-shaped pre-filters / recovery; each returns a new graph
    def findLowCoverage(self, minCoverage: int) -> "CortexGraph":
        """FindLowCoverage.java:33-66: the records with coverage(0) < minCoverage."""
        h = N._P()
        N.check(N.lib().cc_find_low_coverage(self._h, int(minCoverage), C.byref(h)))
        return CortexGraph._adopt(h, self._device)

    def findShared(self, roi: "CortexGraph", child: int, parents, ignore) -> "CortexGraph":
        """FindShared.java:40-118: records of `roi` with coverage in a colour of this graph outside child/parents/ignore."""
        pa = np.ascontiguousarray(list(parents), dtype=np.int32)
        ig = np.ascontiguousarray(list(ignore), dtype=np.int32)
        h = N._P()
        N.check(N.lib().cc_find_shared(self._h, roi._h, int(child), _ptr(pa) if pa.size else None, pa.size,
                                       _ptr(ig) if ig.size else None, ig.size, C.byref(h)))
        return CortexGraph._adopt(h, self._device)

    def recoverExcludedKmers(self, dirty: "CortexGraph", child: int):
        """RecoverExcludedKmers.java:31-106 -> (new one-colour graph, number of recovered records)."""
        h = N._P()
        rec = C.c_uint64(0)
        N.check(N.lib().cc_recover_excluded_kmers(self._h, dirty._h, int(child), C.byref(h), C.byref(rec)))
        return CortexGraph._adopt(h, self._device), int(rec.value)

    def covStats(self, child: int, parents) -> list[tuple[int, int]]:
        """CovStats.java:33-72 -> [(child coverage, count)] in ascending coverage."""
        pa = np.ascontiguousarray(list(parents), dtype=np.int32)
        nrows = C.c_uint64(0)
        N.check(N.lib().cc_cov_stats(self._h, int(child), _ptr(pa) if pa.size else None, pa.size, None, None, 0, C.byref(nrows)))
        cov = np.zeros(max(nrows.value, 1), dtype=np.int32)
        cnt = np.zeros(max(nrows.value, 1), dtype=np.int32)
        N.check(N.lib().cc_cov_stats(self._h, int(child), _ptr(pa) if pa.size else None, pa.size, _ptr(cov), _ptr(cnt), cov.size, C.byref(nrows)))
        return [(int(a), int(b)) for a, b in zip(cov[:nrows.value], cnt[:nrows.value])]

    def writeGraph(self, out_path) -> None:
        """CortexGraphWriter over the whole graph (header from the colours, then every record)."""
        N.check(N.lib().cc_write_graph(self._h, os.fspath(out_path).encode()))

    def lastStats(self) -> N.Stats:
        st = N.Stats()
        N.check(N.lib().cc_last_stats(self._h, C.byref(st)))
        return st

    # reference's cache counters: there is no LRU here (every lookup is a device search)
    def getCacheHitsByIndex(self): return 0
    def getCacheHitsByKmer(self): return 0

    def toString(self) -> str:                                          # :368-396 (RamUsageEstimator lines omitted)
        info = "file: %s\n----\nbinary version: %d\nkmer size: %d\nbitfields: %d\ncolors: %d\n" % (
            self.cortexFile, self.getVersion(), self.getKmerSize(), self.getKmerBits(), self.getNumColors())
        yn = lambda b: "yes" if b else "no"
        for i, col in enumerate(self.getColors()):
            info += ("-- Color %d --\n  sample name: '%s'\n  mean read length: %d\n  total sequence loaded: (not parsed)\n"
                     "  sequence error rate: (not parsed)\n  tip clipping: %s\n  remove_low_coverage_supernodes: %s\n"
                     "  remove_low_coverage_kmers: %s\n  cleaned against graph: %s\n") % (
                i, col.sampleName, col.meanReadLength, yn(col.tipClippingApplied), yn(col.lowCovgSupernodesRemoved),
                yn(col.lowCovgKmersRemoved), yn(col.cleanedAgainstGraph))
        info += "----\nkmers: %d\n----\n" % self.getNumRecords()
        return info

    __str__ = toString


class CortexMap:
    """CortexMap (S/utils/io/graph/cortex/CortexMap.java:14-160).  The reference pre-loads every record into a
    HashMap<CortexBinaryKmer, CortexRecord> (:22-36) and answers findRecord from it; the key of a query is
    CortexBinaryKmer(byte[]) (:77-99), which -- unlike CortexGraph.findRecord -- first takes the alphanumerically lowest
    orientation (CortexBinaryKmer.java:15-17) and compares packed words, not bytes.  Here the graph's device-resident index
    is the map: the packed key goes to cc_find_packed.  Everything else delegates to the wrapped graph, as in the reference.
    (A HashMap does not need sorted input; the index does: an unsorted file raises CC_ERR_UNSORTED -- Sort repairs it.)"""

    def __init__(self, cortexFile, device: int = 0):
        self.graph = cortexFile if isinstance(cortexFile, CortexGraph) else CortexGraph(cortexFile, device=device)

    def _get(self, cbk: CortexBinaryKmer):
        if len(cbk.binaryKmer) != self.graph.getKmerBits():      # a long[] of another length equals no key of the map
            return None
        from .kmer import _swap64
        words = np.array([_swap64(int(x) & 0xFFFFFFFFFFFFFFFF) for x in cbk.binaryKmer], dtype=np.uint64)
        idx = int(self.graph.findPacked(words.reshape(1, -1))[0])
        return self.graph.getRecord(idx) if idx >= 0 else None

    def findRecord(self, kmer):                                         # :77-99
        if isinstance(kmer, CortexBinaryKmer):
            return self._get(kmer)
        return self._get(CortexBinaryKmer(self.graph._as_kmer_bytes(kmer)))

    def __getattr__(self, name):                                        # position / iterator / getRecord / header getters: :38-75,101-160
        return getattr(self.graph, name)

    def __iter__(self):
        return iter(self.graph)


class ShardedCortexGraph:
    """One graph over several GPUs of this process (cc_open_sharded): the record array is cut into k-mer-range shards, one
    per entry of `devices`; lookups are routed to the owning shard over peer memory and the novelty scan returns one
    globally ordered list.  The batch surface is CortexGraph's (findRecordIndices / findWindows / findPacked / findNovel /
    writeRois) and so are the answers; per-record access goes through `shard(r)`."""

    PLACEMENTS = {"range": 0, "replicate": 1, "auto": 2}

    def __init__(self, source, devices, placement: str = "range"):
        """placement: "range" = k-mer-range shards (graphs larger than one GPU), "replicate" = a full copy per device (no
        exchange: lookups scale with the device count), "auto" = replicas when they fit (cc_open_sharded_placed)."""
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = N._P()
        place = self.PLACEMENTS[placement]
        if isinstance(source, (bytes, bytearray, memoryview)):
            buf = bytes(source)
            N.check(N.lib().cc_open_sharded_memory_placed(buf, len(buf), devs, len(devices), place, C.byref(h)))
            self.cortexFile = None
        else:
            self.cortexFile = os.fspath(source)
            N.check(N.lib().cc_open_sharded_placed(self.cortexFile.encode(), devs, len(devices), place, C.byref(h)))
        self._h, self._keep = h, None
        self._info()

    @classmethod
    def fromDevice(cls, bodies, counts, k: int, c: int, devices, keepalive=None):
        """bodies[r]: device pointer of shard r's records on devices[r] (on-disk layout), counts[r] records each."""
        self = cls.__new__(cls)
        n = len(devices)
        devs = (C.c_int * n)(*[int(d) for d in devices])
        ptrs = (N._P * n)(*[int(b) for b in bodies])
        cnts = (C.c_uint64 * n)(*[int(x) for x in counts])
        h = N._P()
        N.check(N.lib().cc_open_sharded_device(ptrs, cnts, k, getKmerBits(k), c, devs, n, C.byref(h)))
        self._h, self._keep, self.cortexFile = h, keepalive, None
        self._info()
        return self

    def _info(self):
        nd, nr, k, c = C.c_int(0), C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        N.check(N.lib().cc_sharded_info(self._h, C.byref(nd), C.byref(nr), C.byref(k), C.byref(c)))
        self.numShards, self.numRecords, self.kmerSize, self.numColors = nd.value, nr.value, k.value, c.value
        self.kmerBits = getKmerBits(k.value)
        pl = C.c_int(0)
        N.check(N.lib().cc_sharded_placement(self._h, C.byref(pl)))
        self.placement = "replicate" if pl.value == 1 else "range"

    def getNumRecords(self): return self.numRecords
    def getKmerSize(self): return self.kmerSize
    def getKmerBits(self): return self.kmerBits
    def getNumColors(self): return self.numColors

    def shard(self, r: int):
        """(borrowed CortexGraph over shard r, device, first global record index)."""
        h, dev, first = N._P(), C.c_int(0), C.c_uint64(0)
        N.check(N.lib().cc_sharded_shard(self._h, r, C.byref(h), C.byref(dev), C.byref(first)))
        g = CortexGraph._adopt(h, dev.value)
        g.firstIndex = first.value
        g.dispose = lambda: None            # owned by the sharded handle
        g._owner = self
        return g, dev.value, first.value

    def findRecordIndices(self, kmers) -> np.ndarray:
        q = np.ascontiguousarray(kmers, dtype=np.uint8)
        if q.ndim != 2 or q.shape[1] != self.kmerSize:
            raise ValueError("queries must be [nq, %d] ASCII bytes" % self.kmerSize)
        out = np.empty(q.shape[0], dtype=np.int64)
        N.check(N.lib().cc_find_ascii_sharded(self._h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def findWindows(self, seq) -> np.ndarray:
        a = np.frombuffer(seq.encode("latin-1") if isinstance(seq, str) else bytes(seq), dtype=np.uint8) \
            if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, dtype=np.uint8)
        nw = max(a.size - self.kmerSize + 1, 0)
        out = np.empty(nw, dtype=np.int64)
        if nw:
            N.check(N.lib().cc_find_windows_sharded(self._h, _ptr(a), a.size, _ptr(out)))
        return out

    def findPacked(self, words, flags=None) -> np.ndarray:
        w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, self.kmerBits)
        f = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        out = np.empty(w.shape[0], dtype=np.int64)
        N.check(N.lib().cc_find_packed_sharded(self._h, _ptr(w), _ptr(f), w.shape[0], _ptr(out)))
        return out

    def findPackedDevice(self, words, flags, outs):
        """Device-resident batch: words[r] int64 [nq_r, s] / flags[r] uint8 [nq_r] or None / outs[r] int64 [nq_r] are torch
        tensors on the shard's device."""
        n = self.numShards
        pw = (N._P * n)(*[w.data_ptr() for w in words])
        pf = (N._P * n)(*[(f.data_ptr() if f is not None else None) for f in flags]) if flags is not None else None
        po = (N._P * n)(*[o.data_ptr() for o in outs])
        nq = (C.c_uint64 * n)(*[int(w.shape[0]) for w in words])
        N.check(N.lib().cc_find_packed_sharded_dev(self._h, pw, pf, nq, po))

    def findNovel(self, child: int, parents, cap: int | None = None, want_index: bool = True):
        par = np.asarray(list(parents), dtype=np.int32)
        O = 8 * self.kmerBits + 5
        cap = self.numRecords if cap is None else cap
        guess = min(cap, max(65536, self.numRecords // 32))
        for _ in range(2):
            out = np.empty((max(guess, 1), O), dtype=np.uint8)
            idx = np.empty(max(guess, 1), dtype=np.uint64) if want_index else None
            cnt = C.c_uint64(0)
            N.check(N.lib().cc_find_novel_sharded(self._h, child, _ptr(par), par.size, _ptr(out), _ptr(idx), guess, C.byref(cnt)))
            if min(cnt.value, cap) <= guess:
                break
            guess = min(cnt.value, cap)
        m = min(cnt.value, guess)
        return cnt.value, out[:m], (idx[:m] if want_index else None)

    def writeRois(self, child: int, parents, out_path) -> int:
        par = np.asarray(list(parents), dtype=np.int32)
        cnt = C.c_uint64(0)
        N.check(N.lib().cc_write_roi_file_sharded(self._h, child, _ptr(par), par.size, os.fspath(out_path).encode(), C.byref(cnt)))
        return cnt.value

    def lastStats(self) -> N.ShardedStats:
        st = N.ShardedStats()
        N.check(N.lib().cc_sharded_last_stats(self._h, C.byref(st)))
        return st

    def dispose(self):
        if getattr(self, "_h", None):
            N.lib().cc_dispose_sharded(self._h)
            self._h = None

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass


def packCanonical(seq, k: int, device: int = 0):
    """K3 over a host sequence: canonical packed words [nw, s] (native order) and flags [nw]
    (bit0 flipped, bit1 not ACGTacgt, bit2 lower case)."""
    a = np.frombuffer(seq.encode("latin-1") if isinstance(seq, str) else bytes(seq), dtype=np.uint8) \
        if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, dtype=np.uint8)
    s = getKmerBits(k)
    nw = max(a.size - k + 1, 0)
    words = np.zeros((nw, s), dtype=np.uint64)
    flags = np.zeros(nw, dtype=np.uint8)
    if nw:
        N.check(N.lib().cc_pack_canonical(device, _ptr(a), a.size, k, _ptr(words), _ptr(flags)))
    return words, flags
