"""numpy restatement of Corticall's k-mer hot path.  TEST INFRASTRUCTURE ONLY.

Second, independent oracle (the first is oracle/ctx_oracle.c).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module; the product (corticall_b200/, the CUDA
library) never does.  Both oracles are pinned against the reference's own golden vectors in
tests/test_oracle.py (fixtures extracted by tests/golden/make_golden.py), and against each other on
seeded random graphs.

Citations: S/ = public/java/src/uk/ac/ox/well/cortexjdk/ inside the reference checkout.
"""
from __future__ import annotations

import struct

import numpy as np

ERROR_RATE_BYTES = bytes([0, 0xD8, 0xA3, 0x70, 0x3D, 0x0A, 0xD7, 0xA3, 0xF8, 0x3F, 0, 0, 0, 0, 0, 0])  # CortexGraphWriter.java:76


class CortexFormatError(Exception):
    """Stands in for CortexJDKException (S/utils/exceptions/CortexJDKException.java)."""


# ----------------------------------------------------------------------------- header

def parse_header(buf: bytes) -> dict:
    """S/utils/io/graph/cortex/CortexGraph.java:66-149 (field order) and docs/ctx_spec.md tables 1-3."""
    if len(buf) < 22 or buf[:6].upper() != b"CORTEX":
        raise CortexFormatError("does not appear to be a Cortex graph")        # :74-76
    version, k, s, c = struct.unpack_from("<4I", buf, 6)
    if version != 6:
        raise CortexFormatError("not a version 6 Cortex graph")                # :82-84
    p = 22
    mean_read_len = list(struct.unpack_from("<%dI" % c, buf, p)); p += 4 * c
    total_seq = list(struct.unpack_from("<%dQ" % c, buf, p)); p += 8 * c        # (LE per spec; Java mis-reads BE, unused)
    names = []
    for _ in range(c):
        (L,) = struct.unpack_from("<I", buf, p); p += 4
        raw = buf[p:p + L]; p += L
        nul = raw.find(b"\0")                                                  # fixStringsWithEarlyTerminators :50-64
        names.append((raw if nul < 0 else raw[:nul]).decode("latin-1"))
    p += 16 * c                                                                # error rates skipped :114-117
    colors = []
    for i in range(c):
        tip, sup, kmr, cleaned = struct.unpack_from("<4B", buf, p); p += 4
        sup_t, kmer_t, G = struct.unpack_from("<3I", buf, p); p += 12
        raw = buf[p:p + G]; p += G
        nul = raw.find(b"\0")
        colors.append(dict(sample_name=names[i], mean_read_length=mean_read_len[i], total_sequence=total_seq[i],
                           tip_clipping=bool(tip), low_covg_supernodes_removed=bool(sup),
                           low_covg_kmers_removed=bool(kmr), cleaned_against_graph=bool(cleaned),
                           low_cov_supernodes_threshold=sup_t, low_cov_kmer_threshold=kmer_t,
                           cleaned_against_graph_name=(raw if nul < 0 else raw[:nul]).decode("latin-1")))
    if buf[p:p + 6].upper() != b"CORTEX":
        raise CortexFormatError("no proper header terminator")                 # :140-142
    p += 6
    rec = 8 * s + 5 * c                                                        # :148
    n = (len(buf) - p) // rec if rec else 0                                    # :149
    return dict(version=version, kmer_size=k, kmer_bits=s, num_colors=c, colors=colors,
                data_offset=p, record_size=rec, num_records=n)


def write_header(k: int, s: int, colors: list[dict]) -> bytes:
    """CortexGraphWriter.initialize, S/utils/io/graph/cortex/CortexGraphWriter.java:45-94."""
    c = len(colors)
    out = [b"CORTEX", struct.pack("<4I", 6, k, s, c)]
    out.append(b"".join(struct.pack("<I", col.get("mean_read_length", 0)) for col in colors))
    out.append(b"".join(struct.pack("<Q", col.get("total_sequence", 0)) for col in colors))
    for col in colors:
        nm = col["sample_name"].encode("latin-1")
        out.append(struct.pack("<I", len(nm)) + nm)
    out.append(ERROR_RATE_BYTES * c)
    for col in colors:
        g = col.get("cleaned_against_graph_name", "").encode("latin-1")
        out.append(struct.pack("<4B", int(col.get("tip_clipping", False)), int(col.get("low_covg_supernodes_removed", False)),
                               int(col.get("low_covg_kmers_removed", False)), int(col.get("cleaned_against_graph", False))))
        out.append(struct.pack("<3I", col.get("low_cov_supernodes_threshold", 0), col.get("low_cov_kmer_threshold", 0), len(g)) + g)
    out.append(b"CORTEX")
    return b"".join(out)


def roi_header(k: int, s: int, child_name: str) -> bytes:
    """FindROIs.makeCortexHeader, S/commands/discover/roi/FindROIs.java:85-105."""
    return write_header(k, s, [dict(sample_name=child_name, cleaned_against_graph_name="")])


# ----------------------------------------------------------------------------- records

def record_dtype(s: int, c: int) -> np.dtype:
    """ctx_spec.md table 5: s LE uint64 words (word 0 most significant), c LE uint32, c uint8; packed."""
    return np.dtype([("kmer", "<u8", (s,)), ("cov", "<u4", (c,)), ("edges", "u1", (c,))])


def records_view(buf: bytes, hdr: dict) -> np.ndarray:
    dt = record_dtype(hdr["kmer_bits"], hdr["num_colors"])
    assert dt.itemsize == hdr["record_size"]
    return np.frombuffer(buf, dtype=dt, count=hdr["num_records"], offset=hdr["data_offset"])


def java_binary_kmer(words: np.ndarray) -> np.ndarray:
    """The long[] a CortexRecord holds = byte-swapped disk words (CortexGraph.java:208-209)."""
    return words.astype("<u8").byteswap().view(np.int64)


def java_coverage(cov: np.ndarray) -> np.ndarray:
    """BinaryUtils.toUnsignedInt returns (int) l: wraps >= 2^31 (S/utils/io/utils/BinaryUtils.java:6-17)."""
    return cov.astype(np.uint32).view(np.int32)


_CODE_TO_CHAR = np.frombuffer(b"ACGT", dtype=np.uint8)


def decode_kmers(words: np.ndarray, k: int) -> np.ndarray:
    """CortexRecord.decodeBinaryKmer :291-307, vectorised: native words [N,s] -> ASCII [N,k]."""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    words = words.reshape(len(words), words.shape[-1] if words.ndim > 1 else 1)
    n, s = words.shape
    out = np.empty((n, k), dtype=np.uint8)
    for i in range(k):                      # base i counted from the left; base k-1 is in the lowest 2 bits
        bit = 2 * (k - 1 - i)
        w = s - 1 - bit // 64
        out[:, i] = _CODE_TO_CHAR[((words[:, w] >> np.uint64(bit % 64)) & np.uint64(3)).astype(np.intp)]
    return out


_CHAR_TO_CODE = np.full(256, 255, dtype=np.uint8)
for _ch, _v in ((b"A", 0), (b"C", 1), (b"G", 2), (b"T", 3), (b"a", 0), (b"c", 1), (b"g", 2), (b"t", 3)):
    _CHAR_TO_CODE[_ch[0]] = _v                # charToBinaryNucleotide, CortexRecord.java:347-360


def encode_kmers(ascii_kmers: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """CortexRecord.encodeBinaryKmer :313-334, vectorised.  Returns (native words [N,s], ok[N]);
    rows with a byte outside ACGTacgt (Java throws) get ok=False and zero words."""
    a = np.ascontiguousarray(ascii_kmers, dtype=np.uint8)
    n, k = a.shape
    s = (k + 31) // 32
    codes = _CHAR_TO_CODE[a]
    ok = (codes != 255).all(axis=1)
    codes = np.where(codes == 255, 0, codes).astype(np.uint64)
    words = np.zeros((n, s), dtype=np.uint64)
    for i in range(k):
        bit = 2 * (k - 1 - i)
        words[:, s - 1 - bit // 64] |= codes[:, i] << np.uint64(bit % 64)
    words[~ok] = 0
    return words, ok


_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn.", b"TGCAtgcaNn."):
    _COMP[_a] = _b                            # SequenceUtils.complement :61-86 (default: itself)


def reverse_complement(ascii_rows: np.ndarray) -> np.ndarray:
    """SequenceUtils.reverseComplement :127-135 over rows."""
    return _COMP[np.ascontiguousarray(ascii_rows, dtype=np.uint8)[..., ::-1]]


def lowest_orientation(ascii_rows: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """SequenceUtils.alphanumericallyLowestOrientation :206-225 over rows: first position where
    seq[i] != comp(seq[n-1-i]) decides (signed bytes); all equal -> forward."""
    a = np.ascontiguousarray(ascii_rows, dtype=np.uint8)
    rc = reverse_complement(a)
    diff = a != rc
    has = diff.any(axis=1)
    first = diff.argmax(axis=1)
    rows = np.arange(len(a))
    flipped = has & (a[rows, first].view(np.int8) > rc[rows, first].view(np.int8))
    return np.where(flipped[:, None], rc, a), flipped


def windows(seq: bytes | np.ndarray, k: int) -> np.ndarray:
    """All k-windows of seq as rows (Call.loadChildWalk substring loop, Call.java:2364-2365)."""
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
    if len(a) < k:
        return np.empty((0, k), dtype=np.uint8)
    return np.lib.stride_tricks.sliding_window_view(a, k)


def pack_windows(seq: bytes | np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """K3 oracle: canonical native words [W,s] and flags (bit0 flipped, bit1 not ACGTacgt, bit2 has lowercase)."""
    w = windows(seq, k)
    canon, flipped = lowest_orientation(w)
    words, ok = encode_kmers(canon)
    lower = ((canon >= ord("a")) & (canon <= ord("z"))).any(axis=1)
    flags = np.where(ok, flipped.astype(np.uint8) | (lower.astype(np.uint8) << 2), 2).astype(np.uint8)
    return words, flags


# ----------------------------------------------------------------------------- novelty (FindROIs)

def is_novel(cov: np.ndarray, child: int, parents: list[int]) -> np.ndarray:
    """FindROIs.isNovel, S/commands/discover/roi/FindROIs.java:72-82 (signed ints)."""
    jc = java_coverage(cov)
    lack = np.ones(len(jc), dtype=bool)
    for p in parents:
        lack &= jc[:, p] == 0
    return (jc[:, child] > 0) & lack


def find_rois(buf: bytes, child: int, parents: list[int]) -> tuple[bytes, np.ndarray]:
    """FindROIs.execute :52-67 + CortexGraphWriter.addRecord: body bytes of the ROI graph (records of
    8s+5 bytes, input order) and the input indices of the novel records."""
    hdr = parse_header(buf)
    rec = records_view(buf, hdr)
    keep = np.flatnonzero(is_novel(rec["cov"], child, parents))
    out = np.zeros(len(keep), dtype=record_dtype(hdr["kmer_bits"], 1))
    out["kmer"] = rec["kmer"][keep]
    out["cov"][:, 0] = rec["cov"][keep, child]
    out["edges"][:, 0] = rec["edges"][keep, child]
    return out.tobytes(), keep.astype(np.uint64)


# ----------------------------------------------------------------------------- lookups

def _keys_as_bytes(words: np.ndarray) -> np.ndarray:
    """[N,s] native words -> fixed-width big-endian byte strings whose memcmp order == k-mer order."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    w = w.reshape(len(w), w.shape[-1] if w.ndim > 1 else 1)
    s = w.shape[1]
    be = w.astype(">u8")
    return np.ascontiguousarray(be).view(np.dtype(("V", 8 * s))).reshape(len(w))


def find_packed(table_words: np.ndarray, query_words: np.ndarray) -> np.ndarray:
    """Exact-match index of each canonical packed query in the sorted table, -1 for a miss."""
    tb = _keys_as_bytes(table_words)
    qb = _keys_as_bytes(query_words)
    # np.void has no ordering; compare through big-endian byte columns with lexsort-free searchsorted
    # on a structured (hi.., lo) unsigned view instead.
    s = tb.dtype.itemsize // 8
    dt = np.dtype([("w%d" % i, ">u8") for i in range(s)])
    t = tb.view(dt)
    q = qb.view(dt)
    pos = np.searchsorted(t, q, side="left")
    posc = np.minimum(pos, max(len(t) - 1, 0))
    hit = (pos < len(t)) & (t[posc] == q) if len(t) else np.zeros(len(q), dtype=bool)
    return np.where(hit, pos, -1).astype(np.int64)


def find_batch(buf: bytes, ascii_queries: np.ndarray) -> np.ndarray:
    """CortexGraph.findRecord :272-317 for N >= 3 sorted, duplicate-free graphs: canonicalise (ASCII),
    then equality against decoded uppercase ACGT record k-mers, so any query with a byte outside ACGT misses."""
    hdr = parse_header(buf)
    rec = records_view(buf, hdr)
    q = np.ascontiguousarray(ascii_queries, dtype=np.uint8).reshape(-1, hdr["kmer_size"])
    canon, _ = lowest_orientation(q)
    upper_acgt = np.isin(canon, np.frombuffer(b"ACGT", dtype=np.uint8)).all(axis=1)
    words, _ = encode_kmers(canon)
    idx = find_packed(rec["kmer"], words)
    return np.where(upper_acgt, idx, -1)


def find_record_faithful(buf: bytes, query: bytes, cached: set[int] | None = None) -> int | None:
    """Line-by-line CortexGraph.findRecord :272-317 (pure Python; small graphs only).
    `cached` = record indices present in the LRU (record 0 after construction, :162).  Returns the
    index, None for null; raises CortexFormatError where the reference throws."""
    hdr = parse_header(buf)
    rec = records_view(buf, hdr)
    k = hdr["kmer_size"]
    kmers = decode_kmers(rec["kmer"], k)
    canon, _ = lowest_orientation(np.frombuffer(query, dtype=np.uint8)[None, :])
    q = canon[0].view(np.int8)
    if cached is None:
        cached = {0} if hdr["num_records"] else set()
    for i in cached:                                                  # :274-276
        if i < hdr["num_records"] and bytes(kmers[i]) == bytes(canon[0]):
            return i

    def cmp(a, b):                                                    # CortexByteKmer.compareTo :41-49
        for x, y in zip(a.tolist(), b.tolist()):
            if x < y:
                return -1
            if x > y:
                return 1
        return 0

    start, stop = 0, hdr["num_records"] - 1
    mid = start + int((stop - start) / 2)                             # Java '/' truncates toward zero
    while start != mid and mid != stop:
        a, m, z = kmers[start].view(np.int8), kmers[mid].view(np.int8), kmers[stop].view(np.int8)
        if cmp(a, z) > 0 or cmp(a, m) > 0:
            raise CortexFormatError("Records are not sorted")         # :295-301
        if cmp(q, z) > 0 or cmp(q, a) < 0:
            return None
        if cmp(a, q) == 0:
            return start
        if cmp(m, q) == 0:
            return mid
        if cmp(z, q) == 0:
            return stop
        if cmp(q, a) > 0 and cmp(q, m) < 0:
            stop = mid
            mid = start + int((stop - start) / 2)
        elif cmp(q, m) > 0 and cmp(q, z) < 0:
            start = mid
            mid = start + int((stop - start) / 2)
    return None


# ----------------------------------------------------------------------------- text forms

def edges_to_string(edge: int) -> str:
    """CortexRecord.getEdgesAsBytes :117-140."""
    s = "acgtACGT"
    left, right = (edge >> 4) & 0xF, edge & 0xF
    out = ["."] * 8
    for i in range(4):
        if left & (1 << (3 - i)):
            out[i] = s[i]
        if right & (1 << i):
            out[i + 4] = s[i + 4]
    return "".join(out)


def edges_from_string(text: str) -> int:
    """Inverse of edges_to_string (CortexRecord.encodeBinaryEdges :379-408 for the non-flipped case)."""
    v = 0
    for i in range(4):
        if text[i] != ".":
            v |= 1 << (7 - i)
        if text[i + 4] != ".":
            v |= 1 << i
    return v


def record_to_string(kmer_ascii: np.ndarray, cov: np.ndarray, edges: np.ndarray) -> str:
    """CortexRecord.toString :166-178."""
    parts = [bytes(kmer_ascii).decode()] + [str(int(v)) for v in java_coverage(np.asarray(cov))]
    parts += [edges_to_string(int(e)) for e in edges]
    return " ".join(parts)


# ----------------------------------------------------------------------------- TempGraphAssembler

def temp_graph_assembler(haplotypes: list[tuple[str, list[str]]], k: int) -> bytes:
    """S/utils/assembler/TempGraphAssembler.java:19-127: a sorted multi-colour .ctx from haplotype strings
    (coverage = occurrences, edges from neighbouring bases, swapped + complemented when the canonical
    orientation is the reverse complement).  Pure Python; test-sized inputs only."""
    nc = len(haplotypes)
    table: dict[bytes, tuple[list[int], list[set], list[set]]] = {}
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for color, (_, seqs) in enumerate(haplotypes):
        for seq in seqs:
            seq = seq.upper()
            for i in range(len(seq) - k + 1):
                sk = seq[i:i + k].encode()
                prev_b = None if i == 0 else seq[i - 1]
                next_b = None if i == len(seq) - k else seq[i + k]
                canon, fl = lowest_orientation(np.frombuffer(sk, dtype=np.uint8)[None, :])
                key = bytes(canon[0])
                cov, ins, outs = table.setdefault(key, ([0] * nc, [set() for _ in range(nc)], [set() for _ in range(nc)]))
                cov[color] += 1
                if not fl[0]:
                    if prev_b: ins[color].add(prev_b)
                    if next_b: outs[color].add(next_b)
                else:
                    if next_b: ins[color].add(comp[next_b])
                    if prev_b: outs[color].add(comp[prev_b])
    s = (k + 31) // 32
    colors = [dict(sample_name=name, cleaned_against_graph_name="") for name, _ in haplotypes]
    body = np.zeros(len(table), dtype=record_dtype(s, nc))
    for r, key in enumerate(sorted(table)):                               # TreeMap<CanonicalKmer,...> order :31,:51
        cov, ins, outs = table[key]
        body["kmer"][r], _ = encode_kmers(np.frombuffer(key, dtype=np.uint8)[None, :])
        body["cov"][r] = cov
        for c in range(nc):
            e = 0
            for i, b in enumerate("ACGT"):
                if b in ins[c]: e |= 1 << (7 - i)
                if b in outs[c]: e |= 1 << i
            body["edges"][r, c] = e
    return write_header(k, s, colors) + body.tobytes()


# ----------------------------------------------------------------------------- Join / CortexCollection

def join(bufs: list[bytes]) -> bytes:
    """S/commands/utils/Join.java:23-57 over S/utils/io/graph/cortex/CortexCollection.java:34-62,245-293:
    the ascending union of the graphs' k-mers, colours concatenated in argument order, coverage 0 / no edges in the
    colours of a graph that lacks the k-mer; header written by CortexGraphWriter from the input colours -- with
    total_sequence byte-reversed, because the reference reads that field big-endian (BinaryFile.java:34-38) and
    writes it little-endian (CortexGraphWriter.java:60-63)."""
    hdrs = [parse_header(b) for b in bufs]
    k, s = hdrs[0]["kmer_size"], hdrs[0]["kmer_bits"]
    for h in hdrs:
        if h["kmer_size"] != k:
            raise CortexFormatError("Graph kmer sizes are not equal")          # CortexCollection.java:43-45
    recs = [records_view(b, h) for b, h in zip(bufs, hdrs)]
    be = [np.ascontiguousarray(r["kmer"]).astype(">u8").view(np.dtype((np.void, 8 * s))).reshape(-1) for r in recs]
    allk = np.unique(np.concatenate(be)) if sum(len(x) for x in be) else np.zeros(0, dtype=be[0].dtype)
    ctot = sum(h["num_colors"] for h in hdrs)
    out = np.zeros(len(allk), dtype=record_dtype(s, ctot))
    out["kmer"] = allk.view(">u8").reshape(-1, s).astype("<u8")
    c0 = 0
    for r, keys, h in zip(recs, be, hdrs):
        c = h["num_colors"]
        pos = np.searchsorted(allk, keys)
        out["cov"][pos, c0:c0 + c] = r["cov"]
        out["edges"][pos, c0:c0 + c] = r["edges"]
        c0 += c
    colors = []
    for h in hdrs:
        for col in h["colors"]:
            col = dict(col)
            col["total_sequence"] = int.from_bytes(int(col["total_sequence"]).to_bytes(8, "little"), "big")
            colors.append(col)
    return write_header(k, s, colors) + out.tobytes()


def remove(primary_buf: bytes, secondary_bufs: list[bytes]) -> tuple[bytes, int]:
    """S/commands/utils/Remove.java:30-88: walk the merged collection (join() above); `found` = coverage > 0 (Java int) in a
    colour >= PGRAPH.getNumColors(); the records that are not found are written with the primary's colours under the primary's
    header (re-emitted by CortexGraphWriter).  Returns (file bytes, records removed)."""
    merged = join([primary_buf] + list(secondary_bufs))
    hm, hp = parse_header(merged), parse_header(primary_buf)
    rec = records_view(merged, hm)
    cp = hp["num_colors"]
    found = (java_coverage(rec["cov"][:, cp:]) > 0).any(axis=1) if hm["num_colors"] > cp else np.zeros(len(rec), dtype=bool)
    keep = rec[~found]
    out = np.zeros(len(keep), dtype=record_dtype(hp["kmer_bits"], cp))
    out["kmer"] = keep["kmer"]
    out["cov"] = keep["cov"][:, :cp]
    out["edges"] = keep["edges"][:, :cp]
    return _rewritten_header(hp, hp["colors"]) + out.tobytes(), int(found.sum())


def sort_graph(buf: bytes) -> bytes:
    """S/commands/utils/Sort.java:19-50: Arrays.sort of the records by k-mer string (stable), same header re-emitted by
    CortexGraphWriter (total_sequence byte-reversed, see join())."""
    h = parse_header(buf)
    rec = records_view(buf, h)
    be = np.ascontiguousarray(rec["kmer"]).astype(">u8").view(np.dtype((np.void, 8 * h["kmer_bits"]))).reshape(-1)
    order = np.argsort(be, kind="stable")
    colors = []
    for col in h["colors"]:
        col = dict(col)
        col["total_sequence"] = int.from_bytes(int(col["total_sequence"]).to_bytes(8, "little"), "big")
        colors.append(col)
    return write_header(h["kmer_size"], h["kmer_bits"], colors) + rec[order].tobytes()


# ----------------------------------------------------------------------------- scan-shaped pre-filters

def _rewritten_header(h: dict, colors: list[dict]) -> bytes:
    """What CortexGraphWriter emits for colours that were READ by CortexGraph: total_sequence byte-reversed (see join())."""
    out = []
    for col in colors:
        col = dict(col)
        col["total_sequence"] = int.from_bytes(int(col["total_sequence"]).to_bytes(8, "little"), "big")
        out.append(col)
    return write_header(h["kmer_size"], h["kmer_bits"], out)


def find_low_coverage(buf: bytes, min_coverage: int) -> bytes:
    """S/commands/prefilter/FindLowCoverage.java:33-66: `if (cr.getCoverage(0) >= MIN_COVERAGE) numKept++ else cgw.addRecord(cr)`
    -- the file holds the records BELOW the limit, under the input header."""
    h = parse_header(buf)
    rec = records_view(buf, h)
    low = java_coverage(rec["cov"][:, 0]) < min_coverage
    return _rewritten_header(h, h["colors"]) + rec[low].tobytes()


def find_shared(graph_buf: bytes, roi_buf: bytes, child: int, parents: list[int], ignore: list[int]) -> bytes:
    """S/commands/prefilter/FindShared.java:40-118: a ROI record is shared when GRAPH.findRecord(its k-mer) has coverage > 0 in a
    colour c with c != childColor, c not in parentColors, c not in ignoreColors (:67); shared ROI records are written in order
    under the ROI header.  A k-mer missing from GRAPH dereferences null (:67) -> raised here as KeyError."""
    hg, hr = parse_header(graph_buf), parse_header(roi_buf)
    g, r = records_view(graph_buf, hg), records_view(roi_buf, hr)
    idx = find_packed(np.ascontiguousarray(g["kmer"]), np.ascontiguousarray(r["kmer"]))
    free = np.array([c != child and c not in set(parents) and c not in set(ignore) for c in range(hg["num_colors"])], dtype=bool)
    # the null record is dereferenced only for a colour that passes the three exclusions (short-circuit && at :67)
    if (idx < 0).any() and free.any():
        raise KeyError("ROI record %d is not in the graph" % int(np.nonzero(idx < 0)[0][0]))
    cov = java_coverage(g["cov"][np.maximum(idx, 0)]) if len(g) else np.zeros((len(r), hg["num_colors"]), dtype=np.int64)
    shared = (cov[:, free] > 0).any(axis=1) if free.any() else np.zeros(len(r), dtype=bool)
    return _rewritten_header(hr, hr["colors"]) + r[shared].tobytes()


def recover_excluded_kmers(graph_buf: bytes, dirty_buf: bytes, child: int) -> tuple[bytes, int]:
    """S/commands/discover/recover/RecoverExcludedKmers.java:31-106.  Records with coverage(child) > 0 are written as they are
    (:50-52); otherwise, if another colour has coverage (:53-61) and DIRTY.findRecord(k-mer) exists with coverage(0) > 0 (:63-65),
    a copy with coverages[child] = that coverage is written (:74,:80).  The header has ONE colour (makeHeader :98-106) and
    CortexGraphWriter.addRecord (CortexGraphWriter.java:106-138) writes header.getNumColors() = 1 coverage and edge of the
    record it is given: coverage[0] and edges[0] of the pedigree record."""
    hg, hd = parse_header(graph_buf), parse_header(dirty_buf)
    g, d = records_view(graph_buf, hg), records_view(dirty_buf, hd)
    cov = java_coverage(g["cov"]).copy()
    keep = cov[:, child] > 0
    others = np.ones(hg["num_colors"], dtype=bool); others[child] = False
    cand = ~keep & (cov[:, others] > 0).any(axis=1)
    idx = np.full(len(g), -1, dtype=np.int64)
    if cand.any():
        idx[cand] = find_packed(np.ascontiguousarray(d["kmer"]), np.ascontiguousarray(g["kmer"][cand]))
    dcov = np.where(idx >= 0, java_coverage(d["cov"][:, 0])[np.maximum(idx, 0)], 0)
    recovered = cand & (idx >= 0) & (dcov > 0)
    cov[recovered, child] = dcov[recovered]
    sel = keep | recovered
    out = np.zeros(int(sel.sum()), dtype=record_dtype(hg["kmer_bits"], 1))
    out["kmer"] = g["kmer"][sel]
    out["cov"][:, 0] = np.ascontiguousarray(cov[sel, 0]).view(np.uint32)
    out["edges"][:, 0] = g["edges"][sel, 0]
    return _rewritten_header(hg, [hg["colors"][child]]) + out.tobytes(), int(recovered.sum())


def cov_stats(buf: bytes, child: int, parents: list[int]) -> list[tuple[int, int]]:
    """S/commands/utils/CovStats.java:33-72: for every record with coverage(child) > 0, numberOfParents > 0 and
    numberOfChildren > 0 (colours with coverage > 0 that are parents / neither child nor parent, :50-58):
    hist[coverage(child)] += numberOfParents + numberOfChildren (Java int); rows printed in ascending coverage (TreeMap)."""
    h = parse_header(buf)
    cov = java_coverage(records_view(buf, h)["cov"])
    c = h["num_colors"]
    is_parent = np.array([cc in set(parents) and cc != child for cc in range(c)], dtype=bool)
    is_other = np.array([cc not in set(parents) and cc != child for cc in range(c)], dtype=bool)
    pos = cov > 0
    npar = pos[:, is_parent].sum(axis=1)
    noth = pos[:, is_other].sum(axis=1)
    counts = (cov[:, child] > 0) & (npar > 0) & (noth > 0)
    hist: dict[int, int] = {}
    for cv, w in zip(cov[counts, child].tolist(), (npar + noth)[counts].tolist()):
        hist[cv] = hist.get(cv, 0) + w
    wrap = lambda v: ((v + 2 ** 31) % 2 ** 32) - 2 ** 31
    return [(cv, wrap(hist[cv])) for cv in sorted(hist)]
