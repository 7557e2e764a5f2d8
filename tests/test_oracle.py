"""Pins BOTH CPU oracles (oracle/ctx_oracle.c and oracle/oracle_np.py) to the reference's own golden
vectors (tests/golden/, extracted from the reference's TestNG sources) and to each other.
Mirrors T/utils/kmer/CortexGraphTest.java, T/utils/sequence/SequenceUtilsTest.java,
T/utils/kmer/CortexGraphWriterTest.java and T/utils/traversal/TraversalEngineTest.java:48-95."""
import random

import numpy as np
import pytest

from oracle import oracle_np as onp
from oracle import orc
from tools import synth
import torch


# ---------------------------------------------------------------- header / numRecordsTest / getShortSampleNames

def test_header_fixture(fixture_ctx):
    h = onp.parse_header(fixture_ctx)
    assert (h["version"], h["kmer_size"], h["kmer_bits"], h["num_colors"]) == (6, 31, 1, 2)
    assert h["data_offset"] == 148 and h["record_size"] == 18
    assert h["num_records"] == 66                                        # CortexGraphTest.numRecordsTest :147-152
    assert [c["sample_name"] for c in h["colors"]] == ["one", "two"]     # getShortSampleNames :139-145
    g = orc.Graph(fixture_ctx)
    assert g.ok and g.h.num_records == 66 and g.h.data_offset == 148 and g.h.record_size == 18
    assert [g.color_name(0), g.color_name(1)] == ["one", "two"]
    assert g.color_for_sample_name("ONE") == 0 and g.color_for_sample_name("two") == 1
    assert g.color_for_sample_name("1") == 1 and g.color_for_sample_name("nope") == -1


# ---------------------------------------------------------------- recordsAreCorrect :186-198

def test_records_match_golden_table(fixture_ctx, kats):
    h = onp.parse_header(fixture_ctx)
    rec = onp.records_view(fixture_ctx, h)
    kmers = onp.decode_kmers(rec["kmer"], h["kmer_size"])
    g = orc.Graph(fixture_ctx)
    for i, row in enumerate(kats["fixture_records"]):
        expect = "%s %d %d %s %s" % (row["kmer"], *row["coverage"], *row["edges"])
        assert onp.record_to_string(kmers[i], rec["cov"][i], rec["edges"][i]) == expect
        bk, cov, ed = g.get_record(i)
        assert g.kmer_string(bk) == row["kmer"].encode()
        assert cov.tolist() == row["coverage"]
        assert [onp.edges_to_string(int(e)) for e in ed] == row["edges"]
        # the C oracle holds the Java long[] convention (byte-swapped disk word)
        assert bk.tolist() == onp.java_binary_kmer(rec["kmer"][i]).tolist()
    assert g.get_record(66) is None                                      # i >= N -> null (CortexGraph.java:190,236)
    # strictly ascending in file order
    ks = [bytes(k) for k in kmers]
    assert ks == sorted(ks) and len(set(ks)) == 66


def test_get_record_backwards(fixture_ctx, kats):                        # testGetRecord :255-265
    g = orc.Graph(fixture_ctx)
    for i in range(10, -1, -1):
        bk, cov, _ = g.get_record(i)
        assert g.kmer_string(bk).decode() == kats["fixture_records"][i]["kmer"]


def test_encode_binary_kmer(fixture_ctx):                                # testEncodeBinaryKmer :267-280
    g = orc.Graph(fixture_ctx)
    h = onp.parse_header(fixture_ctx)
    rec = onp.records_view(fixture_ctx, h)
    for i in range(10, -1, -1):
        bk, _, _ = g.get_record(i)
        ks = np.frombuffer(g.kmer_string(bk), dtype=np.uint8)
        enc = np.zeros(1, dtype=np.int64)
        assert orc.lib().orc_encode_binary_kmer(ks.ctypes.data, 31, enc.ctypes.data) == 0
        assert enc.tolist() == bk.tolist()
        w, ok = onp.encode_kmers(ks[None, :])
        assert ok[0] and w[0, 0] == rec["kmer"][i][0]
    bad = np.frombuffer(b"ACGTN", dtype=np.uint8)
    assert orc.lib().orc_encode_binary_kmer(bad.ctypes.data, 5, np.zeros(1, dtype=np.int64).ctypes.data) == -1


# ---------------------------------------------------------------- findRecord :310-331

def test_sorted_find_record(fixture_ctx, kats):
    g = orc.Graph(fixture_ctx)
    qs = np.array([list(r["kmer"].encode()) for r in kats["fixture_records"]], dtype=np.uint8)
    assert g.find_batch(qs).tolist() == list(range(66))
    assert onp.find_batch(fixture_ctx, qs).tolist() == list(range(66))
    rc = onp.reverse_complement(qs)                                      # queries given on the other strand
    assert g.find_batch(rc).tolist() == list(range(66))
    assert onp.find_batch(fixture_ctx, rc).tolist() == list(range(66))
    for i in (0, 1, 32, 33, 64, 65):
        assert onp.find_record_faithful(fixture_ctx, bytes(qs[i])) == i


def test_find_non_existent_record(fixture_ctx, kats):
    g = orc.Graph(fixture_ctx)
    q = kats["missing_query"].encode()
    assert g.find_record(q) == -1
    assert onp.find_batch(fixture_ctx, np.frombuffer(q, dtype=np.uint8)[None, :]).tolist() == [-1]
    assert onp.find_record_faithful(fixture_ctx, q) is None
    low = kats["fixture_records"][5]["kmer"].lower().encode()            # lowercase never equals a decoded record
    assert g.find_record(low) == -1
    assert onp.find_batch(fixture_ctx, np.frombuffer(low, dtype=np.uint8)[None, :]).tolist() == [-1]


def test_all_fasta_windows_hit(fixture_ctx, fixture_fa):                 # BASELINE.json configs[0]
    g = orc.Graph(fixture_ctx)
    seen = set()
    for seq in fixture_fa:
        idx = g.find_windows(seq)
        assert (idx >= 0).all()
        assert idx.tolist() == onp.find_batch(fixture_ctx, onp.windows(seq, 31)).tolist()
        seen.update(idx.tolist())
    assert seen == set(range(66))


# ---------------------------------------------------------------- SequenceUtilsTest :19-72

def test_complement_table(kats):
    for a, b in zip(kats["complement_in"], kats["complement_out"]):
        assert orc.lib().orc_complement(ord(a)) == ord(b)
        assert onp.reverse_complement(np.array([[ord(a)]], dtype=np.uint8))[0, 0] == ord(b)
    assert orc.lib().orc_complement(ord("n")) == ord("n") and orc.lib().orc_complement(ord("X")) == ord("X")


def test_reverse_complement_vectors(kats):
    for seq, exp in kats["reverse_complement"]:
        a = np.frombuffer(seq.encode(), dtype=np.uint8)
        out = np.empty_like(a)
        orc.lib().orc_reverse_complement(a.ctypes.data, len(a), out.ctypes.data)
        assert out.tobytes().decode() == exp
        assert onp.reverse_complement(a[None, :])[0].tobytes().decode() == exp


def test_lowest_orientation_property():
    """alphanumericallyLowestOrientation == min(fw, rc) by String.compareTo, 40 000 random k in {21,31,41,51}."""
    rng = random.Random(0)
    comp = str.maketrans("ACGT", "TGCA")
    for k in (21, 31, 41, 51):
        fws = ["".join(rng.choice("ACGT") for _ in range(k)) for _ in range(10000)]
        arr = np.array([list(f.encode()) for f in fws], dtype=np.uint8)
        canon, flipped = onp.lowest_orientation(arr)
        for j, fw in enumerate(fws):
            rc = fw.translate(comp)[::-1]
            exp = fw if fw < rc else rc
            assert canon[j].tobytes().decode() == exp
            if j < 500:
                got, fl = orc.lowest_orientation(fw.encode())
                assert got.decode() == exp and fl == (exp != fw)
                assert fl == bool(flipped[j])


def test_lowest_orientation_non_acgt():
    for q in (b"NACGT", b"TTTTN", b"acgTT", b"AC.GT", b"AAAAA", b"ACGT", bytes([200, 65, 67, 71, 84])):
        got, fl = orc.lowest_orientation(q)
        canon, flipped = onp.lowest_orientation(np.frombuffer(q, dtype=np.uint8)[None, :])
        assert got == canon[0].tobytes() and fl == bool(flipped[0])


def test_hash_collision_pair_distinct(kats):                             # CanonicalKmerTest :8-14
    a, b = (x.encode() for x in kats["hash_collision_pair"])

    def jhash(bs):                                                       # java.util.Arrays.hashCode(byte[])
        h = 1
        for x in bs:
            h = (31 * h + (x - 256 if x > 127 else x)) & 0xFFFFFFFF
        return h
    ca, _ = orc.lowest_orientation(a)
    cb, _ = orc.lowest_orientation(b)
    assert jhash(ca) == jhash(cb) and ca != cb


# ---------------------------------------------------------------- TempGraphAssembler KATs (TraversalEngineTest :48-95)

def test_assembler_kats(kats):
    for blk in kats["assembler_kats"]:
        haps = [(name, [seq]) for name, seq in blk["haplotypes"]]
        ctx = onp.temp_graph_assembler(haps, blk["k"])
        h = onp.parse_header(ctx)
        assert h["num_records"] == len(blk["records"])
        rec = onp.records_view(ctx, h)
        kmers = onp.decode_kmers(rec["kmer"], blk["k"])
        got = {onp.record_to_string(kmers[i], rec["cov"][i], rec["edges"][i]) for i in range(len(rec))}
        assert got == set(blk["records"])
        g = orc.Graph(ctx)                                               # the C oracle reads what the writer wrote
        assert g.ok and g.h.num_records == len(blk["records"])
        for i in range(len(rec)):
            bk, cov, ed = g.get_record(i)
            text = " ".join([g.kmer_string(bk).decode()] + [str(v) for v in cov] + [onp.edges_to_string(int(e)) for e in ed])
            assert text in blk["records"]


# ---------------------------------------------------------------- FindROIs (no reference test; pinned by source :72-82)

def test_novelty_fixture(fixture_ctx):
    g = orc.Graph(fixture_ctx)
    for child, parents, n in ((0, [1], 19), (1, [0], 47), (0, [], 19), (0, [0], 0), (0, [1, 1], 19)):
        out_c, idx_c = g.find_rois(child, parents)
        out_n, idx_n = onp.find_rois(fixture_ctx, child, parents)
        assert len(idx_c) == n and idx_c.tolist() == idx_n.tolist() and out_c == out_n
        out_f, idx_f = g.find_rois(child, parents, faithful=True)
        assert out_f == out_c and idx_f.tolist() == idx_c.tolist()
    # the ROI file FindROIs writes is itself a valid, sorted 1-colour graph holding exactly those k-mers
    body, idx = g.find_rois(0, [1])
    roi = orc.roi_header(31, 1, "one") + body
    assert roi[:len(onp.roi_header(31, 1, "one"))] == onp.roi_header(31, 1, "one") and len(onp.roi_header(31, 1, "one")) == 79
    rg = orc.Graph(roi)
    assert rg.ok and rg.h.num_colors == 1 and rg.h.num_records == 19 and rg.color_name(0) == "one"
    h = onp.parse_header(fixture_ctx)
    rec = onp.records_view(fixture_ctx, h)
    for j, i in enumerate(idx.tolist()):
        bk, cov, ed = rg.get_record(j)
        assert bk.tolist() == onp.java_binary_kmer(rec["kmer"][i]).tolist()
        assert cov[0] == rec["cov"][i][0] and ed[0] == rec["edges"][i][0]


def test_novelty_signed_coverage():
    """Coverage >= 2^31 is present-but-not-positive."""
    lib = orc.lib()
    def nov(cov, child, parents):
        c = np.array(cov, dtype=np.uint32).view(np.int32); p = np.array(parents, dtype=np.int32)
        return bool(lib.orc_is_novel(c.ctypes.data, p.ctypes.data, len(p), child))
    assert nov([5, 0, 0], 0, [1, 2])
    assert not nov([0x80000000, 0, 0], 0, [1, 2])
    assert not nov([0xFFFFFFFF, 0, 0], 0, [1, 2])
    assert not nov([5, 0x80000000, 0], 0, [1, 2])
    assert nov([0x7FFFFFFF, 0, 9], 0, [1])
    cov = np.array([[5, 0, 0], [0x80000000, 0, 0], [0xFFFFFFFF, 0, 0], [5, 0x80000000, 0], [0x7FFFFFFF, 0, 0]], dtype=np.uint32)
    assert onp.is_novel(cov, 0, [1, 2]).tolist() == [True, False, False, False, True]


# ---------------------------------------------------------------- both oracles agree on seeded random graphs

@pytest.mark.parametrize("k,c,n", [(31, 4, 5000), (47, 4, 5000), (63, 21, 2000), (5, 3, 100), (95, 2, 1500), (33, 1, 700)])
def test_oracles_agree_on_synthetic(k, c, n):
    ctx = synth.make_ctx_file(1234 + k, n, k, c, novel_permille=20, adv_period=97, trailing=b"\x01\x02\x03")
    h = onp.parse_header(ctx)
    assert h["num_records"] == n and h["record_size"] == 8 * ((k + 31) // 32) + 5 * c
    g = orc.Graph(ctx)
    assert g.ok and g.h.num_records == n
    rec = onp.records_view(ctx, h)
    kmers = onp.decode_kmers(rec["kmer"], k)
    ks = [bytes(x) for x in kmers]
    assert ks == sorted(ks) and len(set(ks)) == n
    rcs = onp.reverse_complement(kmers)
    assert all(bytes(a) <= bytes(b) for a, b in zip(kmers[:200], rcs[:200]))          # records are canonical
    for i in (0, n // 2, n - 1):
        bk, cov, ed = g.get_record(i)
        assert g.kmer_string(bk) == ks[i] and cov.tolist() == onp.java_coverage(rec["cov"][i]).tolist()
    parents = list(range(1, c))
    out_c, idx_c = g.find_rois(0, parents)
    out_n, idx_n = onp.find_rois(ctx, 0, parents)
    assert out_c == out_n and idx_c.tolist() == idx_n.tolist()
    if c > 1:
        assert 0 < len(idx_c) < n
    # lookups: table hits on both strands, misses, N and lowercase
    tw = [np_to_t(rec["kmer"][:, w]) for w in range(h["kmer_bits"])]
    q_ascii, _, valid = synth.make_queries(99, tw, k, 3000, corrupt_permille=30)
    qa = q_ascii.numpy()
    a = g.find_batch(qa)
    b = onp.find_batch(ctx, qa)
    assert a.tolist() == b.tolist()
    assert (a[~valid.numpy()] == -1).all() and (a >= 0).sum() > 500 and (a == -1).sum() > 500
    # pack/canonicalise oracle pair on a genome with N's and a lowercase run
    seq = synth.random_genome(5 + k, 2000, n_permille=3).numpy().copy()
    seq[100:140] += 32
    w1, f1 = orc.pack_windows(seq, k)
    w2, f2 = onp.pack_windows(seq, k)
    assert (w1 == w2).all() and (f1 == f2).all()


def np_to_t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy())


def test_small_graph_quirk():
    """N <= 2: the search loop never runs; only the LRU (record 0 after construction) can answer."""
    for n in (0, 1, 2, 3):
        ctx = synth.make_ctx_file(7, n, 11, 2, adv_period=0) if n else synth.header_bytes(11, 2)
        g = orc.Graph(ctx)
        assert g.ok and g.h.num_records == n
        h = onp.parse_header(ctx)
        kmers = onp.decode_kmers(onp.records_view(ctx, h)["kmer"], 11)
        got = [g.find_record(bytes(kmers[i])) for i in range(n)]
        exp = {0: [], 1: [0], 2: [0, -1], 3: [0, 1, 2]}[n]
        assert got == exp
        assert [onp.find_record_faithful(ctx, bytes(kmers[i])) for i in range(n)] == [None if e < 0 else e for e in exp]
        assert g.find_record(b"A" * 11) in (-1, 0)


def test_unsorted_detected_by_probe():
    ctx = bytearray(synth.make_ctx_file(3, 9, 15, 1, adv_period=0))
    h = onp.parse_header(bytes(ctx))
    S, off = h["record_size"], h["data_offset"]
    first, last = bytes(ctx[off:off + S]), bytes(ctx[off + 8 * S:off + 9 * S])
    ctx[off:off + S], ctx[off + 8 * S:off + 9 * S] = last, first          # start > stop
    g = orc.Graph(bytes(ctx))
    assert g.find_record(b"ACGTACGTACGTACG") == -2
    with pytest.raises(onp.CortexFormatError):
        onp.find_record_faithful(bytes(ctx), b"ACGTACGTACGTACG", cached=set())


def test_bad_headers(fixture_ctx):
    assert orc.Graph(b"NOTCTX" + fixture_ctx[6:]).rc == 1
    assert orc.Graph(fixture_ctx[:6] + b"\x05\0\0\0" + fixture_ctx[10:]).rc == 2
    assert orc.Graph(fixture_ctx[:142] + b"XORTEX" + fixture_ctx[148:]).rc == 3
    assert orc.Graph(b"cortex" + fixture_ctx[6:]).rc == 0                  # equalsIgnoreCase
    with pytest.raises(onp.CortexFormatError):
        onp.parse_header(b"NOTCTX" + fixture_ctx[6:])


def test_writer_round_trip():                                            # CortexGraphWriterTest :19-78
    rng = random.Random(1)
    haps = [(nm, ["".join(rng.choice("ACGT") for _ in range(100))]) for nm in ("mom", "dad", "kid")]
    ctx = onp.temp_graph_assembler(haps, 5)
    h = onp.parse_header(ctx)
    rec = onp.records_view(ctx, h)
    again = onp.write_header(h["kmer_size"], h["kmer_bits"], h["colors"]) + rec.tobytes()
    assert again == ctx
    assert [c["sample_name"] for c in h["colors"]] == ["mom", "dad", "kid"]


# ---------------------------------------------------------------- Join / CortexCollection (CortexCollectionTest :34-111)

def test_join_oracle_matches_assembler():
    """Joining per-sample graphs gives the multi-colour graph TempGraphAssembler builds from the same haplotypes
    (coverage = occurrence count per sample), which is how CortexCollectionTest checks merged iteration."""
    rng = random.Random(5)
    haps = [(nm, ["".join(rng.choice("ACGT") for _ in range(300))]) for nm in ("mom", "dad", "kid")]
    whole = onp.temp_graph_assembler(haps, 7)
    parts = [onp.temp_graph_assembler([h], 7) for h in haps]
    joined = onp.join(parts)
    hw, hj = onp.parse_header(whole), onp.parse_header(joined)
    assert hj["num_colors"] == 3 and [c["sample_name"] for c in hj["colors"]] == ["mom", "dad", "kid"]
    rw, rj = onp.records_view(whole, hw), onp.records_view(joined, hj)
    assert hw["num_records"] == hj["num_records"]
    assert (rw["kmer"] == rj["kmer"]).all() and (rw["cov"] == rj["cov"]).all()
    # edges of a k-mer inside one sample's graph are the union over that sample only: identical to the assembler's per-colour edges
    assert (rw["edges"] == rj["edges"]).all()
    one = onp.join([parts[0]])
    assert onp.records_view(one, onp.parse_header(one)).tobytes() == onp.records_view(parts[0], onp.parse_header(parts[0])).tobytes()


# ------------------------------------------------------------------ scan-shaped pre-filters: hand-checked cases

def _tiny_graph(kmers, covs, names, k=5):
    """A .ctx image from ASCII k-mers (already canonical, ascending) and per-colour coverages; edges = colour index + 1."""
    s = 1
    words, _ = onp.encode_kmers(np.array([list(x.encode()) for x in kmers], dtype=np.uint8))
    rec = np.zeros(len(kmers), dtype=onp.record_dtype(s, len(names)))
    rec["kmer"] = words.reshape(-1, s)
    rec["cov"] = np.array(covs, dtype=np.int64).astype(np.uint32).reshape(len(kmers), len(names))
    rec["edges"] = np.arange(1, len(names) + 1, dtype=np.uint8)[None, :]
    colors = [dict(sample_name=n, mean_read_length=100, total_sequence=7 + i, graph_name="undefined", tip_clipping=0,
                   low_covg_supernodes_removed=0, low_covg_kmers_removed=0, cleaned_against_graph=0,
                   low_cov_supernodes_threshold=0, low_cov_kmer_threshold=0) for i, n in enumerate(names)]
    return onp.write_header(k, s, colors) + rec.tobytes()


def test_prefilter_oracles_hand_checked():
    km = ["AAAAA", "AAACC", "ACGTA", "CCCAA", "GATTA"]
    #            kid  mom  dad  ref
    covs = [[3, 0, 0, 0],          # child only
            [0, 2, 0, 5],          # not in child; in mom and ref
            [9, 1, 0, 4],          # child + mom + ref
            [0, 0, 0, 0],          # nowhere
            [2 ** 31, 1, 1, 1]]    # child coverage wraps negative in Java
    g = _tiny_graph(km, covs, ["kid", "mom", "dad", "ref"])
    # FindLowCoverage on a one-colour ROI: keeps coverage(0) < 4, signed
    roi = _tiny_graph(km, [[3], [0], [9], [4], [2 ** 31]], ["kid"])
    low = onp.find_low_coverage(roi, 4)
    hl = onp.parse_header(low)
    assert onp.decode_kmers(onp.records_view(low, hl)["kmer"], 5).tobytes().decode() == "AAAAA" + "AAACC" + "GATTA"
    # FindShared: free colours = {ref} when child=0, parents={1,2}; shared where ref coverage > 0
    sh = onp.find_shared(g, roi, 0, [1, 2], [])
    assert onp.records_view(sh, onp.parse_header(sh))["cov"][:, 0].tolist() == [0, 9, 2 ** 31]
    assert onp.parse_header(onp.find_shared(g, roi, 0, [1, 2], [3]))["num_records"] == 0        # ref ignored: nothing is free
    with pytest.raises(KeyError):
        onp.find_shared(g, _tiny_graph(["TTTTT"], [[1]], ["kid"]), 0, [1], [])
    # ... but with EVERY colour excluded the null record is never dereferenced (short-circuit && at FindShared.java:67): nothing is shared
    absent = _tiny_graph(["AAAAC"], [[1]], ["kid"])          # canonical, and not a k-mer of g
    assert onp.parse_header(onp.find_shared(g, absent, 0, [1, 2], [3]))["num_records"] == 0
    assert orc.find_shared_mask(orc.Graph(g), orc.Graph(absent), 0, [1, 2], [3]).tolist() == [False]
    assert orc.find_shared_mask(orc.Graph(g), orc.Graph(absent), 0, [1, 2], []) is None
    # RecoverExcludedKmers: dirty graph holds AAACC (cov 6) and CCCAA (cov 8) and GATTA (cov 0)
    dirty = _tiny_graph(["AAACC", "CCCAA", "GATTA"], [[6], [8], [0]], ["kid"])
    rec, nrec = onp.recover_excluded_kmers(g, dirty, 0)
    hr = onp.parse_header(rec)
    rv = onp.records_view(rec, hr)
    assert hr["num_colors"] == 1 and hr["colors"][0]["sample_name"] == "kid" and nrec == 1
    assert onp.decode_kmers(rv["kmer"], 5).tobytes().decode() == "AAAAA" + "AAACC" + "ACGTA"     # CCCAA has no coverage anywhere; GATTA: child <= 0, dirty 0
    assert rv["cov"][:, 0].tolist() == [3, 6, 9] and rv["edges"][:, 0].tolist() == [1, 1, 1]
    rec1, n1 = onp.recover_excluded_kmers(g, _tiny_graph(["AAAAA"], [[5]], ["mom"]), 1)             # child = colour 1: output still shows colour 0
    r1 = onp.records_view(rec1, onp.parse_header(rec1))
    assert n1 == 1 and onp.parse_header(rec1)["colors"][0]["sample_name"] == "mom"
    assert onp.decode_kmers(r1["kmer"], 5).tobytes().decode() == "AAAAA" + "AAACC" + "ACGTA" + "GATTA"
    assert r1["cov"][:, 0].tolist() == [3, 0, 9, 2 ** 31]
    # CovStats: only ACGTA counts (child 9 > 0, one parent, one other): weight 2
    assert onp.cov_stats(g, 0, [1, 2]) == [(9, 2)]
    assert onp.cov_stats(g, 1, [0]) == [(1, 2)]              # child = mom, parent = kid: only ACGTA (kid + ref); GATTA's kid coverage is negative


@pytest.mark.parametrize("k,c,n", [(31, 3, 3000), (47, 4, 4000), (63, 6, 1500), (95, 2, 800)])
def test_prefilter_oracles_c_and_numpy_agree(k, c, n):
    """The two independent restatements of FindLowCoverage / FindShared / RecoverExcludedKmers / CovStats -- the record-by-record C
    loops (through the oracle's own getRecord / findRecord) and the vectorised numpy ones -- give the same files and tables on
    synthetic graphs with adversarial coverages (>= 2^31, i.e. negative Java ints)."""
    ctx = synth.make_ctx_file(500 + k, n, k, c, novel_permille=60, adv_period=37)
    hg = onp.parse_header(ctx)
    g = onp.records_view(ctx, hg)
    G = orc.Graph(ctx)
    rng = np.random.default_rng(k)
    # a one-colour ROI graph: a subset of the records with their colour-0 coverage
    pick = np.sort(rng.choice(n, n // 4, replace=False))
    roi_rec = np.zeros(len(pick), dtype=onp.record_dtype(hg["kmer_bits"], 1))
    roi_rec["kmer"], roi_rec["cov"][:, 0], roi_rec["edges"][:, 0] = g["kmer"][pick], g["cov"][pick, 0], g["edges"][pick, 0]
    roi = onp.write_header(k, hg["kmer_bits"], [hg["colors"][0]]) + roi_rec.tobytes()
    R = orc.Graph(roi)
    hr = onp.parse_header(roi)
    for m in (1, 7, 40, -3):
        want = onp.find_low_coverage(roi, m)
        mask = orc.find_low_coverage_mask(R, m)
        assert onp._rewritten_header(hr, hr["colors"]) + roi_rec[mask].tobytes() == want
    for parents, ignore in (([1], []), ([1, 2], [c - 1]), ([-1], [-1]), ([], list(range(1, c)))):
        want = onp.find_shared(ctx, roi, 0, parents, ignore)
        mask = orc.find_shared_mask(G, R, 0, parents, ignore)
        assert mask is not None and onp._rewritten_header(hr, hr["colors"]) + roi_rec[mask].tobytes() == want
    stray = np.zeros(1, dtype=onp.record_dtype(hg["kmer_bits"], 1)); stray["kmer"][0, -1] = 2
    stray_file = onp.write_header(k, hg["kmer_bits"], [hg["colors"][0]]) + stray.tobytes()
    if orc.Graph(ctx).find_record(onp.decode_kmers(stray["kmer"], k)[0].tobytes()) < 0:
        if c > 2:            # a colour beside child and parent exists: its coverage is read from the null record
            assert orc.find_shared_mask(G, orc.Graph(stray_file), 0, [1], []) is None
            with pytest.raises(KeyError):
                onp.find_shared(ctx, stray_file, 0, [1], [])
        else:                # every colour excluded: the null record is never touched, the k-mer is simply not shared
            assert orc.find_shared_mask(G, orc.Graph(stray_file), 0, [1], []).tolist() == [False]
            assert onp.parse_header(onp.find_shared(ctx, stray_file, 0, [1], []))["num_records"] == 0
    # dirty graph: every third record plus coverage 0 / small / negative
    dpick = np.arange(0, n, 3)
    drec = np.zeros(len(dpick), dtype=onp.record_dtype(hg["kmer_bits"], 1))
    drec["kmer"] = g["kmer"][dpick]
    drec["cov"][:, 0] = rng.integers(0, 4, len(dpick)); drec["cov"][::11, 0] = 0x80000001
    for child in (0, 1):
        dirty = onp.write_header(k, hg["kmer_bits"], [hg["colors"][child]]) + drec.tobytes()
        want, nrec = onp.recover_excluded_kmers(ctx, dirty, child)
        written, cov0, crec = orc.recover_excluded_kmers_decisions(G, orc.Graph(dirty), child)
        out = np.zeros(int((written > 0).sum()), dtype=onp.record_dtype(hg["kmer_bits"], 1))
        out["kmer"], out["cov"][:, 0], out["edges"][:, 0] = g["kmer"][written > 0], cov0[written > 0].view(np.uint32), g["edges"][written > 0, 0]
        assert crec == nrec == int((written == 2).sum()) and nrec > 0
        assert onp._rewritten_header(hg, [hg["colors"][child]]) + out.tobytes() == want
    for child, parents in ((0, [1, 2] if c > 2 else [1]), (1, [0]), (0, [-1]), (c - 1, [0, 1])):
        key, weight = orc.cov_stats_pairs(G, child, parents)
        hist = {}
        for kk, ww in zip(key.tolist(), weight.tolist()):
            if kk:
                hist[kk] = hist.get(kk, 0) + ww
        assert sorted(hist.items()) == onp.cov_stats(ctx, child, [p for p in parents])


def test_remove_oracle_hand_checked_and_c_numpy_agree():
    """Remove.java:30-88 (parity unpinned by reference tests: none exist).  Hand-computed case first, then the record-by-record C
    restatement (CortexCollection.next + Remove's loop) against the vectorised numpy one on overlapping synthetic graphs."""
    # primary (2 colours): AAAAA, ACGTA, CCCAA, GATTA; secondary A (1 colour): ACGTA cov 4, GATTA cov 0, TTTAA cov 2^31; secondary B: CCCAA cov 1, GGGAA cov 0
    p = _tiny_graph(["AAAAA", "ACGTA", "CCCAA", "GATTA"], [[3, 1], [9, 0], [2, 2], [5, 7]], ["kid", "mom"])
    a = _tiny_graph(["ACGTA", "GATTA", "TTTAA"], [[4], [0], [2 ** 31]], ["x"])
    b = _tiny_graph(["CCCAA", "GGGAA"], [[1], [0]], ["y"])
    out, removed = onp.remove(p, [a, b])
    h = onp.parse_header(out)
    r = onp.records_view(out, h)
    # ACGTA and CCCAA are covered by a secondary graph -> removed.  GATTA (secondary coverage 0) stays.  GGGAA and TTTAA exist only in
    # secondaries with coverage <= 0 (TTTAA's wraps negative): the reference writes them with all-zero primary colours.
    assert removed == 2 and h["num_colors"] == 2 and [c["sample_name"] for c in h["colors"]] == ["kid", "mom"]
    assert onp.decode_kmers(r["kmer"], 5).tobytes().decode() == "AAAAA" + "GATTA" + "GGGAA" + "TTTAA"
    assert r["cov"].tolist() == [[3, 1], [5, 7], [0, 0], [0, 0]] and r["edges"][2:].tolist() == [[0, 0], [0, 0]]
    got, rem2 = orc.remove_records(orc.Graph(p), [orc.Graph(a), orc.Graph(b)])
    assert rem2 == removed and got == out[h["data_offset"]:]
    assert onp.remove(p, [])[0][onp.parse_header(onp.remove(p, [])[0])["data_offset"]:] == p[onp.parse_header(p)["data_offset"]:]
    for k, shapes in ((31, [(2, 1500), (1, 1200), (3, 800)]), (47, [(4, 3000), (1, 2500)]), (95, [(1, 700), (2, 0), (1, 900)])):
        pool = synth.random_canonical_keys(5 + k, 4000, k, "cpu")
        ctxs = []
        for gi, (c, n) in enumerate(shapes):
            g = torch.Generator().manual_seed(gi)
            pick = torch.sort(torch.randperm(len(pool[0]), generator=g)[:n]).values
            cov, edges = synth.coverage_and_edges(50 + gi, n, c, "cpu", novel_permille=100, adv_period=7)
            body = synth.assemble_records([w[pick] for w in pool], cov, edges) if n else torch.zeros((0, 8 * len(pool) + 5 * c), dtype=torch.uint8)
            ctxs.append(synth.header_bytes(k, c, ["g%d_%d" % (gi, j) for j in range(c)]) + body.numpy().tobytes())
        want, removed = onp.remove(ctxs[0], ctxs[1:])
        got, rem2 = orc.remove_records(orc.Graph(ctxs[0]), [orc.Graph(x) for x in ctxs[1:]])
        assert rem2 == removed and removed > 0 and got == want[onp.parse_header(want)["data_offset"]:]
