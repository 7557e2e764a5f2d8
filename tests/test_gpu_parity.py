"""Parity of the CUDA path (through the C ABI of libcorticall_cuda) against the CPU oracle and the reference's
golden vectors.  Mirrors T/utils/kmer/CortexGraphTest.java (recordsAreCorrect, numRecordsTest, testGetRecord,
testSortedFindRecord, testFindNonExistentRecord) on the device-backed CortexGraph, then widens to the shapes
BASELINE.json names.  Integer / byte work: every comparison is bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import corticall_b200 as cb                      # noqa: E402
from corticall_b200 import _native as N          # noqa: E402
from oracle import orc                           # noqa: E402  (the checker)
from tools import synth                          # noqa: E402


@pytest.fixture(scope="module")
def graph(fixture_ctx):
    g = cb.CortexGraph(fixture_ctx)
    yield g
    g.dispose()


def t2words(body_words):
    return [w for w in body_words]


# ------------------------------------------------------------------ the reference's own tests, on the GPU graph

def test_header(graph):
    assert (graph.getVersion(), graph.getKmerSize(), graph.getKmerBits(), graph.getNumColors()) == (6, 31, 1, 2)
    assert graph.getNumRecords() == 66                                   # numRecordsTest :147-152
    assert graph.getColor(0).getSampleName() == "one" and graph.getColor(1).getSampleName() == "two"   # :139-145
    assert graph.getColorForSampleName("ONE") == 0 and graph.getColorForSampleName("1") == 1
    assert graph.getColorForSampleName("nope") == -1
    assert graph.getColorsForSampleNames(["two", "one", "x"]) == [1, 0, -1]
    assert graph.getColor(0).getMeanReadLength() == 63 and graph.getColor(0).getCleanedAgainstGraphName() == "undefined"


def test_records_are_correct(graph, kats):                               # recordsAreCorrect :186-198
    n = 0
    for cr, row in zip(graph, kats["fixture_records"]):
        assert cr.getKmerAsString() == row["kmer"]
        assert cr.getCoverages() == row["coverage"]
        assert cr.getEdgeAsStrings() == row["edges"]
        assert cr.toString() == "%s %d %d %s %s" % (row["kmer"], *row["coverage"], *row["edges"])
        n += 1
    assert n == 66
    assert sum(1 for _ in graph) == 66                                   # re-iteration works (TraversalEngineTest :35-45)


def test_get_record_backwards(graph, kats):                              # testGetRecord :255-265
    for i in range(10, -1, -1):
        assert graph.getRecord(i).getKmerAsString() == kats["fixture_records"][i]["kmer"]
    assert graph.getRecord(66) is None
    with pytest.raises(cb.CortexJDKException):
        graph.getRecord(-1)


def test_encode_binary_kmer(graph):                                      # testEncodeBinaryKmer :267-280
    for i in range(10, -1, -1):
        cr = graph.getRecord(i)
        assert cb.CortexRecord.encodeBinaryKmer(cr.getKmerAsBytes()) == cr.getBinaryKmer()
        assert cb.CortexRecord.decodeBinaryKmer(cr.getBinaryKmer(), 31, 1) == cr.getKmerAsBytes()


def test_sorted_find_record(graph, kats):                                # testSortedFindRecord :310-320
    for row in kats["fixture_records"]:
        cr = graph.findRecord(row["kmer"])
        assert cr is not None and cr.getKmerAsString() == row["kmer"] and cr.getCoverages() == row["coverage"]
        rc = cb.SequenceUtils.reverseComplement(row["kmer"])
        assert graph.findRecord(rc) == cr
        assert graph.findRecord(cb.CanonicalKmer(rc)) == cr and graph.findRecord(cb.CortexByteKmer(row["kmer"])) == cr


def test_find_non_existent_record(graph, kats):                          # testFindNonExistentRecord :322-331
    assert graph.findRecord(kats["missing_query"]) is None
    assert graph.findRecord(kats["fixture_records"][5]["kmer"].lower()) is None
    assert graph.findRecord("A" * 30) is None


def test_cortex_map(graph, fixture_ctx, kats):                           # CortexMap.java:14-160
    """CortexMap = the reference's HashMap<CortexBinaryKmer, CortexRecord> view: built here the way the reference builds it
    (a dict over the iterated records) and compared with the mirror, whose map is the device index."""
    want = {cr.getCortexBinaryKmer(): cr for cr in cb.CortexGraph(fixture_ctx)}
    m = cb.CortexMap(fixture_ctx)
    assert m.getNumRecords() == len(want) and m.getKmerSize() == 31 and sum(1 for _ in m) == len(want)
    queries = []
    for row in kats["fixture_records"]:
        rc = cb.SequenceUtils.reverseComplement(row["kmer"])
        rc = rc.decode() if isinstance(rc, bytes) else rc
        queries += [row["kmer"], rc, row["kmer"].lower(), rc[:15] + rc[15:].lower(), row["kmer"][1:], "A" + row["kmer"][:-1]]
    queries += [kats["missing_query"].replace("N", "A"), "A" * 31, "T" * 31, "ACGT" * 7 + "ACG", "C" * 33]
    hits = 0
    for q in queries:
        key = cb.CortexBinaryKmer(q)
        got = m.findRecord(q)
        assert got == want.get(key), q
        assert m.findRecord(q.encode()) == got and m.findRecord(key) == got
        hits += got is not None
    assert hits >= 2 * len(kats["fixture_records"])
    cr = graph.getRecord(7)
    assert m.findRecord(cb.CanonicalKmer(cr.getKmerAsString())) == cr and m.findRecord(cb.CortexByteKmer(cr.getKmerAsBytes())) == cr
    with pytest.raises(RuntimeError):                                    # charToBinaryNucleotide: 'N' is not a valid nucleotide
        m.findRecord(kats["missing_query"])
    m.graph.dispose()


def test_all_fasta_windows_hit(graph, fixture_ctx, fixture_fa):          # BASELINE.json configs[0]
    og = orc.Graph(fixture_ctx)
    seen = set()
    for seq in fixture_fa:
        for algo in (cb.CC_ALGO_AUTO, cb.CC_ALGO_BSEARCH, cb.CC_ALGO_MERGE):
            idx = graph.findWindows(seq, algo)
            assert idx.tolist() == og.find_windows(seq).tolist() and (idx >= 0).all()
        seen.update(idx.tolist())
        assert graph.containsWindows(seq).all()
    assert seen == set(range(66))


def test_novelty_fixture(graph, fixture_ctx, tmp_path):
    og = orc.Graph(fixture_ctx)
    for child, parents, n in ((0, [1], 19), (1, [0], 47), (0, [], 19), (0, [0], 0), (0, [1, 1], 19)):
        cnt, recs, idx = graph.findNovel(child, parents)
        want, widx = og.find_rois(child, parents)
        assert cnt == n and recs.tobytes() == want and idx.tolist() == widx.tolist()
    # FindROIs.execute end to end: the file on disk is header + records, and reads back as a sorted 1-colour graph
    out = tmp_path / "rois.ctx"
    assert graph.writeRois(0, [1], out) == 19
    want, _ = og.find_rois(0, [1])
    assert out.read_bytes() == orc.roi_header(31, 1, "one") + want
    roi = cb.CortexGraph(out)
    assert roi.getNumColors() == 1 and roi.getNumRecords() == 19 and roi.getSampleName(0) == "one"
    ks = [cr.getKmerAsString() for cr in roi]
    assert ks == sorted(ks)
    assert roi.findRecord(ks[7]).getKmerAsString() == ks[7]
    roi.dispose()
    with pytest.raises(cb.CortexJDKException):
        graph.findNovel(2, [0])                                          # colour out of range (Java: ArrayIndexOutOfBounds)
    with pytest.raises(cb.CortexJDKException):
        graph.findNovel(0, [-1])


# ------------------------------------------------------------------ synthetic graphs of the named shapes vs the oracle

SHAPES = [  # k, c, n
    (31, 4, 20000),      # config #4's record shape (28 B, 4-byte aligned)
    (47, 4, 50000),      # config #2/#3 (36 B)
    (63, 21, 6000),      # config #5 (121 B, odd)
    (31, 1, 5000),       # 13 B
    (5, 3, 300),
    (95, 2, 3000),       # 3 words
    (127, 5, 2000),      # 4 words
    (33, 7, 4097),
    (47, 4, 31), (47, 4, 32), (47, 4, 33), (47, 4, 3), (31, 2, 1),
    (31, 45, 2500),      # 233-byte records: 256-record tiles do not fit -> the general (per-tile look-back) scan kernel
]


@pytest.mark.parametrize("k,c,n", SHAPES)
def test_scan_and_lookup_vs_oracle(k, c, n):
    ctx = synth.make_ctx_file(4321 + k + n, n, k, c, novel_permille=15, adv_period=97, trailing=b"\x01\x02\x03")
    g = cb.CortexGraph(ctx)
    og = orc.Graph(ctx)
    assert g.getNumRecords() == n == og.h.num_records
    # K1 decode: all columns
    words, cov, edges = g.decodeRecords(0, n)
    raw = g.getRawRecords(0, n)
    s = g.getKmerBits()
    assert (raw[:, :8 * s].copy().view("<u8") == words).all()
    assert (raw[:, 8 * s:8 * s + 4 * c].copy().view("<i4") == cov).all()
    assert (raw[:, 8 * s + 4 * c:] == edges).all()
    for i in sorted({0, n // 3, n - 1}):
        bk, ocov, oed = og.get_record(i)
        assert words[i].byteswap().view(np.int64).tolist() == bk.tolist() and cov[i].tolist() == ocov.tolist()
    if n > 40:
        w2, c2, e2 = g.decodeRecords(7, 33)                              # unaligned sub-range
        assert (w2 == words[7:40]).all() and (c2 == cov[7:40]).all() and (e2 == edges[7:40]).all()
    # K1+K2: several colour choices
    choices = [(0, list(range(1, c))), (0, []), (c - 1, [0]), (0, [c - 1, c - 1])]
    for child, parents in choices:
        cnt, recs, idx = g.findNovel(child, parents)
        want, widx = og.find_rois(child, parents)
        assert cnt == len(widx) and recs.tobytes() == want and idx.tolist() == widx.tolist(), (child, parents)
    # capped output: count is still the total, only `cap` stored
    cnt, recs, idx = g.findNovel(0, [], cap=5)
    want, widx = og.find_rois(0, [])
    assert cnt == len(widx) and recs.tobytes() == want[:min(5, cnt) * (8 * s + 5)]
    # K3+K4 lookups (hits on both strands, misses, N, lower case), all three algorithms
    tw = [torch.from_numpy(words[:, w].copy().view(np.int64)) for w in range(s)]
    q_ascii, canon, valid = synth.make_queries(99, tw, k, 4000, corrupt_permille=30)
    qa = q_ascii.numpy()
    if n >= 3:
        want_idx = og.find_batch(qa)
    else:
        # N <= 2: the reference's search loop never runs and only its LRU answers; the GPU path
        # returns the exact match.  Compare with the oracle's answer once every record is cached (after iteration).
        for i in range(n):
            og.get_record(i)
        want_idx = og.find_batch(qa)
    for algo in (cb.CC_ALGO_AUTO, cb.CC_ALGO_BSEARCH, cb.CC_ALGO_MERGE):
        got = g.findRecordIndices(qa, algo)
        assert got.tolist() == want_idx.tolist(), algo
    assert (want_idx[~valid.numpy()] == -1).all()
    # packed queries
    pw = np.stack([cw.numpy().view(np.uint64) for cw in canon], axis=1)
    flags = np.where(valid.numpy(), 0, 2).astype(np.uint8)
    got = g.findPacked(pw, flags)
    assert got.tolist() == want_idx.tolist()
    g.dispose()


@pytest.mark.parametrize("k", [5, 21, 31, 32, 33, 47, 63, 64, 65, 95, 128])
def test_pack_canonical_vs_oracle(k):
    seq = synth.random_genome(11 + k, 20000, n_permille=2).numpy().copy()
    seq[300:420] += 32                                                   # a lower-case run
    seq[5000] = ord(".")
    seq[7000:7003] = [ord("n"), 200, 0]
    for ln in (len(seq), 9001, k, k + 1, k - 1):
        sub = seq[:ln]
        w, f = cb.packCanonical(sub, k)
        ow, of = orc.pack_windows(sub, k)
        assert w.shape == ow.shape and (w == ow).all() and (f == of).all()
    # unaligned start of the sequence buffer
    w, f = cb.packCanonical(seq[3:4000], k)
    ow, of = orc.pack_windows(seq[3:4000], k)
    assert (w == ow).all() and (f == of).all()


@pytest.mark.parametrize("k", [5, 21, 31, 32, 33, 47, 63, 64, 65, 96, 127])
def test_pack_canonical_soft_masked_genome(k):
    """Mixed-case sequence (half the bases lower case, in runs and singly, as in a soft-masked genome) with a few N / n: the
    orientation is the reference's ASCII byte rule (SequenceUtils.java:211-219; every upper-case letter sorts before every
    lower-case one), evaluated on the device from packed codes + case bits.  Near-palindromes put the first difference deep
    inside the k-mer and at a case-only position."""
    rng = np.random.default_rng(100 + k)
    n = 30000
    seq = synth.random_genome(5 + k, n).numpy().copy()
    # near-palindromic stretches: a random half followed by its reverse complement, then single-base edits
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    for at in range(500, n - 400, 1500):
        half = seq[at:at + 150].copy()
        seq[at + 150:at + 300] = comp[half[::-1]]
    low = rng.random(n) < 0.5
    for at in range(0, n, 977):                         # runs of one case
        low[at:at + int(rng.integers(1, 200))] = rng.random() < 0.5
    seq[low] += 32
    seq[rng.integers(0, n, 12)] = ord("N")
    seq[rng.integers(0, n, 5)] = ord("n")
    w, f = cb.packCanonical(seq, k)
    ow, of = orc.pack_windows(seq, k)
    assert (f == of).all() and (w == ow).all()
    assert ((f & 4) != 0).sum() > n // 4 and ((f & 1) != 0).sum() > n // 8
    # the fused windows lookup treats every such window as a miss (findRecord compares with upper-case record k-mers)
    ctx = synth.make_ctx_file(1, min(1000, 4 ** k // 8), k, 1, adv_period=0)
    g = cb.CortexGraph(ctx)
    idx = g.findWindows(seq[:5000])
    assert (idx[(f[:len(idx)] & 6) != 0] == -1).all()
    g.dispose()


def test_lowest_orientation_property_gpu():
    """SequenceUtilsTest :59-72 on the GPU packer: canonical == min(fw, rc) for random k in {21,31,41,51}."""
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for k in (21, 31, 41, 51):
        seq = synth.random_genome(k, 10000 + k - 1).numpy()
        w, f = cb.packCanonical(seq, k)
        sb = seq.tobytes()
        for i in range(0, 10000, 97):
            fw = sb[i:i + k]
            rc = fw.translate(comp)[::-1]
            exp = min(fw, rc)
            assert cb.CortexRecord.decodeBinaryKmer([int(x) for x in w[i].byteswap().view(np.int64)], k, (k + 31) // 32) == exp
            assert bool(f[i] & 1) == (exp != fw)


@pytest.mark.parametrize("mis", [0, 1, 4, 7, 8, 13, 15])
@pytest.mark.parametrize("k,c", [(47, 4), (63, 21), (31, 1)])
def test_misaligned_device_body(mis, k, c):
    """A device body at any byte offset (a shard cut out of a larger buffer): tiles are fetched as aligned supersets."""
    n = 30011
    body, words = synth.make_graph_body(77 + mis, n, k, c, novel_permille=30, adv_period=101)
    S = body.shape[1]
    buf = torch.zeros(n * S + 64, dtype=torch.uint8, device="cuda")
    buf[16 + mis:16 + mis + n * S] = body.reshape(-1).cuda()
    g = cb.CortexGraph.fromDevice(buf.data_ptr() + 16 + mis, k, c, n, firstIndex=1000, keepalive=buf)
    cnt, recs, idx = g.findNovel(0, list(range(1, c)))
    wcnt, wrecs, widx = orc.find_rois_body(body.numpy().reshape(-1), n, k, (k + 31) // 32, c, 0, list(range(1, c)))
    assert cnt == wcnt and recs.tobytes() == wrecs.tobytes() and (idx == widx + 1000).all()
    w, cv, e = g.decodeRecords(0, n)
    assert (w == np.stack([x.numpy().view(np.uint64) for x in words], axis=1)).all()
    # lookups report indices rebased by firstIndex
    q = synth.words_to_ascii([x[:50] for x in words], k).numpy()
    assert g.findRecordIndices(q).tolist() == list(range(1000, 1050))
    g.dispose()


def test_host_streamed_scan():
    """cc_find_novel_host: the record array stays on the host and is streamed through the device in chunks."""
    k, c, n = 47, 4, 300_000
    body, _ = synth.make_graph_body(5, n, k, c, novel_permille=8, adv_period=1009)
    flat = body.numpy().reshape(-1)
    N.set_option("host_chunk_mb", 1)                                     # many chunks -> exercises the carry between launches
    try:
        par = np.array([1, 2, 3], dtype=np.int32)
        out = np.empty((n, 21), dtype=np.uint8); idx = np.empty(n, dtype=np.uint64)
        cnt = C.c_uint64(); st = N.Stats()
        N.check(N.lib().cc_find_novel_host(0, flat.ctypes.data, k, 2, c, n, 0, par.ctypes.data, 3, out.ctypes.data,
                                           idx.ctypes.data, n, C.byref(cnt), C.byref(st)))
    finally:
        N.set_option("host_chunk_mb", 64)
    wcnt, wrecs, widx = orc.find_rois_body(flat, n, k, 2, c, 0, [1, 2, 3])
    assert cnt.value == wcnt and out[:wcnt].tobytes() == wrecs.tobytes() and (idx[:wcnt] == widx).all()
    assert st.h2d_bytes == n * 36 and st.launches >= 10


def test_unsorted_graph_is_rejected():
    ctx = bytearray(synth.make_ctx_file(3, 900, 15, 1, adv_period=0))
    g0 = orc.Graph(bytes(ctx))
    S, off = g0.h.record_size, g0.h.data_offset
    a, b = bytes(ctx[off + 10 * S:off + 11 * S]), bytes(ctx[off + 500 * S:off + 501 * S])
    ctx[off + 10 * S:off + 11 * S], ctx[off + 500 * S:off + 501 * S] = b, a
    g = cb.CortexGraph(bytes(ctx))
    with pytest.raises(cb.CortexJDKException) as ei:                     # CortexGraph.java:295-301
        g.findRecord("ACGTACGTACGTACG")
    assert ei.value.status == N.CC_ERR_UNSORTED and "not sorted" in str(ei.value)
    # the scan does not need sorted input (FindROIs just iterates)
    cnt, _, _ = g.findNovel(0, [])
    assert cnt == orc.Graph(bytes(ctx)).find_rois(0, [])[1].size
    g.dispose()


def test_empty_graph():
    ctx = synth.header_bytes(31, 3)
    g = cb.CortexGraph(ctx)
    assert g.getNumRecords() == 0 and list(g) == []
    cnt, recs, idx = g.findNovel(0, [1, 2])
    assert cnt == 0 and len(recs) == 0
    assert g.findRecord("A" * 31) is None
    assert g.findWindows("ACGT" * 20).tolist() == [-1] * 50
    g.dispose()


def test_dense_novel_output():
    """Every record novel (the staging buffer must grow and the ordered write-out must hold at 100 % density)."""
    k, c, n = 31, 2, 150_000
    body, _ = synth.make_graph_body(9, n, k, c, adv_period=0)
    body[:, 8 + 4:8 + 8] = 0                                             # parent coverage := 0
    body[:, 8] |= 1                                                      # child coverage > 0
    ctx = synth.header_bytes(k, c) + body.numpy().tobytes()
    g = cb.CortexGraph(ctx)
    cnt, recs, idx = g.findNovel(0, [1])
    want, widx = orc.Graph(ctx).find_rois(0, [1])
    assert cnt == n and recs.tobytes() == want and idx.tolist() == widx.tolist()
    g.dispose()


def test_full_size_config2_scan():
    """BASELINE.json configs[1] at full size (2.5e7 records, k=47, 4 colours, 0.9 GB), generated on the device:
    bit-exact against the oracle run over the same bytes, plus size-independent properties."""
    k, c, n = 47, 4, 25_000_000
    body, words = synth.make_graph_body(20261018, n, k, c, device="cuda")
    torch.cuda.synchronize()
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
    cnt, recs, idx = g.findNovel(0, [1, 2, 3])
    # properties: sorted unique indices; every emitted record equals the source record's projection
    assert (np.diff(idx.astype(np.int64)) > 0).all()
    host = body.cpu().numpy()
    assert (recs[:, :16] == host[idx.astype(np.int64), :16]).all()
    assert (recs[:, 16:20] == host[idx.astype(np.int64), 16:20]).all() and (recs[:, 20] == host[idx.astype(np.int64), 32]).all()
    # independent predicate evaluated with torch on the device
    cov = body[:, 16:32].contiguous().view(torch.int32)
    mask = (cov[:, 0] > 0) & (cov[:, 1] == 0) & (cov[:, 2] == 0) & (cov[:, 3] == 0)
    assert int(mask.sum()) == cnt and torch.equal(torch.nonzero(mask).flatten().cpu(), torch.from_numpy(idx.astype(np.int64)))
    # the oracle over the full array
    wcnt, wrecs, widx = orc.find_rois_body(host.reshape(-1), n, k, 2, c, 0, [1, 2, 3], cap=cnt + 10)
    assert wcnt == cnt and wrecs.tobytes() == recs.tobytes() and (widx == idx).all()
    # idempotence: scanning the ROI set again with no parents returns it unchanged
    roi_ctx = orc.roi_header(k, 2, "child") + recs.tobytes()
    rg = cb.CortexGraph(roi_ctx)
    c2, r2, _ = rg.findNovel(0, [])
    assert c2 == cnt and r2.tobytes() == recs.tobytes()
    # and every ROI k-mer is found in the big graph at its recorded index
    got = g.findPacked(recs[:, :16].copy().view("<u8"))
    assert (got == idx.astype(np.int64)).all()
    rg.dispose(); g.dispose()


# ------------------------------------------------------------------ multi-GPU helpers on one device

@pytest.mark.parametrize("nshards", [1, 2, 5, 8])
def test_bucket_by_owner_and_scatter(nshards):
    """cc_bucket_by_owner_dev / cc_scatter_results_dev against numpy: counts per owner, stable grouping, original slots."""
    k, n, nq = 47, 20000, 50000
    table = synth.random_canonical_keys(3, n, k, "cpu")
    ascii_q, canon, valid = synth.make_queries(5, table, k, nq, corrupt_permille=20)
    words = torch.stack(canon, dim=1).contiguous().cuda()
    flags = torch.where(valid, 0, 2).to(torch.uint8).cuda()
    spl = torch.stack([torch.stack([w[n * r // nshards] for w in table]) for r in range(1, nshards)]).cuda() if nshards > 1 else None
    from corticall_b200.host.sharded import CudaOps
    ops = CudaOps(None, torch.device("cuda", 0))
    counts, sorted_words, slots = ops.bucket(words, flags, spl, nshards)
    torch.cuda.synchronize()
    w = words.cpu().numpy().view(np.uint64)
    keyf = lambda a: [tuple(int(x) for x in row) for row in a]
    sp = keyf(spl.cpu().numpy().view(np.uint64)) if nshards > 1 else []
    import bisect
    owner = np.array([bisect.bisect_right(sp, t) for t in keyf(w)])
    ok = valid.numpy()
    assert counts.cpu().tolist() == np.bincount(owner[ok], minlength=nshards).tolist()
    tot = int(ok.sum())
    got_slots = slots.cpu().numpy()[:tot].astype(np.int64)
    assert sorted(got_slots.tolist()) == np.nonzero(ok)[0].tolist()              # a permutation of the routed queries
    assert (sorted_words.cpu().numpy().view(np.uint64)[:tot] == w[got_slots]).all()
    assert (np.diff(owner[got_slots]) >= 0).all()                                 # grouped by owner, ascending
    # the return leg
    vals = torch.arange(tot, dtype=torch.int64, device="cuda") * 3 + 1
    out = torch.full((nq,), -1, dtype=torch.int64, device="cuda")
    ops.scatter(vals, slots[:tot], out)
    torch.cuda.synchronize()
    exp = np.full(nq, -1, dtype=np.int64)
    exp[got_slots] = np.arange(tot) * 3 + 1
    assert (out.cpu().numpy() == exp).all()


# ------------------------------------------------------------------ the callers: FindROIs.execute and the Call helpers

def test_findrois_command_and_call_helpers(tmp_path):
    """FindROIs -g trio.ctx -p mom -p dad -c child -o rois.ctx, then Call's loadRois / loadChildWalk / getRegions /
    section ROI set on contigs, each against the oracle's per-window findRecord and a Python set of ROI k-mers."""
    k, c, n = 31, 4, 40000
    ctx = synth.make_ctx_file(99, n, k, c, novel_permille=40, adv_period=211)
    path = tmp_path / "trio.ctx"
    path.write_bytes(ctx)
    graph = cb.CortexGraph(path)
    out = tmp_path / "rois.ctx"
    cmd = cb.FindROIs(graph, ["mom", "dad", "REF"], "Child", out)            # names are matched case-insensitively
    nn = cmd.execute()
    og = orc.Graph(ctx)
    want, widx = og.find_rois(0, [1, 2, 3])
    assert nn == len(widx) and out.read_bytes() == orc.roi_header(k, 1, "child") + want
    with pytest.raises(cb.CortexJDKException):                               # unknown sample -> colour -1 -> Java AIOOBE
        cb.FindROIs(graph, ["mom", "nobody"], "child", tmp_path / "x.ctx").execute()

    rois = cb.CallHelpers.loadRois(cb.CortexGraph(out))
    roi_set = {cr.getKmerAsString() for cr in rois}
    assert len(roi_set) == nn
    # a contig: stitched from graph k-mers (so many windows hit), with an N and a lower-case stretch
    words, _, _ = graph.decodeRecords(0, n)
    kmers = [graph._make_record(words[i], [0] * c, [0] * c).getKmerAsString() for i in list(widx[:40].astype(int)) + list(range(0, 400, 7))]
    contig = "".join(kmers[:60])
    contig = contig[:500] + "N" + contig[501:900] + contig[900:960].lower() + contig[960:]
    walk = cb.CallHelpers.loadChildWalk(contig, graph)
    want_idx = og.find_windows(contig.encode())
    assert [v.recordIndex for v in walk] == want_idx.tolist()
    seen = {}
    for i, v in enumerate(walk):
        sk = contig[i:i + k]
        seen[sk] = seen[sk] + 1 if sk in seen else 0
        assert v.bases == sk and v.copyIndex == seen[sk]
        if want_idx[i] >= 0:
            assert v.record.getKmerAsString() == cb.CanonicalKmer(sk).getKmerAsString()
        else:
            assert v.record is None
    canon = [cb.SequenceUtils.alphanumericallyLowestOrientation(contig[i:i + k]) for i in range(len(contig) - k + 1)]
    member = [x in roi_set for x in canon]
    assert rois.containsWindows(contig).tolist() == member
    regions, start = [], -1
    for i, m in enumerate(member + [False]):
        if m and start < 0:
            start = i
        if not m and start >= 0:
            regions.append((start, i - 1)); start = -1
    assert cb.CallHelpers.getRegions(rois, contig) == regions and len(regions) > 0
    assert [x.getKmerAsString() for x in cb.CallHelpers.sectionRois(rois, contig)] == sorted({x for x, m in zip(canon, member) if m})
    # Call.makeNoveltyTrack :2062-2070 (the '*' track over the gap-free query) and Call.trimQuery :1946-1986, restated literally
    track = [" "] * (len(contig) + 1)
    for i, m in enumerate(member):
        if m:
            track[i:i + k] = "*" * k
    assert cb.CallHelpers.noveltyMask(rois, contig) == "".join(track)
    sub = contig[700:1300]                                                    # holds the lower-case stretch
    rc_piece = cb.SequenceUtils.reverseComplement(contig[1500:1600].encode()).decode()
    for targets in ({"t1": sub[40:200], "t2": rc_piece}, {"t": "ACGT" * 30}, {"a": contig[905:950], "b": contig[480:540]}):
        ws_canon = [cb.CanonicalKmer(contig[i:i + k]) for i in range(len(contig) - k + 1)]
        pos = {}
        first_novel = last_novel = -1
        for i, ck in enumerate(ws_canon):
            pos.setdefault(ck, []).append(i)
            if ck.getKmerAsString() in roi_set:
                first_novel = i if first_novel == -1 else first_novel
                last_novel = i
        fi, li = 2 ** 31 - 1, 0
        for t in targets.values():
            for i in range(len(t) - k + 1):
                ck = cb.CanonicalKmer(t[i:i + k])
                if ck in pos:
                    fi, li = min(fi, pos[ck][0]), max(li, pos[ck][-1])
        if first_novel < fi:
            fi = first_novel
        if last_novel > li:
            li = last_novel
        assert cb.CallHelpers.trimQuery(contig, targets, rois) == (fi, li + 1, contig[fi:li + k]), list(targets)
    graph.dispose(); rois.dispose()


@pytest.mark.parametrize("world,k,vsub", [(1, 47, 1), (2, 47, 1), (4, 47, 1), (8, 47, 1), (3, 16, 1), (3, 31, 1), (2, 63, 1), (3, 65, 1), (5, 96, 1),
                                          (1, 47, 64), (2, 47, 32), (8, 47, 8), (3, 31, 5), (4, 65, 16)])
def test_routed_lookup_emulated_ranks(world, k, vsub):
    """The peer-memory lookup path (route -> search -> gather) with all ranks emulated on one device: plain device
    tensors stand in for the symmetric allocations (every 'peer pointer' is a local pointer) and the legs of all ranks
    run one after another on one stream, which is exactly the ordering the cross-rank barriers enforce."""
    from corticall_b200.host.sharded import RoutedLookup
    c, n = 4, 60000
    nw = (k + 31) // 32
    ctx = synth.make_ctx_file(31, n, k, c, adv_period=0)
    og = orc.Graph(ctx)
    whole = cb.CortexGraph(ctx)
    words_all, _, _ = whole.decodeRecords(0, n)
    body = torch.from_numpy(whole.getRawRecords(0, n)).cuda()
    table = [torch.from_numpy(words_all[:, w].copy().view(np.int64)) for w in range(nw)]
    dev = torch.device("cuda", 0)
    nq_per = [5000 + 700 * r for r in range(world)]
    cap = max(nq_per)
    blocks = [torch.zeros(RoutedLookup.block_elems(world, cap, k, vsub), dtype=torch.int64, device=dev) for _ in range(world)]
    first = [n * r // world for r in range(world)]
    shards, rls, qs = [], [], []
    for r in range(world):
        lo, hi = n * r // world, n * (r + 1) // world
        shards.append(cb.CortexGraph.fromDevice(body[lo:hi].data_ptr(), k, c, hi - lo, firstIndex=lo, keepalive=body))
    # splitters: first key of every (virtual) shard but the first; with vsub = 1 these are the rank boundaries
    spl = RoutedLookup.virtual_splitters(None, 0, world, vsub, dev, emulate_graphs=shards)
    if vsub == 1 and world > 1:
        assert torch.equal(spl.cpu(), torch.stack([torch.stack([t[n * r // world] for t in table]) for r in range(1, world)]))
    for r in range(world):
        g = shards[r]
        rls.append(RoutedLookup(g, spl, r, world, dev, cap, k, shard_first=first, emulate=blocks, vsub=vsub))
        a, canon, valid = synth.make_queries(200 + r, table, k, nq_per[r], corrupt_permille=25)
        qs.append((a, torch.stack(canon, dim=1).contiguous().cuda(), torch.where(valid, 0, 2).to(torch.uint8).cuda(),
                   torch.full((nq_per[r],), -7, dtype=torch.int64, device=dev)))
    for r in range(world):
        rls[r].route(qs[r][1], qs[r][2])
    for r in range(world):
        rls[r].search()
    for r in range(world):
        rls[r].gather(qs[r][3])
    torch.cuda.synchronize()
    for r in range(world):
        want = og.find_batch(qs[r][0].numpy())
        assert (qs[r][3].cpu().numpy() == want).all(), r
        routed, valid = int(rls[r].sent[:world * vsub].sum()), int((qs[r][2] == 0).sum())
        assert valid <= routed <= valid + 3 * world * vsub * (nq_per[r] // 1024 + 1)      # runs are padded to multiples of 4 keys
    # a second batch through the same buffers (stale segment contents must not leak)
    for r in range(world):
        rls[r].route(qs[r][1][:100], qs[r][2][:100])
    for r in range(world):
        rls[r].search()
    for r in range(world):
        rls[r].gather(qs[r][3][:100])
    torch.cuda.synchronize()
    for r in range(world):
        assert (qs[r][3][:100].cpu().numpy() == og.find_batch(qs[r][0][:100].numpy())).all()
    for g in shards:
        g.dispose()
    whole.dispose()


def test_routed_lookup_reports_segment_overflow():
    """cap smaller than the batch: balanced batches work, a skewed batch is reported (dropped keys answer -1, sent > cap)."""
    from corticall_b200.host.sharded import RoutedLookup
    k, c, n, world = 31, 1, 20000, 2
    ctx = synth.make_ctx_file(5, n, k, c, adv_period=0)
    og = orc.Graph(ctx)
    whole = cb.CortexGraph(ctx)
    words_all, _, _ = whole.decodeRecords(0, n)
    body = torch.from_numpy(whole.getRawRecords(0, n)).cuda()
    table = [torch.from_numpy(words_all[:, 0].copy().view(np.int64))]
    dev = torch.device("cuda", 0)
    spl = torch.stack([torch.stack([t[n // 2] for t in table])]).cuda()
    nq, cap = 6000, 4000
    blocks = [torch.zeros(RoutedLookup.block_elems(world, cap, k), dtype=torch.int64, device=dev) for _ in range(world)]
    shards = [cb.CortexGraph.fromDevice(body[lo:hi].data_ptr(), k, c, hi - lo, firstIndex=lo, keepalive=body) for lo, hi in ((0, n // 2), (n // 2, n))]
    rls = [RoutedLookup(shards[r], spl, r, world, dev, cap, k, shard_first=[0, n // 2], emulate=blocks, max_batch=nq) for r in range(world)]
    a, canon, valid = synth.make_queries(1, table, k, nq, hit_fraction_permille=1000, corrupt_permille=0)
    qw = torch.stack(canon, dim=1).contiguous().cuda(); qf = torch.zeros(nq, dtype=torch.uint8, device=dev)
    out = torch.empty(nq, dtype=torch.int64, device=dev)

    def run(words):
        rls[0].route(words, qf); rls[1].route(words[:0], qf[:0])
        rls[0].search(); rls[1].search()
        rls[0].gather(out)
        torch.cuda.synchronize()

    run(qw)                                              # ~3000 per owner: fits
    rls[0].check_overflow()
    assert (out.cpu().numpy() == og.find_batch(a.numpy())).all()
    low = qw[qw[:, 0] < spl[0, 0]][:2500]
    skew = torch.cat([low, low])[:nq].contiguous()       # 5000 keys for owner 0 > cap
    qf = torch.zeros(skew.shape[0], dtype=torch.uint8, device=dev)
    run(skew)
    with pytest.raises(OverflowError):
        rls[0].check_overflow()
    got = out[:skew.shape[0]].cpu().numpy()
    # the segment holds `cap` slots; a few of them are the pad keys of the runs (multiples of 4 keys per tile and owner)
    assert cap - 3 * (nq // 1024 + 1) <= (got >= 0).sum() <= cap and ((got >= 0) | (got == -1)).all()
    for g in shards:
        g.dispose()
    whole.dispose()


@pytest.mark.parametrize("k", [5, 31, 32, 33, 47, 63, 64, 95, 128])
def test_pack_rows_vs_oracle(k):
    """cc_pack_kmers_dev (independent k-byte rows, the query-list form of K3) against the oracle's per-row canonicalise+pack,
    including N / lower-case / odd bytes, a ragged last warp and a misaligned base pointer."""
    nq = 70_003
    s = (k + 31) // 32
    seq = synth.random_genome(17 + k, nq * k + 5, n_permille=1).numpy().copy()
    seq[1000:1400] += 32
    seq[5000] = ord("."); seq[9000] = 200
    for off in (0, 5):
        rows = seq[off:off + nq * k].reshape(nq, k)
        buf = torch.from_numpy(seq).cuda()
        words = torch.zeros((nq, s), dtype=torch.int64, device="cuda"); flags = torch.zeros(nq, dtype=torch.uint8, device="cuda")
        N.check(N.lib().cc_pack_kmers_dev(0, buf.data_ptr() + off, nq, k, words.data_ptr(), flags.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        got_w, got_f = words.cpu().numpy().view(np.uint64), flags.cpu().numpy()
        # oracle: every row is a length-k sequence with exactly one window
        step = max(1, nq // 3000)
        for i in list(range(0, nq, step)) + [nq - 1, 1000 // k, 1400 // k, 5000 // k, 9000 // k]:
            ow, of = orc.pack_windows(rows[i], k)
            assert (got_w[i] == ow[0]).all() and got_f[i] == of[0], (k, off, i)


@pytest.mark.parametrize("k", [1, 7, 16, 47, 64, 97])
def test_pack_rows_every_row(k):
    """Every row of a large list (not a sample): flag classes from a numpy classification of the bytes, and for the pure
    upper-case rows the packed canonical words from an independent torch formulation (tools/synth.py)."""
    nq = 200_003
    s = (k + 31) // 32
    seq = synth.random_genome(5 + k, nq * k + 16, n_permille=1).numpy().copy()
    rng = np.random.default_rng(k)
    low_at = rng.integers(0, nq * k, 300)
    seq[low_at] |= 32                                     # lower case (or 'n')
    seq[rng.integers(0, nq * k, 50)] = ord("-")
    for off in (0, 3, 16):
        rows = seq[off:off + nq * k].reshape(nq, k)
        buf = torch.from_numpy(seq).cuda()
        words = torch.full((nq, s), -1, dtype=torch.int64, device="cuda"); flags = torch.full((nq,), 99, dtype=torch.uint8, device="cuda")
        N.check(N.lib().cc_pack_kmers_dev(0, buf.data_ptr() + off, nq, k, words.data_ptr(), flags.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        got_w, got_f = words.cpu(), flags.cpu().numpy()
        up = rows & 0xDF
        is_letter = (up == 65) | (up == 67) | (up == 71) | (up == 84)
        bad = ~is_letter.all(axis=1)
        low = ~bad & ((rows & 32) != 0).any(axis=1)
        assert ((got_f == 2) == bad).all()
        assert (((got_f & 4) != 0) == low).all()
        assert (got_w[torch.from_numpy(bad)] == 0).all()
        clean = ~bad & ~low
        assert clean.sum() > nq // 2 or k > 64
        code = np.zeros(256, dtype=np.uint64); code[67] = 1; code[71] = 2; code[84] = 3
        crow = code[rows[clean]]
        fw = [np.zeros(crow.shape[0], dtype=np.uint64) for _ in range(s)]
        for i in range(k):                               # base i sits 2*(k-1-i) bits above the bottom of the s-word number
            bit = 2 * (k - 1 - i)
            fw[s - 1 - bit // 64] |= crow[:, i] << np.uint64(bit % 64)
        canon, flipped = synth.canonical_words([torch.from_numpy(w.view(np.int64)) for w in fw], k)
        gw = got_w[torch.from_numpy(clean)]
        for w in range(s):
            assert torch.equal(gw[:, w], canon[w]), (k, off, w)
        assert ((got_f[clean] & 1) == flipped.numpy().astype(np.uint8)).all()
        # the sampled lower-case rows against the oracle's byte rule
        for i in np.nonzero(low)[0][:40]:
            ow, of = orc.pack_windows(rows[i], k)
            assert (got_w[i].numpy().view(np.uint64) == ow[0]).all() and got_f[i] == of[0], (k, off, i)


@pytest.mark.parametrize("fused", [1, 0])
def test_large_ascii_batch(fused):
    """A large query list through both forms of cc_find_ascii (pack + search in one kernel / pack, then search) returns
    what the oracle's findRecord does."""
    k, c, n = 47, 4, 30000
    ctx = synth.make_ctx_file(8, n, k, c, adv_period=0)
    g = cb.CortexGraph(ctx)
    og = orc.Graph(ctx)
    words, _, _ = g.decodeRecords(0, n)
    tw = [torch.from_numpy(words[:, w].copy().view(np.int64)) for w in range(2)]
    q_ascii, _, _ = synth.make_queries(3, tw, k, 90_001, corrupt_permille=20)
    N.set_option("rows_fused", fused)
    try:
        got = g.findRecordIndices(q_ascii.numpy())
    finally:
        N.set_option("rows_fused", 1)
    sub = np.arange(0, 90_001, 7)
    assert (got[sub] == og.find_batch(q_ascii.numpy()[sub])).all()
    assert (got >= 0).sum() > 30000
    g.dispose()


@pytest.mark.parametrize("k", [15, 31, 47, 65])
def test_shard_index_spans_its_own_key_range(k):
    """A k-mer-range shard opened on its own (cc_open_device, first_index = offset): its prefix table spans only the shard's
    [first key, last key]; queries below, inside and above that range return the global index or -1."""
    c, n = 2, 50_000
    ctx = synth.make_ctx_file(77, n, k, c, adv_period=0)
    whole = cb.CortexGraph(ctx)
    og = orc.Graph(ctx)
    s = whole.getKmerBits()
    words_all, _, _ = whole.decodeRecords(0, n)
    body = torch.from_numpy(whole.getRawRecords(0, n)).cuda()
    tw = [torch.from_numpy(words_all[:, w].copy().view(np.int64)) for w in range(s)]
    q_ascii, canon, valid = synth.make_queries(9, tw, k, 30_000, hit_fraction_permille=800, corrupt_permille=5)
    want = og.find_batch(q_ascii.numpy())
    pw = np.stack([cw.numpy().view(np.uint64) for cw in canon], axis=1)
    fl = np.where(valid.numpy(), 0, 2).astype(np.uint8)
    for lo, hi in ((0, n // 7), (n // 3, n // 3 + 1), (n // 3, 2 * n // 3), (n - 5, n), (n // 2, n // 2 + 2)):
        g = cb.CortexGraph.fromDevice(body[lo:hi].data_ptr(), k, c, hi - lo, firstIndex=lo, keepalive=body)
        exp = np.where((want >= lo) & (want < hi), want, -1)
        for algo in (cb.CC_ALGO_AUTO, cb.CC_ALGO_BSEARCH):
            assert (g.findPacked(pw, fl, algo) == exp).all(), (lo, hi, algo)
        assert (g.findRecordIndices(q_ascii.numpy()) == exp).all(), (lo, hi)
        g.dispose()
    whole.dispose()


@pytest.mark.parametrize("k", [21, 31, 47, 63, 95])
def test_line_index_on_clustered_keys(k):
    """The bucket-line index on key sets that are nothing like uniform: dense runs of consecutive k-mers (thousands of keys in one
    bucket: the search continues in the key column, linearly and then by bisection), keys that differ only below the top 64 bits
    (k > 32: same bin, same line), a lone key far from the rest (almost every bin empty), all fill factors and bin counts.  Every
    answer against a numpy searchsorted over the sorted keys (the oracle's order is the same unsigned word order)."""
    rng = np.random.default_rng(k)
    s = (k + 31) // 32
    top_bits = 2 * k - 64 * (s - 1)
    def rand_keys(n):
        w = rng.integers(0, 2 ** 63, size=(n, s), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, s), dtype=np.uint64)
        if top_bits < 64:
            w[:, 0] &= np.uint64((1 << top_bits) - 1)
        return w
    parts = [rand_keys(3000)]
    base = rand_keys(6)
    for b in base[:3]:                                   # runs of consecutive keys: differ in the last word only
        run = np.repeat(b[None, :], 5000, axis=0)
        run[:, -1] = (run[:, -1] & np.uint64(0xffffffffffff0000)) + np.arange(5000, dtype=np.uint64) * np.uint64(3)
        parts.append(run)
    if s > 1:                                            # same top 64 bits, different low words
        for b in base[3:]:
            run = np.repeat(b[None, :], 2000, axis=0)
            run[:, -1] = rng.integers(0, 2 ** 63, size=2000, dtype=np.uint64)
            parts.append(run)
    lone = np.zeros((1, s), dtype=np.uint64)             # the smallest possible key, far below everything else
    parts.append(lone)
    keys = np.concatenate(parts)
    order = np.lexsort([keys[:, w] for w in range(s - 1, -1, -1)])
    keys = keys[order]
    keep = np.ones(len(keys), dtype=bool)
    keep[1:] = (keys[1:] != keys[:-1]).any(axis=1)
    keys = keys[keep]
    n, c = len(keys), 1
    tw = [torch.from_numpy(keys[:, w].copy().view(np.int64)) for w in range(s)]
    cov, edges = synth.coverage_and_edges(1, n, c, "cpu", adv_period=0)
    body = synth.assemble_records(tw, cov, edges).cuda()
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
    # queries: every key, every key +-1 in the last word (mostly misses next to hits), random keys
    q = np.concatenate([keys, keys + np.array([0] * (s - 1) + [1], dtype=np.uint64), keys - np.array([0] * (s - 1) + [1], dtype=np.uint64), rand_keys(5000)])
    if top_bits < 64:
        q[:, 0] &= np.uint64((1 << top_bits) - 1)
    be = lambda a: np.ascontiguousarray(a.astype(">u8")).view(np.dtype((np.void, 8 * s))).reshape(-1)
    kb, qb = be(keys), be(q)
    pos = np.searchsorted(kb, qb)
    want = np.where((pos < n) & (kb[np.minimum(pos, n - 1)] == qb), pos, -1)
    assert (want[:n] == np.arange(n)).all()
    try:
        for fill in (50, 100, 10):
            N.set_option("index_fill_pct", fill)
            for bits in (0, 1, 13):
                g.buildIndex(bits)
                assert (g.findPacked(q) == want).all(), (fill, bits)
        assert (g.findPacked(q, algo=cb.CC_ALGO_BSEARCH) == want).all()
        assert (g.findPacked(q, algo=cb.CC_ALGO_MERGE) == want).all()            # unsorted batch: radix sort, then the merge
        qo = np.lexsort([q[:, w] for w in range(s - 1, -1, -1)])                   # the same batch in ascending order: merge only
        assert (g.findPacked(q[qo], algo=cb.CC_ALGO_MERGE) == want[qo]).all()
        sparse = q[qo][::997]                                                      # windows far longer than a tile can stage
        assert (g.findPacked(sparse, algo=cb.CC_ALGO_MERGE) == want[qo][::997]).all()
        fl = np.zeros(len(q), dtype=np.uint8)
        fl[::5] = 2
        assert (g.findPacked(q[qo], fl, algo=cb.CC_ALGO_MERGE) == np.where(fl == 0, want[qo], -1)).all()
    finally:
        N.set_option("index_fill_pct", 50)
    g.dispose()


# ------------------------------------------------------------------ next row: CortexCollection / Join

@pytest.mark.parametrize("k,colors,sizes", [(31, (1, 1), (5000, 7000)), (47, (1, 1, 1, 1), (20000, 30000, 25000, 9000)),
                                            (63, (2, 3), (4000, 4001)), (47, (4, 1), (50000, 3)), (31, (1, 2), (0, 500)),
                                            (95, (1, 1, 1), (3000, 1, 2999)), (47, (21, 5), (3000, 2500)), (127, (1, 2, 1), (2000, 3000, 10)),
                                            (31, (1, 1), (40000, 40000))])
@pytest.mark.parametrize("tiled", [1, 0])
def test_join_vs_oracle(tmp_path, k, colors, sizes, tiled):
    """cc_join / Join.execute against the oracle's union (CortexCollection.next semantics): overlapping key sets
    (every graph draws from one pool so many k-mers are shared), colour remap, absent k-mers zero-filled, header.
    tiled = 1: the shared-memory tiled union; 0: the per-thread merge-path form it falls back to for very wide records."""
    from oracle import oracle_np as onp
    N.set_option("join_tiled", tiled)
    pool = synth.random_canonical_keys(123 + k, int(max(sizes) * 1.5) + 10, k, "cpu")
    ctxs, graphs = [], []
    for gi, (c, n) in enumerate(zip(colors, sizes)):
        g = torch.Generator().manual_seed(1000 + gi)
        pick = torch.sort(torch.randperm(len(pool[0]), generator=g)[:n]).values
        words = [w[pick] for w in pool]
        cov, edges = synth.coverage_and_edges(50 + gi, n, c, "cpu", adv_period=0)
        body = synth.assemble_records(words, cov, edges) if n else torch.zeros((0, 8 * len(pool) + 5 * c), dtype=torch.uint8)
        names = ["g%d_c%d" % (gi, j) for j in range(c)]
        ctx = synth.header_bytes(k, c, names) + body.numpy().tobytes()
        ctxs.append(ctx)
        p = tmp_path / ("g%d.ctx" % gi)
        p.write_bytes(ctx)
        graphs.append(cb.CortexGraph(p))
    want = onp.join(ctxs)
    out = tmp_path / "joined.ctx"
    try:
        n = cb.Join(graphs, out).execute()
    finally:
        N.set_option("join_tiled", 1)
    got = out.read_bytes()
    hw = onp.parse_header(want)
    assert n == hw["num_records"]
    assert got == want
    # the merged view behaves like a graph: iteration order, colour names, lookups
    coll = cb.CortexCollection(graphs)
    assert coll.getNumColors() == sum(colors) and coll.getSampleName(sum(colors) - 1) == "g%d_c%d" % (len(colors) - 1, colors[-1] - 1)
    if n:
        rec = coll.getRecord(n // 2)
        assert coll.findRecord(rec.getKmerAsString()) == rec
        assert coll.getGraph(sum(colors) - 1) is graphs[-1]
    coll.merged.dispose()
    for g in graphs:
        g.dispose()


def test_join_rejects_mismatched_k(tmp_path):
    a = cb.CortexGraph(synth.make_ctx_file(1, 100, 31, 1, adv_period=0))
    b = cb.CortexGraph(synth.make_ctx_file(2, 100, 33, 1, adv_period=0))
    with pytest.raises(cb.CortexJDKException) as ei:
        cb.CortexGraph.join([a, b])
    assert "kmer sizes are not equal" in str(ei.value)
    a.dispose(); b.dispose()


@pytest.mark.parametrize("k,c,n", [(31, 2, 20000), (47, 4, 60000), (63, 21, 3000), (95, 1, 5000), (47, 4, 1), (31, 3, 0)])
def test_sort_vs_oracle(tmp_path, k, c, n):
    """cc_sort / Sort.execute on a shuffled (hash-ordered, as McCortex writes) graph: bit-exact file against the oracle's
    stable sort; the sorted graph then serves lookups, which the shuffled one refuses."""
    from oracle import oracle_np as onp
    ctx = synth.make_ctx_file(11 + k, n, k, c, adv_period=0) if n else synth.header_bytes(k, c)
    h = onp.parse_header(ctx)
    rec = onp.records_view(ctx, h)
    rng = np.random.default_rng(3)
    shuffled = ctx[:h["data_offset"]] + rec[rng.permutation(n)].tobytes()
    p = tmp_path / "raw.ctx"
    p.write_bytes(shuffled)
    raw = cb.CortexGraph(p)
    out = tmp_path / "sorted.ctx"
    assert cb.Sort(raw, out).execute() == n
    assert out.read_bytes() == onp.sort_graph(shuffled)
    if n > 100:
        with pytest.raises(cb.CortexJDKException):
            raw.findRecord("A" * k)
        sg = cb.CortexGraph(out)
        cr = sg.getRecord(n // 3)
        assert sg.findRecord(cr.getKmerAsString()) == cr
        sg.dispose()
    raw.dispose()


def test_outputs_stay_inside_their_buffers():
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are hunted with canaries: every device
    output of the scan (sparse, dense and capped), the column decode, the packer and the lookups is placed between two
    guard regions that must come back untouched."""
    k, c, n = 47, 4, 40_000
    st = torch.cuda.current_stream().cuda_stream
    L = N.lib()
    GUARD = 4096

    def guarded(nbytes):
        buf = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, buf.data_ptr() + GUARD

    def intact(buf, nbytes):
        return bool((buf[:GUARD] == 0xA5).all()) and bool((buf[GUARD + nbytes:] == 0xA5).all())

    for permille, parents in ((10, [1, 2, 3]), (10, [])):                 # sparse path / every chunk dense
        body, words = synth.make_graph_body(21, n, k, c, device="cuda", novel_permille=permille, adv_period=0)
        torch.cuda.synchronize()
        g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
        par = np.asarray(parents, dtype=np.int32)
        for cap in (n, 100, 1, 0):
            rb, rp = guarded(cap * 21)
            ib, ip = guarded(cap * 8)
            cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            N.check(L.cc_find_novel_dev(g._h, 0, par.ctypes.data if len(par) else None, len(par), rp if cap else None, ip if cap else None, cap,
                                        cnt.data_ptr(), st))
            torch.cuda.synchronize()
            assert intact(rb, cap * 21) and intact(ib, cap * 8), (permille, parents, cap)
            assert int(cnt[0]) > 0
        wb, wp = guarded(n * 16); cb_, cp = guarded(n * 16); eb, ep = guarded(n * 4)
        N.check(L.cc_decode_records_dev(g._h, 0, n, wp, cp, ep, st))
        qa, canon, valid = synth.make_queries(4, [w.cpu() for w in words], k, 70_000)
        qa = qa.cuda()
        for fn in ("ascii", "windows"):
            nq = 70_000 if fn == "ascii" else qa.numel() - k + 1
            ob, op = guarded(nq * 8)
            if fn == "ascii":
                N.check(L.cc_find_ascii_dev(g._h, qa.data_ptr(), nq, op, 0, st))
            else:
                N.check(L.cc_find_windows_dev(g._h, qa.data_ptr(), qa.numel(), op, 0, st))
            torch.cuda.synchronize()
            assert intact(ob, nq * 8), fn
        pb, pp = guarded(70_000 * 16); fb, fp = guarded(70_000)
        N.check(L.cc_pack_kmers_dev(0, qa.data_ptr(), 70_000, k, pp, fp, st))
        torch.cuda.synchronize()
        assert intact(wb, n * 16) and intact(cb_, n * 16) and intact(eb, n * 4) and intact(pb, 70_000 * 16) and intact(fb, 70_000)
        g.dispose()


# ------------------------------------------------------------------ next row: scan-shaped pre-filters / recovery

def _one_colour_graph(words, cov, edges, k, name):
    from oracle import oracle_np as onp
    s = words.shape[1]
    rec = np.zeros(len(words), dtype=onp.record_dtype(s, 1))
    rec["kmer"], rec["cov"][:, 0], rec["edges"][:, 0] = words, cov, edges
    col = dict(sample_name=name, mean_read_length=0, total_sequence=0, graph_name="", tip_clipping=0, low_covg_supernodes_removed=0,
               low_covg_kmers_removed=0, cleaned_against_graph=0, low_cov_supernodes_threshold=0, low_cov_kmer_threshold=0)
    return onp.write_header(k, s, [col]) + rec.tobytes()


@pytest.mark.parametrize("k,c,n", [(47, 4, 60_000), (31, 3, 20_000), (63, 21, 4_000), (95, 2, 3_000)])
def test_prefilters_vs_oracle(tmp_path, k, c, n):
    """FindLowCoverage, FindShared, RecoverExcludedKmers and CovStats end to end (files bit-exact against the numpy oracle)."""
    from oracle import oracle_np as onp
    ctx = synth.make_ctx_file(900 + k, n, k, c, novel_permille=40, adv_period=53)
    graph = cb.CortexGraph(ctx)
    names = [graph.getSampleName(i) for i in range(c)]
    parents = names[1:3] if c >= 3 else names[1:2]
    # ROI graph: the child's novel k-mers w.r.t. the parents only (so that some have coverage in the remaining colours)
    roi_path = tmp_path / "roi.ctx"
    cb.FindROIs(graph, parents, names[0], roi_path).execute()
    roi_bytes = roi_path.read_bytes()
    roi = cb.CortexGraph(str(roi_path))
    assert roi.getNumRecords() > 50

    # FindLowCoverage
    for m in (10, 1, 61, -5):
        kept, excluded = cb.FindLowCoverage(roi, tmp_path / "low.ctx", m).execute()
        want = onp.find_low_coverage(roi_bytes, m)
        assert (tmp_path / "low.ctx").read_bytes() == want, m
        assert excluded == onp.parse_header(want)["num_records"] and kept + excluded == roi.getNumRecords()

    # FindShared: free colours = everything but child, parents and the ignored ones
    pcols = graph.getColorsForSampleNames(parents)
    for ignore in ([], names[-1:], ["no-such-sample"]):
        cb.FindShared(graph, parents, ignore, roi, tmp_path / "shared.ctx").execute()
        icols = [x for x in graph.getColorsForSampleNames(ignore)]
        want = onp.find_shared(ctx, roi_bytes, 0, pcols, icols)
        assert (tmp_path / "shared.ctx").read_bytes() == want, ignore
    if c > 3:
        assert onp.parse_header(onp.find_shared(ctx, roi_bytes, 0, pcols, []))["num_records"] > 0
    # a ROI k-mer that is not in the graph: NullPointerException in the reference, an error here
    stray = cb.CortexGraph(_one_colour_graph(np.full((1, graph.getKmerBits()), 3, dtype=np.uint64), [1], [0], k, names[0]))
    if len(set(pcols) | {0}) < c:
        with pytest.raises(cb.CortexJDKException):
            graph.findShared(stray, 0, pcols, [])
    # ... unless every colour is excluded: then the reference never dereferences the null record (FindShared.java:67 short-circuits)
    none_shared = graph.findShared(stray, 0, pcols, list(range(c)))
    assert none_shared.getNumRecords() == 0
    none_shared.dispose()
    stray.dispose()

    # RecoverExcludedKmers: the dirty graph holds every third k-mer of the pedigree graph plus strangers, coverage 0 / small / >= 2^31
    words, cov, edges = graph.decodeRecords(0, n)
    rng = np.random.default_rng(k)
    pick = np.arange(0, n, 3)
    dcov = rng.integers(0, 4, len(pick)).astype(np.uint32)
    dcov[::17] = 0x80000005
    for child_color in (0, min(1, c - 1)):
        dirty_bytes = _one_colour_graph(words[pick], dcov, rng.integers(0, 256, len(pick)).astype(np.uint8), k, names[child_color])
        dirty = cb.CortexGraph(dirty_bytes)
        recovered = cb.RecoverExcludedKmers(graph, dirty, tmp_path / "rec.ctx").execute()
        want, nrec = onp.recover_excluded_kmers(ctx, dirty_bytes, child_color)
        assert (tmp_path / "rec.ctx").read_bytes() == want, child_color
        assert recovered == nrec and (nrec > 0 or n < 5000)
        dirty.dispose()
    with pytest.raises(cb.CortexJDKException):
        d2 = cb.CortexGraph(_one_colour_graph(words[:5], [1] * 5, [0] * 5, k, "stranger"))
        try:
            cb.RecoverExcludedKmers(graph, d2, tmp_path / "x.ctx").execute()
        finally:
            d2.dispose()

    # CovStats
    import io
    from corticall_b200 import _native as N
    for child_name, pnames in ((names[0], parents), (names[1], [names[0]]), (names[0], ["no-such-sample"])):
        ccol = graph.getColorForSampleName(child_name)
        want = onp.cov_stats(ctx, ccol, [graph.getColorForSampleName(p) for p in pnames])
        for fused in (1, 0):                     # the histogram pass, and the sort-based path it falls back to
            N.set_option("covstats_fused", fused)
            try:
                buf = io.StringIO()
                rows = cb.CovStats(graph, child_name, pnames, buf).execute()
            finally:
                N.set_option("covstats_fused", 1)
            assert rows == want and buf.getvalue() == "".join("%d\t%d\n" % r for r in want), fused
    assert len(onp.cov_stats(ctx, 0, graph.getColorsForSampleNames(parents))) > 0 or c < 4
    roi.dispose(); graph.dispose()


@pytest.mark.parametrize("k,c", [(47, 4), (31, 5)])
def test_covstats_huge_coverages(k, c):
    """CovStats is one histogram pass with three levels (shared memory < 4096, global table < 65536, overflow list beyond):
    child coverages placed in every level, on and around the boundaries, still give the oracle's table."""
    import io
    from oracle import oracle_np as onp
    n = 30_000
    names = ["kid", "mom", "dad", "ref", "other"][:c]
    ctx = bytearray(synth.make_ctx_file(77 + k, n, k, c, novel_permille=40, adv_period=53, names=names))
    hdr = len(synth.header_bytes(k, c, names))
    s_words = (k + 31) // 32
    S = 8 * s_words + 5 * c
    body = np.frombuffer(ctx, dtype=np.uint8, offset=hdr).reshape(n, S)
    cov = body[:, 8 * s_words:8 * s_words + 4 * c].copy().view("<u4").reshape(n, c)
    rng = np.random.default_rng(k)
    special = np.array([4095, 4096, 4097, 65535, 65536, 65537, 1 << 20, (1 << 30) + 7, (1 << 31) - 1, 1 << 31, 0xFFFFFFFF, 5000, 5000, 70000, 70000, 70000],
                       dtype=np.uint32)
    rows = rng.choice(n, size=len(special) * 20, replace=False)
    cov[rows, 0] = np.tile(special, 20)
    cov[rows[::2], 1] = 3          # a parent ...
    cov[rows[::3], c - 1] = 2      # ... and another sample for some of them
    body[:, 8 * s_words:8 * s_words + 4 * c] = cov.astype("<u4").view(np.uint8).reshape(n, 4 * c)
    ctx = bytes(ctx)
    graph = cb.CortexGraph(ctx)
    for child_name, pnames in (("kid", ["mom", "dad"][:max(1, c - 3)]), ("mom", ["kid"])):
        buf = io.StringIO()
        got = cb.CovStats(graph, child_name, pnames, buf).execute()
        want = onp.cov_stats(ctx, graph.getColorForSampleName(child_name), [graph.getColorForSampleName(p) for p in pnames])
        assert got == want
        if child_name == "kid":
            assert any(v >= 65536 for v, _ in want) and any(4096 <= v < 65536 for v, _ in want)
    graph.dispose()


def test_prefilters_empty_graph(tmp_path):
    ctx = synth.make_ctx_file(3, 0, 31, 3)
    g = cb.CortexGraph(ctx)
    low = g.findLowCoverage(5)
    assert low.getNumRecords() == 0
    low.dispose()
    assert g.covStats(0, [1]) == []
    rec, nrec = g.recoverExcludedKmers(g, 0)
    assert rec.getNumRecords() == 0 and nrec == 0
    rec.dispose(); g.dispose()


@pytest.mark.parametrize("k,shapes", [(31, [(2, 15000), (1, 12000), (3, 8000)]), (47, [(4, 30000), (1, 25000)]), (47, [(1, 20000)]),
                                      (63, [(21, 3000), (2, 2500), (1, 1)]), (95, [(1, 700), (2, 0), (1, 900)]), (31, [(1, 0), (1, 500)])])
def test_remove_vs_oracle(tmp_path, k, shapes):
    """cc_remove / Remove.execute (S/commands/utils/Remove.java:30-88): the primary graph minus the k-mers covered by a secondary
    graph, bit-exact file against the oracle (both restatements), including the reference's all-zero records for k-mers only a
    secondary graph holds with coverage <= 0 and adversarial coverages >= 2^31."""
    from oracle import oracle_np as onp
    pool = synth.random_canonical_keys(5 + k, int(max(n for _, n in shapes) * 1.4) + 10, k, "cpu")
    ctxs, graphs = [], []
    for gi, (c, n) in enumerate(shapes):
        g = torch.Generator().manual_seed(gi)
        pick = torch.sort(torch.randperm(len(pool[0]), generator=g)[:n]).values
        cov, edges = synth.coverage_and_edges(50 + gi, n, c, "cpu", novel_permille=100, adv_period=7)
        body = synth.assemble_records([w[pick] for w in pool], cov, edges) if n else torch.zeros((0, 8 * len(pool) + 5 * c), dtype=torch.uint8)
        ctxs.append(synth.header_bytes(k, c, ["g%d_%d" % (gi, j) for j in range(c)]) + body.numpy().tobytes())
        graphs.append(cb.CortexGraph(ctxs[-1]))
    want, removed = onp.remove(ctxs[0], ctxs[1:])
    o_rec, o_removed = orc.remove_records(orc.Graph(ctxs[0]), [orc.Graph(x) for x in ctxs[1:]])
    assert o_removed == removed and o_rec == want[onp.parse_header(want)["data_offset"]:]
    out = tmp_path / "kept.ctx"
    kept, got_removed = cb.Remove(graphs[0], graphs[1:], out).execute()
    assert got_removed == removed and kept == onp.parse_header(want)["num_records"]
    assert out.read_bytes() == want
    for g in graphs:
        g.dispose()
