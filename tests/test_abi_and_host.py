"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every symbol include/corticall_cuda.h
declares, header parsing / error statuses mirror the reference's exceptions (CortexGraph.java:74-76,82-84,140-142,
163-167), compute entry points refuse to run without a device (no CPU fallback), and the host value types
(CanonicalKmer, CortexByteKmer, CortexBinaryKmer, CortexRecord, SequenceUtils) match the reference's golden vectors."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

import corticall_b200 as cb
from corticall_b200 import _native as N
from corticall_b200.host.kmer import _java_array_hash, decodeBinaryKmer, encodeBinaryKmer, native_words
from oracle import orc
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = cb.device_count() > 0


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "corticall_cuda.h")).read()
    return sorted(set(re.findall(r"CC_API\s+[\w\s\*]+?\b(cc_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 35
    L = N.lib()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(N.SIGNATURES) == names                       # the ctypes table binds exactly the header
    assert b"sm_100a" in L.cc_version()


def test_no_cpu_fallback(fixture_ctx):
    if HAS_GPU:
        pytest.skip("a CUDA device is present")
    with pytest.raises(cb.CortexJDKException) as ei:
        cb.CortexGraph(fixture_ctx)
    assert ei.value.status == N.CC_ERR_CUDA and "no CPU fallback" in str(ei.value)
    seq = np.frombuffer(b"ACGT" * 20, dtype=np.uint8)
    words = np.zeros((50, 1), dtype=np.uint64); flags = np.zeros(50, dtype=np.uint8)
    assert N.lib().cc_pack_canonical(0, seq.ctypes.data, seq.size, 31, words.ctypes.data, flags.ctypes.data) == N.CC_ERR_CUDA
    cnt = C.c_uint64()
    par = np.array([1], dtype=np.int32)
    rc = N.lib().cc_find_novel_host(0, seq.ctypes.data, 31, 1, 2, 1, 0, par.ctypes.data, 1, None, None, 0, C.byref(cnt), None)
    assert rc == N.CC_ERR_CUDA
    # the multi-GPU handle: every placement needs devices too; a bad placement is an argument error before anything else
    buf = np.frombuffer(fixture_ctx, dtype=np.uint8)
    devs = (C.c_int * 2)(0, 0)
    for place in (0, 1, 2):
        h = N._P()
        assert N.lib().cc_open_sharded_memory_placed(buf.ctypes.data, buf.size, devs, 2, place, C.byref(h)) == N.CC_ERR_CUDA
        assert "no CPU fallback" in N.last_error() and not h.value
    h = N._P()
    assert N.lib().cc_open_sharded_memory_placed(buf.ctypes.data, buf.size, devs, 2, 7, C.byref(h)) == N.CC_ERR_ARG
    assert N.lib().cc_open_sharded_memory(buf.ctypes.data, buf.size, devs, 2, C.byref(h)) == N.CC_ERR_CUDA


def open_status(image: bytes):
    h = N._P()
    buf = np.frombuffer(image, dtype=np.uint8)
    rc = N.lib().cc_open_memory(buf.ctypes.data, buf.size, 0, C.byref(h))
    if rc == 0:
        N.lib().cc_dispose(h)
    return rc, N.last_error()


def test_header_errors_mirror_the_reference(fixture_ctx, tmp_path):
    rc, msg = open_status(b"NOTCTX" + fixture_ctx[6:])
    assert rc == N.CC_ERR_NOT_CORTEX and "does not appear to be a Cortex graph" in msg
    rc, msg = open_status(fixture_ctx[:6] + b"\x05\0\0\0" + fixture_ctx[10:])
    assert rc == N.CC_ERR_BAD_VERSION and "is not a version 6 Cortex graph" in msg
    rc, msg = open_status(fixture_ctx[:142] + b"XORTEX" + fixture_ctx[148:])
    assert rc == N.CC_ERR_BAD_TRAILER and "proper header terminator" in msg
    rc, msg = open_status(fixture_ctx[:60])
    assert rc == N.CC_ERR_IO
    # kmer_bits that is not ceil(kmer_size / 32) (k = 31 claiming 2 words): every shift of the 2-bit arithmetic would be undefined
    rc, msg = open_status(fixture_ctx[:14] + b"\x02\0\0\0" + fixture_ctx[18:])
    assert rc == N.CC_ERR_IO and "kmer_bits 2 does not match kmer_size 31" in msg
    rc, _ = open_status(b"cortex" + fixture_ctx[6:])           # equalsIgnoreCase: header accepted, fails only for lack of a device
    assert rc == (0 if HAS_GPU else N.CC_ERR_CUDA)
    h = N._P()
    rc = N.lib().cc_open(str(tmp_path / "missing.ctx").encode(), 0, C.byref(h))
    assert rc == N.CC_ERR_IO and "not found" in N.last_error()
    assert N.lib().cc_set_option(b"no_such_option", 1) == N.CC_ERR_ARG


# ------------------------------------------------------------------ host value types vs the reference's vectors

def test_complement_and_reverse_complement(kats):               # SequenceUtilsTest :19-57
    for a, b in zip(kats["complement_in"], kats["complement_out"]):
        assert cb.SequenceUtils.complement(ord(a)) == ord(b)
    for seq, exp in kats["reverse_complement"]:
        assert cb.SequenceUtils.reverseComplement(seq) == exp
        assert cb.SequenceUtils.reverseComplement(seq.encode()) == exp.encode()


def test_lowest_orientation_matches_oracle_and_property():      # SequenceUtilsTest :59-72
    rng = random.Random(0)
    comp = str.maketrans("ACGT", "TGCA")
    for k in (21, 31, 41, 51):
        for _ in range(500):
            fw = "".join(rng.choice("ACGT") for _ in range(k))
            rc = fw.translate(comp)[::-1]
            assert cb.SequenceUtils.alphanumericallyLowestOrientation(fw) == min(fw, rc)
    for q in (b"NACGT", b"TTTTN", b"acgTT", b"AC.GT", b"AAAAA", b"ACGT", bytes([200, 65, 67, 71, 84])):
        assert cb.SequenceUtils.alphanumericallyLowestOrientation(q) == orc.lowest_orientation(q)[0]


def test_canonical_kmer(kats):                                   # CanonicalKmerTest :8-14
    a, b = (cb.CanonicalKmer(x) for x in kats["hash_collision_pair"])
    assert a.hashCode() == b.hashCode() and a != b
    ck = cb.CanonicalKmer("TTTTT")
    assert ck.getKmerAsString() == "AAAAA" and ck.isFlipped() and ck.length() == 5
    assert cb.CanonicalKmer("AAAAC") == "AAAAC" and not cb.CanonicalKmer("AAAAC").isFlipped()
    assert cb.CanonicalKmer("ACGTT").getSubKmer(1, 3).getKmerAsString() == "ACG"   # CGT -> canonical ACG
    assert [k.getKmerAsString() for k in cb.SequenceUtils.kmerizeSequence("ACGTAC", 3)] == ["ACG", "ACG", "GTA", "GTA"]
    assert len({cb.CanonicalKmer("ACGTA"), cb.CanonicalKmer("TACGT")}) == 1          # set membership (Call.java:191-197)


def test_byte_kmer_compare_is_signed():                          # CortexByteKmer.java:41-49
    a, b = cb.CortexByteKmer(bytes([200, 65])), cb.CortexByteKmer(b"AA")
    assert a.compareTo(b) == -1 and b.compareTo(a) == 1 and a.compareTo(a) == 0
    assert orc.lib().orc_byte_kmer_compare(np.frombuffer(a.getKmer(), np.uint8).ctypes.data,
                                           np.frombuffer(b.getKmer(), np.uint8).ctypes.data, 2) == -1
    assert hash(cb.CortexByteKmer("ACGT")) == _java_array_hash(b"ACGT")


def test_encode_decode_binary_kmer(fixture_ctx, kats):           # CortexGraphTest.testEncodeBinaryKmer :267-280
    og = orc.Graph(fixture_ctx)
    for i, row in enumerate(kats["fixture_records"]):
        bk, _, _ = og.get_record(i)
        assert encodeBinaryKmer(row["kmer"].encode()) == bk.tolist()          # == the record's long[] (byte-swapped disk word)
        assert decodeBinaryKmer(bk.tolist(), 31, 1) == row["kmer"].encode()
    rng = random.Random(3)
    for k in (1, 5, 31, 32, 33, 47, 63, 64, 65, 95, 128):
        km = "".join(rng.choice("ACGT") for _ in range(k)).encode()
        enc = np.zeros((k + 31) // 32, dtype=np.int64)
        assert orc.lib().orc_encode_binary_kmer(np.frombuffer(km, np.uint8).ctypes.data, k, enc.ctypes.data) == 0
        assert encodeBinaryKmer(km) == enc.tolist()
        assert decodeBinaryKmer(enc.tolist(), k, len(enc)) == km
        assert [int.from_bytes(int(x).to_bytes(8, "big", signed=True), "little") for x in enc] == native_words(km)
    with pytest.raises(RuntimeError):
        encodeBinaryKmer(b"ACGTN")
    assert cb.CortexBinaryKmer(b"TTTT").getBinaryKmer() == encodeBinaryKmer(b"AAAA")   # (byte[]) ctor canonicalises


def test_cortex_record_string_round_trip(kats):                  # CortexGraphTest.constructRecordsFromString :200-253
    for row in kats["fixture_records"]:
        text = "%s %d %d %s %s" % (row["kmer"], *row["coverage"], *row["edges"])
        cr = cb.CortexRecord.fromString(text)
        assert cr.toString() == text and cr.getKmerAsString() == row["kmer"]
        assert cr.getEdgeAsStrings() == row["edges"] and cr.getCoverages() == row["coverage"]
        for c in range(2):                                        # getLeftAndRightEdges :154-182
            es = row["edges"][c]
            assert [chr(b) for b in cr.getInEdgesAsBytes(c, False)] == [x.upper() for x in es[:4] if x != "."]
            assert [chr(b) for b in cr.getOutEdgesAsBytes(c, False)] == [x for x in es[4:] if x != "."]
            assert cr.getInDegree(c) == sum(x != "." for x in es[:4]) and cr.getOutDegree(c) == sum(x != "." for x in es[4:])
        assert cr.toString(1) == "%s %d %s" % (row["kmer"], row["coverage"][1], row["edges"][1])
    a, b = cb.CortexRecord.fromString("AAAAC 1 ....A..."), cb.CortexRecord.fromString("AAAAG 1 ....A...")
    assert a.compareTo(b) == -1 and a != b and a == cb.CortexRecord.fromString("AAAAC 1 ....A...")
    assert cb.CortexRecord.fromString("AAAAC 1 a.g..C.T").getInEdgesAsStrings(0, True) == ["T", "C"]


def test_synth_header_matches_abi_parser():
    """tools/synth writes headers the library's parser accepts with the right geometry (data offset not 16-aligned)."""
    ctx = synth.make_ctx_file(5, 100, 47, 4, adv_period=0, trailing=b"\x01\x02")
    og = orc.Graph(ctx)
    assert og.ok and og.h.num_records == 100 and og.h.record_size == 36 and og.h.data_offset % 16 != 0
    rc, _ = open_status(ctx)
    assert rc == (0 if HAS_GPU else N.CC_ERR_CUDA)              # parsed fine; only the device is missing


def test_jni_shim_matches_the_java_natives_and_compiles_against_a_stub():
    """The Java drop-in cannot be built here (no JDK): check what can be checked -- every `static native` of NativeCortex.java has
    exactly one JNI function in the shim (and vice versa), every cc_* function the shim calls is declared in the header, and the shim
    is valid C++ against a stand-in jni.h."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    java = open(os.path.join(root, "java/uk/ac/ox/well/cortexjdk/utils/io/graph/cortex/NativeCortex.java")).read()
    shim_path = os.path.join(root, "corticall_b200/csrc/jni_shim.cpp")
    shim = open(shim_path).read()
    header = open(os.path.join(root, "include/corticall_cuda.h")).read()
    natives = set(re.findall(r"static native [\w\[\]]+ (\w+)\(", java))
    shims = set(re.findall(r"^JFN\([\w ]+, (\w+)\)", shim, flags=re.M))          # definitions only, not the macro itself
    assert natives == shims and len(natives) >= 20, (natives ^ shims)
    for fn in set(re.findall(r"\b(cc_\w+)\(", shim)):
        assert re.search(r"\b%s\(" % fn, header), fn
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(root, "tests/jni_stub"), shim_path])


def _build_abi_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(str(tmp_path), "abi_example")
    libdir = os.path.join(root, "corticall_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "tools", "abi_example.c"),
                           "-L", libdir, "-lcorticall_cuda", "-Wl,-rpath," + libdir, "-o", exe])
    return exe, os.path.join(root, "tests", "golden", "two_short_contigs.ctx")


def test_c_abi_from_plain_c_fails_loudly_without_a_device(tmp_path):
    """The header is valid C11 and the library links from plain C; without a CUDA device the first compute call reports CC_ERR_CUDA
    (there is no CPU fallback)."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c", os.path.join(root, "include", "corticall_cuda.h")])
    exe, fixture = _build_abi_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the GPU test")
    p = subprocess.run([exe, fixture, "0", "1"], capture_output=True, text=True)
    assert p.returncode == 7 and "no CUDA device" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_abi_from_plain_c_on_the_fixture(tmp_path):
    """tools/abi_example.c on the reference's fixture: 66 records, 19 novel k-mers for colour 0 against colour 1 (47 the other
    way round), each found again at the index the scan reported."""
    import subprocess
    exe, fixture = _build_abi_example(tmp_path)
    for child, parent, novel in ((0, 1, 19), (1, 0, 47)):
        p = subprocess.run([exe, fixture, str(child), str(parent)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert "version 6 k 31 words 1 colours 2 records 66" in p.stdout
        assert "novel %d\n" % novel in p.stdout and "found %d of %d at the reported index" % (novel, novel) in p.stdout


@pytest.mark.gpu
def test_c_abi_sharded_from_plain_c(tmp_path):
    """The multi-GPU entry points driven from plain C (tools/abi_example.c with CC_DEVICES): one cc_sharded handle over several
    shards -- on every GPU of the box, or three shards on device 0 -- gives the single-GPU novel records and lookup indices."""
    import subprocess
    import torch
    from tools import synth
    exe, fixture = _build_abi_example(tmp_path)
    big = tmp_path / "big.ctx"
    big.write_bytes(synth.make_ctx_file(21, 300000, 47, 4, novel_permille=30))
    ndev = torch.cuda.device_count()
    lists = ["0,0,0"] + ([",".join(str(i) for i in range(ndev))] if ndev >= 2 else [])
    for path in (fixture, str(big)):
        for devs in lists:
            p = subprocess.run([exe, path, "0", "1"], capture_output=True, text=True, env=dict(os.environ, CC_DEVICES=devs))
            assert p.returncode == 0, p.stdout + p.stderr
            assert "sharded over %d devices" % len(devs.split(",")) in p.stdout and "equal the single-GPU answers" in p.stdout


@pytest.mark.gpu
def test_find_records_one_call_per_vertex():
    """cc_find_records: the legacy per-record findRecord for a vertex and its neighbours in one call (indices + record bytes),
    on the fast path (<= 64 k-mers, one launch, mapped pinned staging) and beyond it; equal to findRecordIndices + getRawRecords."""
    import numpy as np
    import torch
    import corticall_b200 as cb
    from oracle import orc
    from tools import synth
    for k, c, n in ((47, 4, 50000), (31, 2, 3000), (95, 3, 2000), (31, 45, 300)):
        ctx = synth.make_ctx_file(3 + k, n, k, c, adv_period=0)
        g = cb.CortexGraph(ctx)
        og = orc.Graph(ctx)
        words, _, _ = g.decodeRecords(0, n)
        tw = [torch.from_numpy(words[:, w].copy().view(np.int64)) for w in range(g.getKmerBits())]
        a, _, _ = synth.make_queries(8, tw, k, 400, corrupt_permille=50)
        qa = a.numpy()
        want = og.find_batch(qa)
        launches0 = cb.launch_count()
        for nq in (1, 9, 64, 65, 400):
            idx, raw = g.findRecords(qa[:nq])
            assert idx.tolist() == want[:nq].tolist(), (k, nq)
            for i in range(nq):
                if idx[i] >= 0:
                    assert (raw[i] == g.getRawRecords(int(idx[i]), 1)[0]).all()
                else:
                    assert not raw[i].any()
        assert g.findRecord(qa[0].tobytes()) == (g.getRecord(int(want[0])) if want[0] >= 0 else None)
        if c < 100:
            lc = cb.launch_count()
            g.findRecords(qa[:9])
            assert cb.launch_count() - lc == 1          # one kernel for the vertex and its eight neighbours
        g.dispose()
