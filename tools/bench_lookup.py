"""K4 experiments on one GPU (not a bench line): batched lookups on the configs[2] shape through the C ABI.
  python tools/bench_lookup.py [table_records] [queries] [k]
Prints lookups/s for packed queries (line index vs binary search), ASCII rows, windows; checks the line index against
the binary search over the key column on the whole batch (two independent device algorithms) and sweeps the index options."""
import sys
import time
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 28
k = int(sys.argv[3]) if len(sys.argv) > 3 else 47
c = 4
s = (k + 31) // 32
L = N.lib()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


table = synth.random_canonical_keys(2, nt, k, "cuda")
cov, edges = synth.coverage_and_edges(2, nt, c, "cuda")
body = synth.assemble_records(table, cov, edges)
del cov, edges
g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, nt, keepalive=body)
t0 = time.perf_counter()
g.buildIndex()
torch.cuda.synchronize()
print("index build n=%.1e k=%d: %.1f ms" % (nt, k, (time.perf_counter() - t0) * 1e3), flush=True)
chunk = 1 << 24
qw = torch.empty((nq, s), dtype=torch.int64, device="cuda")
qf = torch.empty(nq, dtype=torch.uint8, device="cuda")
na = min(nq, 1 << 26)
qa = torch.empty((na, k), dtype=torch.uint8, device="cuda")
for o in range(0, nq, chunk):
    m = min(chunk, nq - o)
    a, canon, valid = synth.make_queries(5, table, k, m, offset=o)
    for w in range(s):
        qw[o:o + m, w] = canon[w]
    qf[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
    if o < na:
        qa[o:min(o + m, na)] = a[:max(0, min(m, na - o))]
    del a, canon, valid
res = torch.empty(nq, dtype=torch.int64, device="cuda")
ref = torch.empty(nq, dtype=torch.int64, device="cuda")
bpl = 8 * s + 8 + nt * 8 * s / nq
ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nq, ref.data_ptr(), cb.CC_ALGO_BSEARCH, st)), reps=2, warm=1)
print("packed bsearch      q=%.2e  %8.3f ms  %.3g lookups/s" % (nq, ms, nq / ms * 1e3), flush=True)
for fill in (50, 40, 60, 75):
    for bits in (0, 10, 6):
        if bits and fill != 50:
            continue
        N.set_option("index_fill_pct", fill)
        g.buildIndex(bits)
        for smem in (1, 0):
            N.set_option("find_bins_smem", smem)
            ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nq, res.data_ptr(), cb.CC_ALGO_AUTO, st)))
            ok = bool(torch.equal(res, ref))
            print("packed lines fill=%d%% bins=2^%d smem=%d  q=%.2e  %8.3f ms  %.3g lookups/s  (%.0f GB/s algorithmic)  agrees_with_bsearch=%s hits=%.3f"
                  % (fill, bits, smem, nq, ms, nq / ms * 1e3, nq / ms * 1e3 * bpl / 1e9, ok, float((res >= 0).float().mean())), flush=True)
        N.set_option("find_bins_smem", 1)
N.set_option("index_fill_pct", 50)
g.buildIndex(0)
for hints in (0, 1, 2, 3):
    N.set_option("lookup_l2_hints", hints)
    ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nq, res.data_ptr(), cb.CC_ALGO_AUTO, st)))
    print("packed lines hints=%d  %8.3f ms  %.3g lookups/s" % (hints, ms, nq / ms * 1e3), flush=True)
N.set_option("lookup_l2_hints", 1)
ra = torch.empty(na, dtype=torch.int64, device="cuda")
ms = timeit(lambda: N.check(L.cc_find_ascii_dev(g._h, qa.data_ptr(), na, ra.data_ptr(), cb.CC_ALGO_AUTO, st)))
print("ascii rows (pack+find fused)  q=%.2e  %8.3f ms  %.3g lookups/s  agrees=%s" % (na, ms, na / ms * 1e3, bool(torch.equal(ra, ref[:na]))), flush=True)
pw = torch.empty((na, s), dtype=torch.int64, device="cuda")
pf = torch.empty(na, dtype=torch.uint8, device="cuda")
ms = timeit(lambda: N.check(L.cc_pack_kmers_dev(0, qa.data_ptr(), na, k, pw.data_ptr(), pf.data_ptr(), st)))
print("pack rows                     q=%.2e  %8.3f ms  %.3g rows/s  %.0f GB/s algorithmic" % (na, ms, na / ms * 1e3, na * (k + 8 * s + 1) / ms / 1e6), flush=True)
del pw, pf
nm = min(nq, 1 << 25)
ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nm, res.data_ptr(), cb.CC_ALGO_MERGE, st)), reps=2, warm=1)
print("packed sort-then-probe        q=%.2e  %8.3f ms  %.3g lookups/s  agrees=%s" % (nm, ms, nm / ms * 1e3, bool(torch.equal(res[:nm], ref[:nm]))), flush=True)
# a batch that arrives in ascending order (the records of another graph): sorted-merge against the line index on the same batch
ns = min(nq, 1 << 27)
srt = synth.sort_unique_words([qw[:ns, w].contiguous() for w in range(s)])
qs = torch.stack(srt, dim=1).contiguous()
ns = qs.shape[0]
for name, algo in (("line index", cb.CC_ALGO_AUTO), ("sorted-merge", cb.CC_ALGO_MERGE)):
    ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qs.data_ptr(), None, ns, res.data_ptr(), algo, st)))
    if algo == cb.CC_ALGO_AUTO:
        keep = res[:ns].clone()
    print("sorted batch, %-13s q=%.2e  %8.3f ms  %.3g lookups/s  agrees=%s" % (name, ns, ms, ns / ms * 1e3, bool(torch.equal(res[:ns], keep))), flush=True)
for frac in (16, 256):
    sub = qs[::frac].contiguous()
    for name, algo in (("line index", cb.CC_ALGO_AUTO), ("sorted-merge", cb.CC_ALGO_MERGE)):
        ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, sub.data_ptr(), None, sub.shape[0], res.data_ptr(), algo, st)))
        print("sorted batch 1/%d, %-13s q=%.2e  %8.3f ms  %.3g lookups/s" % (frac, name, sub.shape[0], ms, sub.shape[0] / ms * 1e3), flush=True)
del srt, qs
seq = synth.random_genome(9, 1 << 27, device="cuda")
nwin = seq.numel() - k + 1
r2 = torch.empty(nwin, dtype=torch.int64, device="cuda")
ms = timeit(lambda: N.check(L.cc_find_windows_dev(g._h, seq.data_ptr(), seq.numel(), r2.data_ptr(), cb.CC_ALGO_AUTO, st)))
print("windows of a random genome    q=%.2e  %8.3f ms  %.3g lookups/s" % (nwin, ms, nwin / ms * 1e3), flush=True)
del qw, qf, res, ref
ww = torch.empty((nwin, s), dtype=torch.int64, device="cuda")
wf = torch.empty(nwin, dtype=torch.uint8, device="cuda")
ms = timeit(lambda: N.check(L.cc_pack_canonical_dev(0, seq.data_ptr(), seq.numel(), k, ww.data_ptr(), wf.data_ptr(), st)))
print("pack windows                  q=%.2e  %8.3f ms  %.3g kmers/s" % (nwin, ms, nwin / ms * 1e3), flush=True)
