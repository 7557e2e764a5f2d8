"""Times every kernel of the path on the shapes BASELINE.json names (GPU box only): scan per shape, column decode,
pack/canonicalise (windows and rows), lookups (bucketed / bsearch / merge, packed and ASCII)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def timeit(fn, reps=10, warm=2):
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


def scan_shape(name, k, c, n, permille=5):
    s = (k + 31) // 32
    S, O = 8 * s + 5 * c, 8 * s + 5
    body, _ = synth.make_graph_body(1, n, k, c, device="cuda", novel_permille=permille)
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
    parents = np.arange(1, c, dtype=np.int32)
    cap = max(n // 8, 1 << 20)
    out = torch.empty(cap * O + 64, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, len(parents), out.data_ptr(), None, cap, cnt.data_ptr(), st)))
    print("scan   %-28s S=%3d n=%.1e  %.3f ms  %6.0f GB/s  %.3g rec/s  novel=%d" % (name, S, n, ms, n * S / ms / 1e6, n / ms * 1e3, int(cnt[0])), flush=True)
    # column decode (keys only = index build; all columns)
    words = torch.empty(n * s, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: N.check(L.cc_decode_records_dev(g._h, 0, n, words.data_ptr(), None, None, st)), reps=5)
    print("decode %-28s keys only           %.3f ms  %6.0f GB/s (read+write)" % (name, ms, n * (S + 8 * s) / ms / 1e6), flush=True)
    g.dispose()


scan_shape("cfg2 trio k47 c4", 47, 4, 25_000_000)
scan_shape("cfg4-slice k31 c4", 31, 4, 30_000_000 if not quick else 10_000_000)
scan_shape("cfg5 k63 c21 (20 parents)", 63, 21, 8_000_000 if not quick else 3_000_000)
scan_shape("k31 c1 (13 B records)", 31, 1, 40_000_000 if not quick else 10_000_000)
scan_shape("k95 c3 (39 B, odd)", 95, 3, 20_000_000 if not quick else 5_000_000)

# ---- pack / canonicalise
for k in (31, 47, 63):
    s = (k + 31) // 32
    ln = 1 << 28 if not quick else 1 << 26
    seq = synth.random_genome(3, ln, device="cuda")
    nw = ln - k + 1
    w = torch.empty(nw * s, dtype=torch.int64, device="cuda"); f = torch.empty(nw, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: N.check(L.cc_pack_canonical_dev(0, seq.data_ptr(), ln, k, w.data_ptr(), f.data_ptr(), st)), reps=5)
    print("pack   windows k=%d  n=%.2e  %.3f ms  %.3g kmers/s  %6.0f GB/s (1+8s+1 B/kmer)" % (k, nw, ms, nw / ms * 1e3, nw * (2 + 8 * s) / ms / 1e6), flush=True)
    del seq, w, f

# ---- lookups
k, c = 47, 4
nt = 100_000_000 if not quick else 20_000_000
nq = 1 << 27 if not quick else 1 << 25
table = synth.random_canonical_keys(2, nt, k, "cuda")
cov, edges = synth.coverage_and_edges(2, nt, c, "cuda")
body = synth.assemble_records(table, cov, edges); del cov, edges
g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, nt, keepalive=body)
ms = timeit(lambda: g.buildIndex(), reps=2, warm=1)
print("index  build (decode keys + validate + table) n=%.1e  %.3f ms" % (nt, ms), flush=True)
a, canon, valid = synth.make_queries(5, table, k, nq)
qw = torch.stack(canon, dim=1).contiguous(); qf = torch.where(valid, 0, 2).to(torch.uint8)
res = torch.empty(nq, dtype=torch.int64, device="cuda")
for bits in (0, 24, 26, 27, 28, 29):
    if bits:
        g.buildIndex(bits)
    for name, algo, qpt in (("bucketed", 0, 0), ("bucketed", 0, 1), ("bucketed", 0, 2), ("bucketed", 0, 4), ("bsearch", 1, 0)):
        if bits and algo:
            continue
        N.set_option("lookup_queries_per_thread", qpt)
        ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nq, res.data_ptr(), algo, st)), reps=5)
        print("lookup packed %-9s bits=%2d qpt=%d n=%.1e q=%.1e  %.3f ms  %.3g lookups/s" % (name, bits, qpt, nt, nq, ms, nq / ms * 1e3), flush=True)
N.set_option("lookup_queries_per_thread", 2)
g.buildIndex(0)
pw = torch.empty(nq * 2, dtype=torch.int64, device="cuda"); pf = torch.empty(nq, dtype=torch.uint8, device="cuda")
ms = timeit(lambda: N.check(L.cc_pack_kmers_dev(0, a.data_ptr(), nq, k, pw.data_ptr(), pf.data_ptr(), st)), reps=5)
print("pack   rows k=47 (query list)              q=%.1e  %.3f ms  %.3g rows/s  %6.0f GB/s (k+8s+1 B/row)" % (nq, ms, nq / ms * 1e3, nq * (k + 17) / ms / 1e6), flush=True)
del pw, pf
ms = timeit(lambda: N.check(L.cc_find_ascii_dev(g._h, a.data_ptr(), nq, res.data_ptr(), 0, st)), reps=5)
print("lookup ascii rows (pack+find fused)       q=%.1e  %.3f ms  %.3g lookups/s" % (nq, ms, nq / ms * 1e3), flush=True)
nm = nq // 4
ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nm, res.data_ptr(), 2, st)), reps=3)
print("lookup packed sorted-merge                q=%.1e  %.3f ms  %.3g lookups/s" % (nm, ms, nm / ms * 1e3), flush=True)
# sliding windows over a genome drawn from the table's k-mers is not available for random tables; use a random genome (all misses)
seq = synth.random_genome(9, 1 << 27 if not quick else 1 << 25, device="cuda")
nwin = seq.numel() - k + 1
res2 = torch.empty(nwin, dtype=torch.int64, device="cuda")
ms = timeit(lambda: N.check(L.cc_find_windows_dev(g._h, seq.data_ptr(), seq.numel(), res2.data_ptr(), 0, st)), reps=5)
print("lookup windows of a genome (all miss)      q=%.1e  %.3f ms  %.3g lookups/s" % (nwin, ms, nwin / ms * 1e3), flush=True)

# ---- join (CortexCollection / Join): 4 single-colour graphs drawn from one pool -> 4-colour trio graph
del g, body, table, qw, qf, res, a, seq, res2
torch.cuda.empty_cache()
nj = 25_000_000 if not quick else 5_000_000
pool = synth.random_canonical_keys(77, int(nj * 1.3), 47, "cuda")
gs, keep = [], []
for gi in range(4):
    gen = torch.Generator(device="cuda").manual_seed(gi)
    pick = torch.sort(torch.randperm(len(pool[0]), generator=gen, device="cuda")[:nj]).values
    cov, edges = synth.coverage_and_edges(gi, nj, 1, "cuda", adv_period=0)
    b = synth.assemble_records([w[pick] for w in pool], cov, edges)
    keep.append(b)
    gg = cb.CortexGraph.fromDevice(b.data_ptr(), 47, 1, nj, keepalive=b)
    gg.buildIndex()
    gs.append(gg)
torch.cuda.synchronize()
import time
for _ in range(2):
    t0 = time.perf_counter()
    j = cb.CortexGraph.join(gs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nout = j.getNumRecords()
    j.dispose()
inb = 4 * nj * 21
print("join   4 x %.1e single-colour k=47 graphs -> %.3e records x 36 B   %.2f ms   %.3g input records/s  %.0f GB/s (in+out bytes)" % (
    nj, nout, dt * 1e3, 4 * nj / dt, (inb + nout * 36) / dt / 1e9), flush=True)
