"""Times the scan-shaped pre-filters on a configs[1]-sized pedigree graph (GPU box only)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from tools import synth

k, c, n = 47, 4, 25_000_000
body, words = synth.make_graph_body(20261018, n, k, c, device="cuda")
g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
S = 8 * 2 + 5 * c


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


ms, rows = timed(lambda: g.covStats(0, [1, 2]))
print("CovStats                n=%.1e  %.2f ms  %.3g records/s  (%d rows)" % (n, ms, n / ms * 1e3, len(rows)))
# dirty graph: every 4th k-mer, one colour
pick = torch.arange(0, n, 4, device="cuda")
dcov = torch.randint(0, 5, (len(pick), 1), device="cuda", dtype=torch.int32)
dedg = torch.zeros((len(pick), 1), dtype=torch.uint8, device="cuda")
dbody = synth.assemble_records([w[pick] for w in words], dcov, dedg)
dirty = cb.CortexGraph.fromDevice(dbody.data_ptr(), k, 1, len(pick), keepalive=dbody)
dirty.buildIndex()


def rec():
    out, nrec = g.recoverExcludedKmers(dirty, 0)
    m = out.getNumRecords()
    out.dispose()
    return m, nrec


ms, (m, nrec) = timed(rec)
print("RecoverExcludedKmers    n=%.1e  %.2f ms  %.3g records/s  (%d written, %d recovered; %.0f MB in, %.0f MB out)" % (
    n, ms, n / ms * 1e3, m, nrec, n * S / 1e6, m * 21 / 1e6))
cnt, recs, idx = g.findNovel(0, [1, 2])
print("novel vs parents only: %d" % cnt)

# ---- Join: 4 single-colour graphs of ~2.5e7 records drawn from one pool of 3.2e7 k-mers -> one 4-colour graph
from corticall_b200 import _native as N
del dirty, dbody
torch.cuda.empty_cache()
pool = synth.random_canonical_keys(77, 32_000_000, k, "cuda")
parts, keep = [], []
for gi in range(4):
    m = (synth.umod(synth.hash_idx(900 + gi, 1, torch.arange(len(pool[0]), device="cuda")), 1000) < 780)
    w = [t[m] for t in pool]
    cv, ed = synth.coverage_and_edges(60 + gi, len(w[0]), 1, "cuda", adv_period=0)
    b = synth.assemble_records(w, cv, ed)
    gg = cb.CortexGraph.fromDevice(b.data_ptr(), k, 1, len(w[0]), keepalive=b)
    gg.buildIndex()
    parts.append(gg)
nin = sum(p.getNumRecords() for p in parts)
for tiled, kb in ((1, 24), (1, 48), (1, 100), (0, 0), (1, 48)):
    N.set_option("join_tiled", tiled)
    if kb:
        N.set_option("join_tile_kb", kb)

    def jn():
        out = cb.CortexGraph.join(parts)
        m = out.getNumRecords()
        out.dispose()
        return m

    ms, m = timed(jn)
    print("Join (tiled=%d, %d KB)  4 x 1 colour, %.2e input records -> %.2e records  %.2f ms  %.3g input records/s  (%.0f MB in, %.0f MB out)" % (
        tiled, kb, nin, m, ms, nin / ms * 1e3, nin * 21 / 1e6, m * 36 / 1e6))
N.set_option("join_tiled", 1)
N.set_option("join_tile_kb", 48)
