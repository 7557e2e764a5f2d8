"""Times cc_join (Join / CortexCollection) call by call: 4 single-colour graphs of ~2.5e7 records drawn from one pool of
3.2e7 k-mers -> one 4-colour graph (GPU box only).  python tools/bench_join.py [reps] [option=value ...]"""
import sys, time
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for kv in sys.argv[2:]:
    name, val = kv.split("=")
    N.set_option(name, int(val))
k = 47
pool = synth.random_canonical_keys(77, 32_000_000, k, "cuda")
parts = []
for gi in range(4):
    m = (synth.umod(synth.hash_idx(900 + gi, 1, torch.arange(len(pool[0]), device="cuda")), 1000) < 780)
    w = [t[m] for t in pool]
    cv, ed = synth.coverage_and_edges(60 + gi, len(w[0]), 1, "cuda", adv_period=0)
    b = synth.assemble_records(w, cv, ed)
    gg = cb.CortexGraph.fromDevice(b.data_ptr(), k, 1, len(w[0]), keepalive=b)
    gg.buildIndex()
    parts.append(gg)
del pool
torch.cuda.empty_cache()
nin = sum(p.getNumRecords() for p in parts)
times = []
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = cb.CortexGraph.join(parts)
    t1 = time.perf_counter()
    m = out.getNumRecords()
    out.dispose()
    t2 = time.perf_counter()
    times.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3))
print("join of 4 x 1 colour, %.3g input records -> %.3g records (%.0f MB in, %.0f MB out)" % (nin, m, nin * 21 / 1e6, m * 36 / 1e6))
print("per call ms (join, dispose):", " ".join("%.2f/%.2f" % t for t in times))
best = min(t[0] for t in times)
print("best %.2f ms = %.3g input records/s" % (best, nin / best * 1e3))
