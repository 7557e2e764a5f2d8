"""Runs one kernel of the path a few times on its bench shape (for ncu captures / quick timing).  GPU box only.
  python tools/run_once.py scan|scan5|lookup|ascii|rows|windows|packwin [reps] [option=value ...]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

what = sys.argv[1] if len(sys.argv) > 1 else "scan"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
NT = 100_000_000
for kv in sys.argv[3:]:
    name, val = kv.split("=")
    if name == "nt":
        NT = int(float(val))
    else:
        N.set_option(name, int(val))
L = N.lib()
st = torch.cuda.current_stream().cuda_stream


def timed(fn):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    return ["%.3f" % e[i].elapsed_time(e[i + 1]) for i in range(reps)]


if what in ("scan", "scan5"):
    k, c, n = (47, 4, 25_000_000) if what == "scan" else (63, 21, 20_000_000)
    s = (k + 31) // 32
    body, _ = synth.make_graph_body(20261018, n, k, c, device="cuda")
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
    parents = np.arange(1, c, dtype=np.int32)
    cap = n // 8
    out = torch.empty(cap * (8 * s + 5) + 64, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    ms = timed(lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, c - 1, out.data_ptr(), None, cap, cnt.data_ptr(), st)))
    print(what, "n", n, "novel", int(cnt[0]), "ms", ms)
else:
    k, c = 47, 4
    nt, nq = NT, 1 << 25
    table = synth.random_canonical_keys(2, nt, k, "cuda")
    cov, edges = synth.coverage_and_edges(2, nt, c, "cuda")
    body = synth.assemble_records(table, cov, edges)
    del cov, edges
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, nt, keepalive=body)
    g.buildIndex()
    a, canon, valid = synth.make_queries(5, table, k, nq)
    qw = torch.stack(canon, dim=1).contiguous()
    qf = torch.where(valid, 0, 2).to(torch.uint8)
    res = torch.empty(nq, dtype=torch.int64, device="cuda")
    if what == "lookup":
        ms = timed(lambda: N.check(L.cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nq, res.data_ptr(), 0, st)))
    elif what == "ascii":
        ms = timed(lambda: N.check(L.cc_find_ascii_dev(g._h, a.data_ptr(), nq, res.data_ptr(), 0, st)))
    elif what == "rows":
        ms = timed(lambda: N.check(L.cc_pack_kmers_dev(0, a.data_ptr(), nq, k, qw.data_ptr(), qf.data_ptr(), st)))
    else:
        seq = synth.random_genome(9, nq + k - 1, device="cuda")
        if what == "windows":
            ms = timed(lambda: N.check(L.cc_find_windows_dev(g._h, seq.data_ptr(), seq.numel(), res.data_ptr(), 0, st)))
        else:
            ms = timed(lambda: N.check(L.cc_pack_canonical_dev(0, seq.data_ptr(), seq.numel(), k, qw.data_ptr(), qf.data_ptr(), st)))
    print(what, "table", nt, "queries", nq, "ms", ms, "per_s %.3g" % (nq / float(ms[-1]) * 1e3))
