"""Experiment: lookup rate when the searched slice of the table (keys + prefix table) fits L2, as it would after a
second-level partition of the batch by key range.  usage: python tools/exp_l2_resident.py"""
import sys
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

L = N.lib(); st = torch.cuda.current_stream().cuda_stream
K, C, nt = 47, 4, 100_000_000
dev = torch.device("cuda", 0)
words = synth.random_canonical_keys(20261019, nt, K, dev)
cov, edges = synth.coverage_and_edges(20261019, nt, C, dev)
body = synth.assemble_records(words, cov, edges); del cov, edges
nq0 = 1 << 27
_, canon, valid = synth.make_queries(5, words, K, nq0)
qw = torch.stack(canon, dim=1).contiguous()


def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for parts in (8, 16, 32, 64, 128, 256):
    lo, hi = nt // 3, nt // 3 + nt // parts
    g = cb.CortexGraph.fromDevice(body[lo:hi].data_ptr(), K, C, hi - lo, firstIndex=lo, keepalive=body)
    g.buildIndex()
    k0 = (words[0][lo], words[1][lo]); k1 = (words[0][hi - 1], words[1][hi - 1])
    ge = (qw[:, 0] > k0[0]) | ((qw[:, 0] == k0[0]) & (qw[:, 1] >= k0[1]))
    le = (qw[:, 0] < k1[0]) | ((qw[:, 0] == k1[0]) & (qw[:, 1] <= k1[1]))
    sel = qw[ge & le & valid].contiguous()
    reps = max(1, (1 << 24) // max(1, sel.shape[0]))
    sel = sel.repeat(reps, 1)[torch.randperm(sel.shape[0] * reps, device=dev)].contiguous()     # >= 16M in-range queries, shuffled
    m = sel.shape[0]
    res = torch.empty(m, dtype=torch.int64, device=dev)
    for hints in (0, 3):
        N.set_option("lookup_l2_hints", hints)
        ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, sel.data_ptr(), None, m, res.data_ptr(), 0, st)))
        print("slice 1/%d: %.2e keys (%.0f MB keys + %.0f MB table), %d in-range queries, hints %d: %.3f ms  %.3g lookups/s  hits %.2f" % (
            parts, hi - lo, (hi - lo) * 16 / 1e6, (hi - lo) * 4 / 1e6, m, hints, ms, m / ms * 1e3, float((res >= 0).float().mean())), flush=True)
    g.dispose()
