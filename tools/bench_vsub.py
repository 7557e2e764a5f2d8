"""Single-GPU proxy for one rank of an N-GPU routed lookup (GPU box only): a table of table/world records is searched for
queries_per_rank queries routed through `vsub` virtual shards (world = 1: route, search and gather are all local), so the
legs' own costs and the effect of L2-sized sub-ranges on the search leg can be read without NVLink in the way.
usage: python tools/bench_vsub.py [shard_records] [queries] [vsub ...]"""
import sys
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from corticall_b200.host.sharded import RoutedLookup
from tools import synth

nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 125_000_000
vsubs = [int(x) for x in sys.argv[3:]] or [1, 2, 4, 8, 16, 32]
K, C = 47, 4
dev = torch.device("cuda", 0)
words = synth.random_canonical_keys(20261019, nt, K, dev)
cov, edges = synth.coverage_and_edges(20261019, nt, C, dev)
body = synth.assemble_records(words, cov, edges)
del cov, edges
g = cb.CortexGraph.fromDevice(body.data_ptr(), K, C, nt, keepalive=body)
g.buildIndex()
qw = torch.empty((nq, 2), dtype=torch.int64, device=dev)
qf = torch.empty(nq, dtype=torch.uint8, device=dev)
for o in range(0, nq, 1 << 24):
    m = min(1 << 24, nq - o)
    _, canon, valid = synth.make_queries(5, words, K, m, offset=o)
    qw[o:o + m, 0], qw[o:o + m, 1] = canon[0], canon[1]
    qf[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
ref = torch.empty(nq, dtype=torch.int64, device=dev)
N.check(N.lib().cc_find_packed_dev(g._h, qw.data_ptr(), qf.data_ptr(), nq, ref.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
out = torch.empty_like(ref)
for hints in (0, 1):
    N.set_option("lookup_l2_hints", hints)
    for vsub in vsubs:
        spl = RoutedLookup.virtual_splitters(g, 0, 1, vsub, dev)
        blocks = [torch.zeros(RoutedLookup.block_elems(1, nq, K, vsub), dtype=torch.int64, device=dev)]
        rl = RoutedLookup(g, spl, 0, 1, dev, nq // vsub * 5 // 4 + 4096 if vsub > 1 else nq, K, shard_first=[0], emulate=blocks, max_batch=nq, vsub=vsub)
        for _ in range(2):
            rl.find_packed(qw, qf, out)
        tot = {}
        for _ in range(3):
            rl.find_packed(qw, qf, out, profile=True)
            for kk, v in rl.phase_ms.items():
                tot[kk] = tot.get(kk, 0.0) + v / 3
        print("hints=%d vsub=%2d  route %.3f  search %.3f  gather %.3f  total %.3f ms  (%.3g lookups/s)  ok=%s" % (
            hints, vsub, tot["route"], tot["search"], tot["gather"], sum(tot.values()), nq / sum(tot.values()) * 1e3, bool(torch.equal(out, ref))), flush=True)
        del rl, blocks
        torch.cuda.empty_cache()
