"""Single-GPU timings of the lookup path's kernels (GPU box only):
  * K3 rows: cc_pack_kmers_dev, cc_find_ascii_dev fused vs pack-then-search
  * K4 on a k-mer-range shard of 1/world of the table (what each rank searches at `world` GPUs)
  * the three routed legs (route / search / gather) of ONE rank with every rank emulated on this device: peer pointers are
    local, so the times are the kernels' own cost without the NVLink transfer.
usage: python tools/bench_lookup_legs.py [world] [table_records] [queries_per_rank]"""
import sys
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from corticall_b200.host.sharded import RoutedLookup
from tools import synth

L = N.lib()
st = torch.cuda.current_stream().cuda_stream
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nt = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
nq = int(float(sys.argv[3])) if len(sys.argv) > 3 else 125_000_000
vsub = int(sys.argv[4]) if len(sys.argv) > 4 else 1
K, C = 47, 4
dev = torch.device("cuda", 0)


import os
REPS, WARM = int(os.environ.get("LEGS_REPS", "5")), int(os.environ.get("LEGS_WARM", "2"))
PROFILE = os.environ.get("LEGS_PROFILE") == "1"     # under `ncu --profile-from-start off`: one marked launch per kernel


def timeit(fn, reps=None, warm=None, prof=False):
    reps, warm = reps or REPS, WARM if warm is None else warm
    if PROFILE and prof:
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
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


words = synth.random_canonical_keys(20261019, nt, K, dev)
cov, edges = synth.coverage_and_edges(20261019, nt, C, dev)
body = synth.assemble_records(words, cov, edges)
del cov, edges
whole = cb.CortexGraph.fromDevice(body.data_ptr(), K, C, nt, keepalive=body)
whole.buildIndex()

# ---- rows: pack / ascii lookups
na = 1 << 25
a, canon, valid = synth.make_queries(5, words, K, na)
pw = torch.empty(na * 2, dtype=torch.int64, device=dev); pf = torch.empty(na, dtype=torch.uint8, device=dev)
res = torch.empty(max(na, nq), dtype=torch.int64, device=dev)
ms = timeit(lambda: N.check(L.cc_pack_kmers_dev(0, a.data_ptr(), na, K, pw.data_ptr(), pf.data_ptr(), st)), prof=True)
print("pack rows k=47            q=%.2e  %.3f ms  %.3g rows/s  %6.0f GB/s (k+8s+1 B/row)" % (na, ms, na / ms * 1e3, na * (K + 17) / ms / 1e6), flush=True)
want_w = torch.stack(canon, dim=1).contiguous()
okrows = valid
assert torch.equal(pw.view(na, 2)[okrows], want_w[okrows]), "pack rows mismatch"
for fused in (1, 0):
    N.set_option("rows_fused", fused)
    ms = timeit(lambda: N.check(L.cc_find_ascii_dev(whole._h, a.data_ptr(), na, res.data_ptr(), 0, st)), prof=bool(fused))
    print("find ascii rows fused=%d   q=%.2e  %.3f ms  %.3g lookups/s" % (fused, na, ms, na / ms * 1e3), flush=True)
    if fused:
        r1 = res[:na].clone()
    else:
        assert torch.equal(r1, res[:na]), "fused and two-pass ascii lookups disagree"
N.set_option("rows_fused", 1)
qw1 = want_w; qf1 = torch.where(valid, 0, 2).to(torch.uint8)
ms = timeit(lambda: N.check(L.cc_find_packed_dev(whole._h, qw1.data_ptr(), qf1.data_ptr(), na, res.data_ptr(), 0, st)), prof=True)
print("find packed whole table   q=%.2e  %.3f ms  %.3g lookups/s" % (na, ms, na / ms * 1e3), flush=True)
assert torch.equal(r1, res[:na]), "ascii and packed lookups disagree"
if os.environ.get("LEGS_SWEEP") == "1":
    for hints in (0, 1, 2, 3):
        N.set_option("lookup_l2_hints", hints)
        ms = timeit(lambda: N.check(L.cc_find_packed_dev(whole._h, qw1.data_ptr(), qf1.data_ptr(), na, res.data_ptr(), 0, st)))
        ms2 = timeit(lambda: N.check(L.cc_find_ascii_dev(whole._h, a.data_ptr(), na, res.data_ptr(), 0, st)))
        print("   whole table, L2 hints %d: packed %.3f ms  %.3g lookups/s; ascii %.3f ms  %.3g lookups/s" % (hints, ms, na / ms * 1e3, ms2, na / ms2 * 1e3), flush=True)
    N.set_option("lookup_l2_hints", 3)
del a, canon, valid, pw, pf, want_w, qw1, qf1, r1

# ---- shards + routed legs, all ranks on this device
first = [nt * r // world for r in range(world)]
shards = []
for r in range(world):
    lo, hi = first[r], nt * (r + 1) // world
    g = cb.CortexGraph.fromDevice(body[lo:hi].data_ptr(), K, C, hi - lo, firstIndex=lo, keepalive=body)
    g.buildIndex()
    shards.append(g)
splitters = RoutedLookup.virtual_splitters(None, 0, world, vsub, dev, emulate_graphs=shards)
cap = int(nq / (world * vsub) * 1.15) + 4096 if world * vsub > 1 else nq
blocks = [torch.zeros(RoutedLookup.block_elems(world, cap, K, vsub), dtype=torch.int64, device=dev) for _ in range(world)]
rls = [RoutedLookup(shards[r], splitters, r, world, dev, cap, K, shard_first=first, emulate=blocks, max_batch=nq, vsub=vsub) for r in range(world)]
qs = []
chunk = 1 << 24
for r in range(world):
    qw = torch.empty((nq, 2), dtype=torch.int64, device=dev); qf = torch.empty(nq, dtype=torch.uint8, device=dev)
    for o in range(0, nq, chunk):
        m = min(chunk, nq - o)
        _, canon, valid = synth.make_queries(20261019, words, K, m, offset=r * nq + o)
        qw[o:o + m, 0], qw[o:o + m, 1] = canon[0], canon[1]
        qf[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
    qs.append((qw, qf))
    if r > 0:
        rls[r].route(qw, qf)          # fills rank 0's inbox segment r
torch.cuda.synchronize()
ms_r = timeit(lambda: rls[0].route(qs[0][0], qs[0][1]), prof=True)
for r in range(1, world):
    rls[r].search()                   # fills rank 0's return segments
ms_s = timeit(lambda: rls[0].search(), prof=True)
out = res[:nq]
ms_g = timeit(lambda: rls[0].gather(out), prof=True)
print("routed legs of one rank, world=%d, vsub=%d, %d queries per rank, table %.1e (shard %.2e records):" % (world, vsub, nq, nt, nt / world))
print("   route  %.3f ms  (%.3g q/s)\n   search %.3f ms  (%.3g q/s)\n   gather %.3f ms  (%.3g q/s)" % (
    ms_r, nq / ms_r * 1e3, ms_s, nq / ms_s * 1e3, ms_g, nq / ms_g * 1e3), flush=True)
print("   sum %.3f ms -> %.3g lookups/s per rank, x%d ranks = %.3g (no NVLink time)" % (
    ms_r + ms_s + ms_g, nq / (ms_r + ms_s + ms_g) * 1e3, world, world * nq / (ms_r + ms_s + ms_g) * 1e3))
chk = torch.empty(nq, dtype=torch.int64, device=dev)
N.check(L.cc_find_packed_dev(whole._h, qs[0][0].data_ptr(), qs[0][1].data_ptr(), nq, chk.data_ptr(), 0, st))
torch.cuda.synchronize()
assert torch.equal(chk, out), "routed result differs from the whole-table lookup"
ms = timeit(lambda: N.check(L.cc_find_packed_dev(shards[0]._h, qs[0][0].data_ptr(), qs[0][1].data_ptr(), nq, chk.data_ptr(), 0, st)))
print("find packed on shard 0 only (all of rank 0's queries, %.0f%% in range)  %.3f ms  %.3g lookups/s" % (100.0 / world, ms, nq / ms * 1e3))
for rl in rls:
    rl.check_overflow()
print("routed == whole-table lookup: ok")
if os.environ.get("LEGS_SWEEP") == "1":
    for hints in (0, 1, 2, 3):
        N.set_option("lookup_l2_hints", hints)
        for nb in (8_388_608, 12_500_000, 16_777_216, 25_000_000):
            N.set_option("index_buckets", nb)
            shards[0].buildIndex()
            ms_s = timeit(lambda: rls[0].search())
            print("   shard 0 search, L2 hints %d, %.3g buckets (%.2f keys/bucket, table %.0f MB): %.3f ms (%.3g q/s)" % (
                hints, nb, nt / world / nb, nb * 4 / 1e6, ms_s, nq / ms_s * 1e3), flush=True)
    N.set_option("index_buckets", 0)
    N.set_option("lookup_l2_hints", 3)
