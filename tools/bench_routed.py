"""Routed (k-mer-range sharded) lookups on N GPUs, experiments (not a bench line):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_routed.py [table] [queries_total]
Times the plain route -> search -> gather sequence and the pipelined form (legs of different sub-batches at the same
time) for several sub-batch counts and CTA budgets; every variant must reproduce the sequential answers."""
import json
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
import corticall_b200 as cb
from corticall_b200 import _native as N
from corticall_b200.host.sharded import PipelinedRoutedLookup, RoutedLookup
from tools import synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq_total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
K, C_ = 47, 4
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = bench.Env(rank, world, local)
dev = env.dev
words = synth.random_canonical_keys(bench.SEED_LOOKUP, nt, K, dev)
lo, hi = nt * rank // world, nt * (rank + 1) // world
cov, edges = synth.coverage_and_edges(bench.SEED_LOOKUP, hi - lo, C_, dev, offset=lo)
body = synth.assemble_records([w[lo:hi] for w in words], cov, edges)
del cov, edges
splitters = torch.stack([torch.stack([w[nt * r // world] for w in words]) for r in range(1, world)]) if world > 1 else None
g = cb.CortexGraph.fromDevice(body.data_ptr(), K, C_, hi - lo, firstIndex=lo, device=local, keepalive=body)
g.buildIndex()
nq = nq_total // world
qw = torch.empty((nq, 2), dtype=torch.int64, device=dev)
qf = torch.empty(nq, dtype=torch.uint8, device=dev)
for o in range(0, nq, 1 << 24):
    m = min(1 << 24, nq - o)
    _, canon, valid = synth.make_queries(bench.SEED_LOOKUP, words, K, m, offset=rank * nq + o)
    qw[o:o + m, 0], qw[o:o + m, 1] = canon[0], canon[1]
    qf[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
del words
res = torch.empty(nq, dtype=torch.int64, device=dev)
ref = None
for name, opts in (("sequential", {}), ("sequential, route staging 3 deep", {"route_stage_depth": 3}), ("sequential, route staging 4 deep", {"route_stage_depth": 4}),
                   ("sequential, route staging 3 deep, 4 CTAs/SM", {"route_stage_depth": 3, "route_blocks_per_sm": 4})):
    for kname, v in opts.items():
        N.set_option(kname, v)
    rl = RoutedLookup(g, splitters, rank, world, dev, cap=int(nq / world * 1.25) + 4096, k=K, max_batch=nq)
    ms = env.timeit(lambda: rl.find_packed(qw, qf, res))
    rl.find_packed(qw, qf, res, profile=True)
    allp = [None] * world
    if world > 1:
        dist.all_gather_object(allp, {k: round(v, 3) for k, v in rl.phase_ms.items()})
    else:
        allp = [rl.phase_ms]
    if ref is None:
        ref = res.clone()
    if rank == 0:
        print(json.dumps({"variant": name, "n_gpus": world, "lookups_per_s": nq * world / ms * 1e3, "ms": ms, "equals_first": bool(torch.equal(res, ref)),
                          "search_ms_all_ranks": [p["search"] for p in allp], "route_ms_all_ranks": [p["route"] for p in allp],
                          "gather_ms_all_ranks": [p["gather"] for p in allp]}), flush=True)
    del rl
    N.set_option("lookup_l2_hints", -1)
    N.set_option("routed_search_blocks_per_sm", 0)
    N.set_option("route_blocks_per_sm", 0)
    N.set_option("route_stage_depth", 2)
    torch.cuda.empty_cache()
if len(sys.argv) > 3 and sys.argv[3] == "pipelines":
    res2 = torch.empty_like(res)
    for nsub, per in ((4, (3, 1, 2)), (4, (2, 2, 2))):
        pl = PipelinedRoutedLookup(g, splitters, rank, world, dev, sub_batch=(nq + nsub - 1) // nsub, k=K, per_sm=per)
        msp = env.timeit(lambda: pl.find_packed(qw, qf, res2), steps=4, warm=2)
        if rank == 0:
            print(json.dumps({"variant": "pipelined", "sub_batches": nsub, "route_search_gather_ctas_per_sm": per, "lookups_per_s": nq * world / msp * 1e3,
                              "ms": msp, "equals_sequential": bool(torch.equal(ref, res2))}), flush=True)
        del pl
        torch.cuda.empty_cache()
g.dispose()
if world > 1:
    dist.destroy_process_group()
