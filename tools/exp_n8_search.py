"""Per-rank device times of the three routed-lookup legs (multi-GPU, under torchrun), configs[2] shape.
History: with the results stored straight into the origins' buffers by the search kernel, the search leg took 2.9 ms on rank 0 but
3.6-4.3 ms on the other ranks at 8 GPUs -- only in the first search after a route, independent of L2 flushes, host syncs or
sleeps, gone (2.96 ms on every rank) when the stores were redirected to local memory (profiles/r1_exp_n8_search.log).  Hence the
results now stay on the owner and the gather leg pulls them.
usage: torchrun --nproc-per-node 8 tools/exp_n8_search.py"""
import os, subprocess, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from corticall_b200.host.sharded import RoutedLookup
from tools import synth

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
K, C, nt, nq = 47, 4, 100_000_000, 1_000_000_000 // world
words = synth.random_canonical_keys(20261019, nt, K, dev)
lo, hi = nt * rank // world, nt * (rank + 1) // world
cov, edges = synth.coverage_and_edges(20261019, hi - lo, C, dev, offset=lo)
body = synth.assemble_records([w[lo:hi] for w in words], cov, edges)
splitters = torch.stack([torch.stack([w[nt * r // world] for w in words]) for r in range(1, world)])
g = cb.CortexGraph.fromDevice(body.data_ptr(), K, C, hi - lo, firstIndex=lo, device=lr, keepalive=body)
g.buildIndex()
qw = torch.empty((nq, 2), dtype=torch.int64, device=dev); qf = torch.empty(nq, dtype=torch.uint8, device=dev)
for o in range(0, nq, 1 << 24):
    m = min(1 << 24, nq - o)
    _, canon, valid = synth.make_queries(20261019, words, K, m, offset=rank * nq + o)
    qw[o:o + m, 0], qw[o:o + m, 1] = canon[0], canon[1]
    qf[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
del words
rl = RoutedLookup(g, splitters, rank, world, dev, cap=int(nq / world * 1.25) + 4096, k=K, max_batch=nq)
rl.route(qw, qf); rl._barrier()
torch.cuda.synchronize(); dist.barrier()


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", str(lr)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return out
    except Exception as e:
        return str(e)


def time_search(reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rl.search(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        rl.search()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gather_all(x):
    out = [None] * world
    dist.all_gather_object(out, x)
    return out


out = torch.empty(nq, dtype=torch.int64, device=dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def pipeline():
    """one batch; returns (route, search, gather) device times of this rank"""
    torch.cuda.synchronize(); dist.barrier()
    e0 = ev(); rl.route(qw, qf); e1 = ev(); rl._barrier()
    e2 = ev(); rl.search(); e3 = ev(); rl._barrier()
    e4 = ev(); rl.gather(out); e5 = ev()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1), 3), round(e2.elapsed_time(e3), 3), round(e4.elapsed_time(e5), 3)


for hints in (3, 3, 0, 3):
    N.set_option("lookup_l2_hints", hints)
    r = gather_all(pipeline())
    if rank == 0:
        print("hints %d  route  %s\n         search %s\n         gather %s" % (hints, [x[0] for x in r], [x[1] for x in r], [x[2] for x in r]), flush=True)
dist.barrier()
dist.destroy_process_group()
