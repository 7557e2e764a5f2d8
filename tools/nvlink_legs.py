"""NVLink traffic of the routed lookup, from the driver's own link counters (not from our arithmetic):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/nvlink_legs.py [table] [queries_total] [calls]
Every rank reads its GPU's NVLink data counters (NVML field values, and `nvidia-smi nvlink -gt d` as a second source) before
and after `calls` batches of RoutedLookup.find_packed, and prints the bytes per batch beside what the exchange has to move:
keys out in the 12-byte wire form for the queries another GPU owns (route leg, runs padded to 4 keys), 4-byte results pulled
back for the same queries (gather leg).  Per-leg rates = those bytes / the legs' device times."""
import json
import os
import re
import subprocess
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
import corticall_b200 as cb
from corticall_b200.host.sharded import RoutedLookup
from tools import synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq_total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 400_000_000
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 20
K, C_ = 47, 4
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = bench.Env(rank, world, local)
dev = env.dev


def counters():
    """(tx_bytes, rx_bytes, source) summed over this GPU's links."""
    out = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "0,1,2,3,4,5,6,7").split(",")[local]))
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF), (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        got = []
        for v in vals:
            if v.nvmlReturn != 0:
                raise RuntimeError("nvml field return %d" % v.nvmlReturn)
            got.append(int(v.value.ullVal) * 1024)          # KiB
        out["nvml"] = got
    except Exception as e:                                   # noqa: BLE001 -- second source below
        out["nvml_error"] = str(e)[:120]
    try:
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(local)], capture_output=True, text=True, timeout=30).stdout
        txs, rxs = re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt), re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt)
        if txs and rxs:
            out["smi"] = [sum(int(x) for x in txs) * 1024, sum(int(x) for x in rxs) * 1024]
        else:
            out["smi_error"] = "nvidia-smi nvlink -gt d reports no counters (N/A)"
        out["smi_head"] = txt[:300]
    except Exception as e:                                   # noqa: BLE001
        out["smi_error"] = str(e)[:120]
    return out


words = synth.random_canonical_keys(bench.SEED_LOOKUP, nt, K, dev)
lo, hi = nt * rank // world, nt * (rank + 1) // world
cov, edges = synth.coverage_and_edges(bench.SEED_LOOKUP, hi - lo, C_, dev, offset=lo)
body = synth.assemble_records([w[lo:hi] for w in words], cov, edges)
del cov, edges
splitters = torch.stack([torch.stack([w[nt * r // world] for w in words]) for r in range(1, world)])
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
# queries this rank sends away: owner by an independent torch formulation of the splitter rule
owner = synth.owner_of_keys([qw[:, 0], qw[:, 1]], splitters)
remote = int(((owner != rank) & (qf == 0)).sum())
del owner
del words
res = torch.empty(nq, dtype=torch.int64, device=dev)
rl = RoutedLookup(g, splitters, rank, world, dev, cap=int(nq / world * 1.25) + 4096, k=K, max_batch=nq)
for _ in range(3):
    rl.find_packed(qw, qf, res)
rl.find_packed(qw, qf, res, profile=True)
phase = dict(rl.phase_ms)
torch.cuda.synchronize()
env.barrier()
c0 = counters()
ms = env.timeit(lambda: rl.find_packed(qw, qf, res), steps=calls, warm=0)
torch.cuda.synchronize()
env.barrier()
c1 = counters()
line = {"rank": rank, "n_gpus": world, "queries_per_rank": nq, "calls": calls, "ms_per_call": ms, "phase_ms": {k: round(v, 3) for k, v in phase.items()},
        "remote_queries_per_call": remote}
for src in ("nvml", "smi"):
    if src in c0 and src in c1:
        tx, rx = (c1[src][0] - c0[src][0]) / calls, (c1[src][1] - c0[src][1]) / calls
        line[src + "_tx_bytes_per_call"], line[src + "_rx_bytes_per_call"] = tx, rx
for k in ("nvml_error", "smi_error"):
    if k in c1:
        line[k] = c1[k]
if "smi_head" in c1 and rank == 0:
    line["smi_head"] = c1["smi_head"]
if remote is not None:
    line["algorithmic_route_bytes_out"] = remote * 12          # wire keys
    line["algorithmic_gather_bytes_in"] = remote * 4           # results pulled back
    if phase.get("route") and phase.get("gather"):
        line["route_gb_per_s_algorithmic"] = remote * 12 / phase["route"] / 1e6
        line["gather_gb_per_s_algorithmic"] = remote * 4 / phase["gather"] / 1e6
allp = [None] * world
dist.all_gather_object(allp, line)
if rank == 0:
    for p in allp:
        print(json.dumps(p), flush=True)
g.dispose()
dist.destroy_process_group()
