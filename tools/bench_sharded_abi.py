"""The single-process multi-GPU entry points at benchmark size (GPU box with >= 2 GPUs; not a bench line):
  python tools/bench_sharded_abi.py [table_records] [queries_total]
One process, one host thread: the configs[2] table is cut into k-mer-range shards over all visible GPUs (cc_open_sharded_device),
every GPU holds its slice of the batch, and cc_find_packed_sharded_dev routes / searches / gathers with CUDA events as the
cross-device barriers.  Device 0's answers are checked against a whole-table graph on device 0."""
import json
import sys
import time
import torch
sys.path.insert(0, ".")
import bench
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq_total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
K, C_ = 47, 4
world = torch.cuda.device_count()
nq = nq_total // world
bodies, qws, qfs, outs = [], [], [], []
for r in range(world):
    dev = torch.device("cuda", r)
    torch.cuda.set_device(r)
    words = synth.random_canonical_keys(bench.SEED_LOOKUP, nt, K, dev)
    lo, hi = nt * r // world, nt * (r + 1) // world
    cov, edges = synth.coverage_and_edges(bench.SEED_LOOKUP, hi - lo, C_, dev, offset=lo)
    bodies.append(synth.assemble_records([w[lo:hi] for w in words], cov, edges))
    del cov, edges
    qw = torch.empty((nq, 2), dtype=torch.int64, device=dev)
    qf = torch.empty(nq, dtype=torch.uint8, device=dev)
    for o in range(0, nq, 1 << 24):
        m = min(1 << 24, nq - o)
        _, canon, valid = synth.make_queries(bench.SEED_LOOKUP, words, K, m, offset=r * nq + o)
        qw[o:o + m, 0], qw[o:o + m, 1] = canon[0], canon[1]
        qf[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
    qws.append(qw); qfs.append(qf); outs.append(torch.empty(nq, dtype=torch.int64, device=dev))
    if r == 0:
        cov, edges = synth.coverage_and_edges(bench.SEED_LOOKUP, nt, C_, dev)
        whole_body = synth.assemble_records(words, cov, edges)
        del cov, edges
    del words
    torch.cuda.synchronize()
sh = cb.ShardedCortexGraph.fromDevice([b.data_ptr() for b in bodies], [b.shape[0] for b in bodies], K, C_, list(range(world)), keepalive=bodies)
for _ in range(2):
    sh.findPackedDevice(qws, qfs, outs)
times = []
for _ in range(5):
    t0 = time.perf_counter()
    sh.findPackedDevice(qws, qfs, outs)        # synchronous: returns when every device's results are in place
    times.append(time.perf_counter() - t0)
st = sh.lastStats()
torch.cuda.set_device(0)
whole = cb.CortexGraph.fromDevice(whole_body.data_ptr(), K, C_, nt, device=0, keepalive=whole_body)
ref = torch.empty(nq, dtype=torch.int64, device="cuda:0")
N.check(N.lib().cc_find_packed_dev(whole._h, qws[0].data_ptr(), qfs[0].data_ptr(), nq, ref.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
ms = sorted(times)[len(times) // 2] * 1e3
print(json.dumps({"api": "cc_open_sharded_device + cc_find_packed_sharded_dev (one process, one host thread)", "n_gpus": world, "table_records": nt,
                  "queries_per_call": nq * world, "lookups_per_s": nq * world / ms * 1e3, "ms_per_call_wall": ms,
                  "device_ms_first_chunk": {"route": st.route_ms, "search": st.search_ms, "gather": st.gather_ms, "route_to_gather_end": st.chunk_ms},
                  "launches": st.launches, "overflow_retries": st.overflow_retries, "device0_equals_whole_table_graph": bool(torch.equal(outs[0], ref))}))

# ---- the same batch against REPLICAS (cc_open_sharded_memory_placed, CC_PLACE_REPLICATE / AUTO): a full copy and index per device,
# every device answers its own queries -- what the library picks by itself ("auto") for a graph of this size
sh.dispose()
del bodies
import numpy as np
hdr = synth.header_bytes(K, C_)
image = np.empty(len(hdr) + whole_body.numel(), dtype=np.uint8)
image[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
torch.from_numpy(image[len(hdr):]).copy_(whole_body.reshape(-1))
t0 = time.perf_counter()
rep = cb.ShardedCortexGraph(memoryview(image), list(range(world)), "auto")
open_s = time.perf_counter() - t0
for o in outs:
    o.fill_(-7)
t0 = time.perf_counter()
rep.findPackedDevice(qws, qfs, outs)           # first call builds the per-device indices
first_s = time.perf_counter() - t0
times = []
for _ in range(5):
    t0 = time.perf_counter()
    rep.findPackedDevice(qws, qfs, outs)
    times.append(time.perf_counter() - t0)
ms = sorted(times)[len(times) // 2] * 1e3
same = bool(torch.equal(outs[0], ref))
last = outs[world - 1].to("cuda:0")
# the last device's answers against the whole-table graph too (its queries copied to device 0)
torch.cuda.set_device(0)
q_last, f_last = qws[world - 1].to("cuda:0"), qfs[world - 1].to("cuda:0")
torch.cuda.synchronize()
N.check(N.lib().cc_find_packed_dev(whole._h, q_last.data_ptr(), f_last.data_ptr(), nq, ref.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print(json.dumps({"api": "cc_open_sharded_memory_placed(CC_PLACE_AUTO) + cc_find_packed_sharded_dev (one process, one host thread)", "placement": rep.placement,
                  "n_gpus": world, "table_records": nt, "queries_per_call": nq * world, "lookups_per_s": nq * world / ms * 1e3, "ms_per_call_wall": ms,
                  "open_seconds_upload_to_all_devices": open_s, "first_call_seconds_with_index_build": first_s,
                  "device0_equals_whole_table_graph": same, "last_device_equals_whole_table_graph": bool(torch.equal(last, ref))}))
