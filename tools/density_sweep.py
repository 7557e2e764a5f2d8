"""Novelty-density sweep of the scan on configs[1]-sized graphs (GPU box only): sparse path vs rewrite of dense chunks."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth
k, c, n = 47, 4, 25_000_000
L = N.lib()
st = torch.cuda.current_stream().cuda_stream
for permille in (0, 1, 5, 10, 20, 30, 50, 100, 300):
    body, _ = synth.make_graph_body(1, n, k, c, device="cuda", novel_permille=permille)
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
    cap = n
    out = torch.empty(cap * 21 + 64, dtype=torch.uint8, device="cuda"); cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    parents = np.arange(1, 4, dtype=np.int32)
    step = lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), None, cap, cnt.data_ptr(), st))
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nov = int(cnt[0])
    print("novel %5.1f %% (%8d records)  %.4f ms  %.0f GB/s algorithmic (36 B in + 21 B per novel out)" % (100.0 * nov / n, nov, ms, (n * 36 + nov * 21) / ms / 1e6), flush=True)
    g.dispose(); del body, out
