"""Runs the configs[1] novelty scan a few times (for ncu / quick timing).  GPU box only."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth

k, c, n = 47, 4, 25_000_000
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for kv in sys.argv[2:]:
    name, val = kv.split("=")
    N.set_option(name, int(val))
L = N.lib()
body, _ = synth.make_graph_body(20261018, n, k, c, device="cuda")
g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
parents = np.arange(1, c, dtype=np.int32)
cap = n // 8
out = torch.empty(cap * 21 + 64, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
e[0].record()
for i in range(reps):
    N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), None, cap, cnt.data_ptr(), st))
    e[i + 1].record()
torch.cuda.synchronize()
print("novel", int(cnt[0]), "ms", ["%.3f" % e[i].elapsed_time(e[i + 1]) for i in range(reps)])
