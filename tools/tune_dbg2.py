import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth
k, c, n = 47, 4, 25_000_000
L = N.lib()
body, _ = synth.make_graph_body(1, n, k, c, device="cuda")
g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
cap = n // 8
out = torch.empty(cap * 21 + 64, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
parents = np.arange(1, 4, dtype=np.int32)
for stg in (4096, 1024):
    for chunk in (16, 8):
        for dbg in (0, 4, 8, 12, 32):
            N.set_option("scan_debug", dbg); N.set_option("scan_chunk_tiles", chunk); N.set_option("scan_stage_buf_bytes", stg)
            step = lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), None, cap, cnt.data_ptr(), st))
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("stg=%d chunk=%2d dbg=%2d (no copy-out=%d, no staging stores=%d, nothing novel=%d)  %.4f ms %.0f GB/s" % (
                stg, chunk, dbg, (dbg >> 2) & 1, (dbg >> 3) & 1, (dbg >> 5) & 1, ms, n * 36 / ms / 1e6), flush=True)
