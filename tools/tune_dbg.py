"""Diagnosis sweep over scan_debug bits of the fast kernel (GPU box only)."""
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
N.set_option("scan_chunk_tiles", 16)
for tile, stages, ctas in ((32768, 3, 2), (32768, 2, 2), (24576, 4, 2)):
    for npar in (3, 1, 0):
        parents = np.arange(1, 1 + npar, dtype=np.int32)
        for dbg in (0, 32, 2):
            N.set_option("scan_debug", dbg); N.set_option("scan_tile_bytes", tile)
            N.set_option("scan_stages", stages); N.set_option("scan_ctas_per_sm", ctas)
            step = lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, npar, out.data_ptr(), None, cap, cnt.data_ptr(), st))
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("tile=%d stages=%d ctas=%d parents=%d dbg=%2d  %.3f ms %.0f GB/s novel=%d" % (
                tile, stages, ctas, npar, dbg, ms, n * 36 / ms / 1e6, int(cnt[0])), flush=True)
