"""Second-level sweep of the fast scan kernel: CTAs/SM x staging size x tile x stages (GPU box only)."""
import itertools, sys
import numpy as np, torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth
k, c, n = 47, 4, 25_000_000
L = N.lib()
body, _ = synth.make_graph_body(1, n, k, c, device="cuda")
g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
parents = np.arange(1, c, dtype=np.int32)
cap = n // 8
out = torch.empty(cap * 21 + 64, dtype=torch.uint8, device="cuda"); cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
res = []
for ctas, stg, tile, stages in itertools.product((2, 3, 4), (2048, 8192), (16384, 20480, 28672, 40960, 49152), (2, 3, 4, 6)):
    for kname, v in (("scan_ctas_per_sm", ctas), ("scan_stage_buf_bytes", stg), ("scan_tile_bytes", tile), ("scan_stages", stages)):
        N.set_option(kname, v)
    step = lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), None, cap, cnt.data_ptr(), st))
    try:
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res.append((ms, ctas, stg, tile, stages, int(cnt[0])))
    except cb.CortexJDKException as e:
        pass
for r in sorted(res)[:25]:
    print("%.4f ms  %.0f GB/s  ctas=%d stg=%d tile=%d stages=%d novel=%d" % (r[0], n * 36 / r[0] / 1e6, *r[1:]))
print("worst", sorted(res)[-1])
