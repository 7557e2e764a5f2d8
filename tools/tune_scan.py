"""Sweeps the scan kernel's tuning knobs on the configs[1] shape and prints ms / GB/s per setting (GPU box only)."""
import itertools
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200 import _native as N
from tools import synth


def main():
    shapes = [(47, 4, 25_000_000)]
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        shapes += [(31, 4, 25_000_000), (63, 21, 8_000_000)]
    L = N.lib()
    for k, c, n in shapes:
        s = (k + 31) // 32
        S, O = 8 * s + 5 * c, 8 * s + 5
        body, _ = synth.make_graph_body(1, n, k, c, device="cuda")
        g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, keepalive=body)
        parents = np.arange(1, c, dtype=np.int32)
        cap = n // 8
        out = torch.empty(cap * O + 64, dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def step():
            N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, len(parents), out.data_ptr(), None, cap, cnt.data_ptr(), st))

        print("shape k=%d c=%d S=%d n=%d" % (k, c, S, n), flush=True)
        for fast, chunk, tile, stages, ctas in itertools.product((1, 0), (1, 2, 4, 8, 16), (16384, 24576, 32768, 49152), (2, 3, 4), (1, 2)):
            if not fast and chunk != 1:
                continue
            N.set_option("scan_fast", fast); N.set_option("scan_chunk_tiles", chunk); N.set_option("scan_tile_bytes", tile)
            N.set_option("scan_stages", stages); N.set_option("scan_ctas_per_sm", ctas)
            try:
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 20
                print("fast=%d chunk=%2d tile=%6d stages=%d ctas=%d  %.3f ms  %.0f GB/s  novel=%d" % (
                    fast, chunk, tile, stages, ctas, ms, n * S / ms / 1e6, int(cnt[0])), flush=True)
            except cb.CortexJDKException as e:
                print("fast=%d chunk=%2d tile=%6d stages=%d ctas=%d  unsupported: %s" % (fast, chunk, tile, stages, ctas, e), flush=True)
        g.dispose()


if __name__ == "__main__":
    main()
