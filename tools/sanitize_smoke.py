"""A small end-to-end pass over every kernel for compute-sanitizer (memcheck): fixture + tiny synthetic graphs."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import corticall_b200 as cb
from corticall_b200.host.sharded import RoutedLookup
from tools import synth

for k, c, n in ((47, 4, 3000), (63, 21, 700), (31, 1, 1500), (31, 45, 300)):
    ctx = synth.make_ctx_file(9, n, k, c, novel_permille=30, adv_period=97)
    g = cb.CortexGraph(ctx)
    cnt, recs, idx = g.findNovel(0, list(range(1, c)))
    cnt2, _, _ = g.findNovel(0, [])                         # dense: rewrite path
    words, cov, edges = g.decodeRecords(0, n)
    tw = [torch.from_numpy(words[:, w].copy().view(np.int64)) for w in range(words.shape[1])]
    q, canon, valid = synth.make_queries(3, tw, k, 2000, corrupt_permille=30)
    for algo in (0, 1, 2):
        r = g.findRecordIndices(q.numpy(), algo)
    seq = synth.random_genome(1, 5000, n_permille=3).numpy()
    g.findWindows(seq)
    cb.packCanonical(seq, k)
    j = cb.CortexGraph.join([g, g])
    s = j.sorted()
    print(k, c, n, "novel", cnt, cnt2, "hits", int((r >= 0).sum()), "join", j.getNumRecords(), flush=True)
    s.dispose(); j.dispose()
    if words.shape[1] == 2:
        dev = torch.device("cuda", 0)
        world, cap = 3, 2000
        blocks = [torch.zeros(RoutedLookup.block_elems(world, cap, k), dtype=torch.int64, device=dev) for _ in range(world)]
        spl = torch.stack([torch.stack([t[n * r // world] for t in tw]) for r in range(1, world)]).cuda()
        body = torch.from_numpy(g.getRawRecords(0, n)).cuda()
        rls, shards = [], []
        for rnk in range(world):
            lo, hi = n * rnk // world, n * (rnk + 1) // world
            sg = cb.CortexGraph.fromDevice(body[lo:hi].data_ptr(), k, c, hi - lo, firstIndex=lo, keepalive=body)
            shards.append(sg)
            rls.append(RoutedLookup(sg, spl, rnk, world, dev, cap, k, shard_first=[n * x // world for x in range(world)], emulate=blocks))
        qw = torch.stack(canon, dim=1).contiguous().cuda(); qf = torch.where(valid, 0, 2).to(torch.uint8).cuda()
        out = torch.empty(2000, dtype=torch.int64, device=dev)
        rls[0].route(qw, qf)
        for rl in rls[1:]:
            rl.route(qw[:0], qf[:0])
        for rl in rls:
            rl.search()
        rls[0].gather(out)
        torch.cuda.synchronize()
        assert (out.cpu().numpy() == r).all()
        for sg in shards:
            sg.dispose()
    g.dispose()
print("sanitize smoke done")
