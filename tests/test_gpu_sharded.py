"""The single-process multi-GPU entry points (cc_open_sharded ...): one graph cut into k-mer-range shards, lookups routed
over peer memory, one globally ordered novelty output.  Every answer is compared bit for bit with the oracle and with the
single-GPU entry points on the same file.  With one GPU the shards share device 0 (the library allows a device to be
listed several times); the tests marked `multi` need two or more devices and skip otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import corticall_b200 as cb                      # noqa: E402
from corticall_b200 import _native as N          # noqa: E402
from oracle import orc                           # noqa: E402  (the checker)
from tools import synth                          # noqa: E402


def device_lists():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    lists = [[0], [0, 0], [0, 0, 0], [0] * 8]
    if n >= 2:
        lists += [list(range(n)), [0, 1, 0, 1]]
    return lists


@pytest.mark.parametrize("k,c,n", [(47, 4, 60000), (31, 4, 30011), (63, 3, 9000), (95, 2, 5000), (21, 2, 7)])
def test_sharded_matches_oracle_and_single_gpu(tmp_path, k, c, n):
    ctx = synth.make_ctx_file(77 + k, n, k, c, novel_permille=20, adv_period=89)
    path = tmp_path / "g.ctx"
    path.write_bytes(ctx)
    og = orc.Graph(ctx)
    single = cb.CortexGraph(ctx)
    s = single.getKmerBits()
    words, _, _ = single.decodeRecords(0, n)
    tw = [torch.from_numpy(words[:, w].copy().view(np.int64)) for w in range(s)]
    q_ascii, canon, valid = synth.make_queries(5, tw, k, 20000, corrupt_permille=30)
    qa = q_ascii.numpy()
    if n < 3:
        for i in range(n):
            og.get_record(i)
    want = og.find_batch(qa)
    pw = np.stack([cw.numpy().view(np.uint64) for cw in canon], axis=1)
    flags = np.where(valid.numpy(), 0, 2).astype(np.uint8)
    seq = synth.random_genome(3, 5000, device="cpu").numpy()
    want_windows = single.findWindows(seq)
    want_cnt, want_recs, want_idx = single.findNovel(0, list(range(1, c)))
    o_recs, o_idx = og.find_rois(0, list(range(1, c)))
    assert want_recs.tobytes() == o_recs and want_idx.tolist() == o_idx.tolist()
    single.writeRois(0, list(range(1, c)), tmp_path / "single.ctx")
    for devs in device_lists():
        for source, placement in ((path, "range"), (ctx, "range"), (path, "replicate"), (ctx, "auto")):
            sh = cb.ShardedCortexGraph(source, devs, placement)
            # a replica per device when asked for (or when the graph fits: "auto" on these sizes), unless there is one device
            assert sh.placement == ("range" if placement == "range" or len(devs) == 1 else "replicate")
            assert sh.getNumRecords() == n and sh.numShards == len(devs) and sh.getKmerSize() == k
            assert sh.findRecordIndices(qa).tolist() == want.tolist(), devs
            assert sh.findPacked(pw, flags).tolist() == want.tolist(), devs
            assert sh.findPacked(pw[valid.numpy()]).tolist() == want[valid.numpy()].tolist(), devs
            assert sh.findWindows(seq).tolist() == want_windows.tolist()
            cnt, recs, idx = sh.findNovel(0, list(range(1, c)))
            assert cnt == want_cnt and recs.tobytes() == o_recs and idx.tolist() == o_idx.tolist(), devs
            cnt, recs, _ = sh.findNovel(0, list(range(1, c)), cap=7, want_index=False)      # capped: total count, first 7 records
            assert cnt == want_cnt and recs.tobytes() == o_recs[:min(7, want_cnt) * (8 * s + 5)]
            assert sh.writeRois(0, list(range(1, c)), tmp_path / "sharded.ctx") == want_cnt
            assert (tmp_path / "sharded.ctx").read_bytes() == (tmp_path / "single.ctx").read_bytes()
            # shards are ordinary graphs over their slices (a replica is the whole graph)
            g1, dev, first = sh.shard(len(devs) - 1)
            if sh.placement == "range":
                assert first == n * (len(devs) - 1) // len(devs) and g1.getNumRecords() == n - first
            else:
                assert first == 0 and g1.getNumRecords() == n
                if n > 10:
                    assert g1.getRecord(n - 3) == single.getRecord(n - 3)
            assert g1.getColor(0).getSampleName() == single.getColor(0).getSampleName()
            sh.dispose()
    single.dispose()


def test_sharded_device_resident_batches_and_skew():
    """Queries already on the devices (cc_find_packed_sharded_dev), uneven batch sizes per device, and a batch so skewed
    that a segment overflows: the chunk is repeated in pieces and the answers stay exact."""
    k, c, n = 47, 4, 200000
    ctx = synth.make_ctx_file(9, n, k, c, adv_period=0)
    og = orc.Graph(ctx)
    ndev = torch.cuda.device_count()
    devs = [0, 0, 0, 0] if ndev < 2 else [i % ndev for i in range(4)]
    sh = cb.ShardedCortexGraph(ctx, devs)
    single = cb.CortexGraph(ctx)
    words, _, _ = single.decodeRecords(0, n)
    tw = [torch.from_numpy(words[:, w].copy().view(np.int64)) for w in range(2)]
    sizes = [30000, 0, 12345, 50001]
    qs, outs, wants = [], [], []
    for r, m in enumerate(sizes):
        a, canon, valid = synth.make_queries(40 + r, tw, k, max(m, 1), corrupt_permille=10)
        a, canon, valid = a[:m], [x[:m] for x in canon], valid[:m]
        dev = torch.device("cuda", devs[r])
        qs.append((torch.stack(canon, dim=1).contiguous().to(dev), torch.where(valid, 0, 2).to(torch.uint8).to(dev)))
        outs.append(torch.full((m,), -9, dtype=torch.int64, device=dev))
        wants.append(og.find_batch(a.numpy()) if m else np.empty(0, dtype=np.int64))
    sh.findPackedDevice([q[0] for q in qs], [q[1] for q in qs], outs)
    for r in range(4):
        assert outs[r].cpu().numpy().tolist() == wants[r].tolist(), r
    assert sh.lastStats().overflow_retries == 0
    # all hits from the lowest quarter of the key space: everything routes to shard 0 and overflows its segments
    low = torch.stack([t[:n // 4] for t in tw], dim=1)
    pick = torch.randint(0, n // 4, (240000,), generator=torch.Generator().manual_seed(1))
    qw = low[pick].contiguous()
    got = sh.findPacked(qw.numpy().view(np.uint64))
    assert got.tolist() == pick.tolist()
    assert sh.lastStats().overflow_retries >= 1
    # and the exchange keeps working afterwards
    sh.findPackedDevice([q[0] for q in qs], [q[1] for q in qs], outs)
    for r in range(4):
        assert outs[r].cpu().numpy().tolist() == wants[r].tolist(), r
    sh.dispose()
    # the same batches against replicas: no exchange, no overflow to handle, same answers
    rep = cb.ShardedCortexGraph(ctx, devs, "replicate")
    for o in outs:
        o.fill_(-9)
    rep.findPackedDevice([q[0] for q in qs], [q[1] for q in qs], outs)
    for r in range(4):
        assert outs[r].cpu().numpy().tolist() == wants[r].tolist(), r
    assert rep.findPacked(qw.numpy().view(np.uint64)).tolist() == pick.tolist() and rep.lastStats().overflow_retries == 0
    rep.dispose(); single.dispose()


def test_sharded_rejects_bad_input(tmp_path):
    ctx = synth.make_ctx_file(3, 1000, 31, 2)
    with pytest.raises(cb.CortexJDKException) as e:
        cb.ShardedCortexGraph(ctx, [0, 99])
    assert e.value.status == N.CC_ERR_ARG
    with pytest.raises(cb.CortexJDKException) as e:
        cb.ShardedCortexGraph(tmp_path / "missing.ctx", [0])
    assert e.value.status == N.CC_ERR_IO
    with pytest.raises(cb.CortexJDKException) as e:
        cb.ShardedCortexGraph(b"NOTCORTEX" + ctx[9:], [0, 0])
    assert e.value.status == N.CC_ERR_NOT_CORTEX
    # records out of order across a shard boundary only: every shard is sorted on its own, the whole is not
    g = cb.CortexGraph(ctx)
    raw = g.getRawRecords(0, 1000).copy()
    hdr = ctx[:len(ctx) - raw.size]
    swapped = np.concatenate([raw[500:], raw[:500]])
    sh = cb.ShardedCortexGraph(hdr + swapped.tobytes(), [0, 0])
    with pytest.raises(cb.CortexJDKException) as e:
        sh.findPacked(np.zeros((4, 1), dtype=np.uint64))
    assert e.value.status == N.CC_ERR_UNSORTED
    # the scan does not need sorted input
    cnt, _, _ = sh.findNovel(0, [1])
    assert cnt == g.findNovel(0, [1])[0]
    sh.dispose(); g.dispose()
    # the routed kernels keep their owner tables in shared memory: at most 64 owners (world x virtual shards)
    import torch as _t
    buf = _t.zeros(4096, dtype=_t.int64, device="cuda")
    ptrs = (N._P * 65)(*([buf.data_ptr()] * 65))
    rc = N.lib().cc_route_queries_dev(0, buf.data_ptr(), buf.data_ptr(), 16, 31, buf.data_ptr(), 65, 0, 64, ptrs, ptrs, buf.data_ptr(), 16,
                                      buf.data_ptr(), None)
    assert rc == N.CC_ERR_ARG and "1..64" in N.last_error()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two or more GPUs")
def test_sharded_real_devices_large():
    """The real multi-GPU path at a size where every leg runs many tiles: 4e6 records, 2e6 queries, oracle on a sample."""
    k, c, n = 47, 4, 4_000_000
    ndev = torch.cuda.device_count()
    body, table = synth.make_graph_body(11, n, k, c, device="cpu")
    ctx = synth.header_bytes(k, c) + body.numpy().tobytes()
    og = orc.Graph(ctx)
    sh = cb.ShardedCortexGraph(ctx, list(range(ndev)))
    a, canon, valid = synth.make_queries(12, table, k, 2_000_000)
    got = sh.findRecordIndices(a.numpy())
    sample = np.random.default_rng(0).choice(a.shape[0], 200_000, replace=False)
    assert got[sample].tolist() == og.find_batch(a.numpy()[sample]).tolist()
    st = sh.lastStats()
    assert st.launches >= 3 * ndev and st.route_ms > 0 and st.search_ms > 0 and st.gather_ms > 0
    cnt, recs, idx = sh.findNovel(0, [1, 2, 3])
    want, widx = og.find_rois(0, [1, 2, 3])
    assert cnt == len(widx) and recs.tobytes() == want and idx.tolist() == widx.tolist()
    sh.dispose()
