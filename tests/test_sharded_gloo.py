"""world_size-2 (and 3) gloo runs of the multi-GPU lookup exchange (corticall_b200/host/sharded.py) on CPU tensors.

What is under test is the HOST logic around the device steps: owner counts -> count all-to-all -> variable
all-to-all of the queries -> local search -> reverse all-to-all -> scatter to the original slots, including ranks
that receive nothing, empty batches and flagged (never-routed) queries.  The three device steps are replaced by
numpy stand-ins defined HERE (the product always uses CudaOps; there is no CPU path in the package); expected
results come from the oracle's findRecord over the unsharded graph."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from corticall_b200.host.sharded import ShardedLookup
from oracle import oracle_np as onp
from oracle import orc
from tools import synth

K, C_, N_ = 47, 4, 6000


def u(t: torch.Tensor) -> np.ndarray:
    return t.numpy().view(np.uint64)


def key_tuple_array(words: np.ndarray) -> np.ndarray:
    """[n, s] uint64 -> structured array that sorts like the multiword unsigned key."""
    s = words.shape[1]
    out = np.zeros(words.shape[0], dtype=[("w%d" % i, ">u8") for i in range(s)])
    for i in range(s):
        out["w%d" % i] = words[:, i]
    return out


class NumpyOps:
    """CPU stand-ins for cc_bucket_by_owner_dev / cc_find_packed_dev / cc_scatter_results_dev (test only)."""

    def __init__(self, shard_words: np.ndarray, first_index: int):
        self.keys = key_tuple_array(shard_words)
        self.first = first_index

    def bucket(self, words, flags, splitters, world):
        w = u(words)
        ok = np.ones(len(w), dtype=bool) if flags is None else (flags.numpy() & 6) == 0
        if world > 1:
            owner = np.searchsorted(key_tuple_array(u(splitters)), key_tuple_array(w), side="right")
        else:
            owner = np.zeros(len(w), dtype=np.int64)
        order = np.argsort(np.where(ok, owner, world), kind="stable")
        order = order[:int(ok.sum())]
        counts = np.bincount(owner[ok], minlength=world).astype(np.int64)
        sorted_words = np.zeros_like(w)
        sorted_words[:len(order)] = w[order]
        slots = np.zeros(len(w), dtype=np.int32)
        slots[:len(order)] = order
        return torch.from_numpy(counts), torch.from_numpy(sorted_words.view(np.int64)), torch.from_numpy(slots)

    def search(self, words):
        q = key_tuple_array(u(words))
        pos = np.searchsorted(self.keys, q, side="left")
        hit = (pos < len(self.keys)) & (self.keys[np.minimum(pos, len(self.keys) - 1)] == q) if len(self.keys) else np.zeros(len(q), bool)
        return torch.from_numpy(np.where(hit, pos + self.first, -1).astype(np.int64))

    def scatter(self, values, slots, out):
        out.numpy()[slots.numpy()] = values.numpy()


def worker(rank: int, world: int, port: int, ctx: bytes, result_dir: str, skew: bool):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = onp.parse_header(ctx)
        rec = onp.records_view(ctx, h)
        table = np.ascontiguousarray(rec["kmer"]).astype(np.uint64)            # [n, s] native words, ascending
        n = len(table)
        lo, hi = n * rank // world, n * (rank + 1) // world
        splitters = torch.from_numpy(np.stack([table[n * r // world] for r in range(1, world)]).view(np.int64))
        ops = NumpyOps(table[lo:hi], lo)
        sl = ShardedLookup(None, splitters, rank, world, torch.device("cpu"), ops=ops)
        tw = [torch.from_numpy(table[:, w].copy().view(np.int64)) for w in range(table.shape[1])]
        nq = 0 if (skew and rank == 1) else 3000 + 500 * rank                   # ragged batches; one rank may own nothing
        ascii_q, canon, valid = synth.make_queries(100 + rank, tw, K, max(nq, 1), corrupt_permille=20)
        ascii_q, canon, valid = ascii_q[:nq], [c[:nq] for c in canon], valid[:nq]
        words = torch.stack(canon, dim=1).contiguous() if nq else torch.zeros((0, 2), dtype=torch.int64)
        if skew and nq:                                                          # every query goes to the last shard
            words = torch.from_numpy(np.repeat(table[-1:].view(np.int64), nq, axis=0).copy())
            valid = torch.ones(nq, dtype=torch.bool)
            ascii_q = None
        flags = torch.where(valid, 0, 2).to(torch.uint8)
        out = torch.empty(nq, dtype=torch.int64)
        sl.find_packed(words, flags, out)
        if ascii_q is not None:
            want = orc.Graph(ctx).find_batch(ascii_q.numpy()) if nq else np.zeros(0, dtype=np.int64)
        else:
            want = np.full(nq, n - 1, dtype=np.int64)
        ok = bool((out.numpy() == want).all())
        hits = int((out.numpy() >= 0).sum())
        with open(os.path.join(result_dir, "r%d" % rank), "w") as f:
            f.write("%d %d %d %s" % (ok, nq, hits, sl.last.get("sent")))
    finally:
        dist.destroy_process_group()


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,skew", [(2, False), (3, False), (2, True)])
def test_sharded_lookup_exchange(tmp_path, world, skew):
    ctx = synth.make_ctx_file(77, N_, K, C_, adv_period=0)
    mp.spawn(worker, args=(world, free_port(), ctx, str(tmp_path), skew), nprocs=world, join=True)
    total_hits = 0
    for r in range(world):
        ok, nq, hits, sent = open(tmp_path / ("r%d" % r)).read().split(" ", 3)
        assert ok == "1", (r, nq, hits, sent)
        total_hits += int(hits)
    assert total_hits > 0


def test_routed_block_layout_is_aligned_and_disjoint():
    """Host-side layout of the symmetric block of the routed lookup (no device needed): every part starts 16-byte aligned, the
    parts do not overlap, and a segment capacity is a multiple of 4 keys so that every (sub-range, source) segment of the inbox
    starts 16-byte aligned whatever the wire width (the route kernel stores aligned 16-byte vectors)."""
    from corticall_b200.host.sharded import RoutedLookup
    for world in (1, 2, 3, 8):
        for vsub in (1, 5, 8):
            if world * vsub > 64:
                continue
            for k in (16, 31, 47, 63, 65, 128):
                for cap_req in (1, 5, 4097, 19_535_346):
                    cap = RoutedLookup.round_cap(cap_req)
                    kw = (2 * k + 31) // 32
                    assert cap % 4 == 0 and cap >= cap_req
                    off_inbox, off_ret, off_counts, total = RoutedLookup._layout(world, vsub, cap, kw)
                    assert off_inbox == 0 and off_ret % 16 == 0 and off_counts % 16 == 0 and total % 8 == 0
                    assert off_ret >= vsub * world * cap * kw * 4                      # inbox fits before the results
                    assert off_counts >= off_ret + world * vsub * cap * 4             # results fit before the counts
                    assert total >= off_counts + 8 * world * vsub
                    assert RoutedLookup.block_elems(world, cap_req, k, vsub) * 8 == total
                    for v in range(vsub):                                              # sub-range blocks of the inbox
                        assert (v * world * cap * kw * 4) % 16 == 0
                        for src in range(world):                                       # and every segment inside them
                            assert ((v * world + src) * cap * kw * 4) % 16 == 0
