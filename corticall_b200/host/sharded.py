"""Multi-GPU lookups: the sorted record array is sharded by k-mer range (contiguous record slices, one per
rank); each rank owns a slice of the query batch, routes every canonical packed query to the rank whose
k-mer range contains it, the owner searches its shard, and the record indices travel back to the query's
original slot.  One process per GPU; torch.distributed is the plumbing.

The reference has no counterpart (it is a single JVM over one mmap'ed file); what is reproduced is the RESULT of
`CortexGraph.findRecord` (CortexGraph.java:272-317) for every query: the global record index, or -1.

Exchange per batch:
  1. owner = number of splitters <= query   (splitter r = first key of shard r; cc_bucket_by_owner_dev: count,
     exclusive scan, stable scatter into per-owner runs + original slots)
  2. counts all-to-all (world x int64), then variable all-to-all of the query words
  3. local search on the owner (cc_find_packed_dev; indices already rebased by the shard's first record index)
  4. reverse all-to-all of the int64 indices, scatter to the original slots (cc_scatter_results_dev)

The three device steps are injectable (`ops`) so the exchange logic can be exercised by world_size-2 gloo tests
on CPU tensors with numpy stand-ins supplied BY THE TEST; the product always uses `CudaOps` -- there is no
fallback selection here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .. import _native as N


class CudaOps:
    """The device steps, through the C ABI."""

    def __init__(self, graph, device):
        self.g = graph
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()

    def bucket(self, words, flags, splitters, world):
        nq, s = words.shape
        counts = torch.zeros(world, dtype=torch.int64, device=words.device)
        sorted_words = torch.empty_like(words)
        slots = torch.empty(nq, dtype=torch.int32, device=words.device)
        st = torch.cuda.current_stream().cuda_stream
        N.check(N.lib().cc_bucket_by_owner_dev(self.index, words.data_ptr(), flags.data_ptr() if flags is not None else None, nq, s,
                                               splitters.data_ptr() if splitters is not None else None, world,
                                               counts.data_ptr(), sorted_words.data_ptr(), slots.data_ptr(), st))
        return counts, sorted_words, slots

    def search(self, words):
        out = torch.empty(words.shape[0], dtype=torch.int64, device=words.device)
        st = torch.cuda.current_stream().cuda_stream
        if words.shape[0]:
            N.check(N.lib().cc_find_packed_dev(self.g._h, words.data_ptr(), None, words.shape[0], out.data_ptr(), N.CC_ALGO_AUTO, st))
        return out

    def scatter(self, values, slots, out):
        st = torch.cuda.current_stream().cuda_stream
        if values.shape[0]:
            N.check(N.lib().cc_scatter_results_dev(self.index, values.data_ptr(), slots.data_ptr(), values.shape[0], out.data_ptr(), st))


class ShardedLookup:
    def __init__(self, graph, splitters, rank: int, world: int, device, ops=None, group=None):
        """graph: this rank's shard (CortexGraph.fromDevice(..., firstIndex=shard offset));
        splitters: int64 [world-1, s] = first key of shards 1..world-1 (identical on every rank)."""
        self.rank, self.world, self.group = rank, world, group
        self.splitters = splitters.contiguous() if splitters is not None else None
        self.ops = ops if ops is not None else CudaOps(graph, device)
        self.last = {}
        self.profile = False          # True: record per-phase device times (CUDA events) into self.last["phase_ms"]

    def find_packed(self, words: torch.Tensor, flags: torch.Tensor | None, out: torch.Tensor) -> torch.Tensor:
        """words int64 [nq, s] canonical packed queries (this rank's part of the batch), flags uint8 [nq] or None
        (bit1/bit2 set = cannot match), out int64 [nq] receives the GLOBAL record index or -1."""
        world = self.world
        marks = []

        def mark(name):
            if self.profile and words.is_cuda:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        out.fill_(-1)                                     # flagged queries are never routed
        counts, sorted_words, slots = self.ops.bucket(words, flags, self.splitters, world)
        mark("bucket")
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts, group=self.group)
        send = counts.tolist()                            # host needs the split sizes
        recv = recv_counts.tolist()
        mark("counts")
        n_send, n_recv = sum(send), sum(recv)
        s = words.shape[1]
        inbox = torch.empty((n_recv, s), dtype=words.dtype, device=words.device)
        dist.all_to_all_single(inbox, sorted_words[:n_send], output_split_sizes=recv, input_split_sizes=send, group=self.group)
        mark("route")
        found = self.ops.search(inbox)
        mark("search")
        back = torch.empty(n_send, dtype=torch.int64, device=words.device)
        dist.all_to_all_single(back, found, output_split_sizes=send, input_split_sizes=recv, group=self.group)
        mark("return")
        self.ops.scatter(back, slots[:n_send], out)
        mark("scatter")
        self.last = {"sent": send, "received": recv}
        if marks:
            torch.cuda.synchronize()
            self.last["phase_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        return out


class RoutedLookup:
    """The same sharded lookup with NO collective library on the data path: every leg is one kernel fused with its
    transfer over peer-mapped memory (NVLink P2P stores):

        cc_route_queries_dev  owner search + block-aggregated reservation + store of each key (compact wire format:
                              ceil(2k/32) 32-bit words) straight into the owner's inbox segment for this rank; only the
                              tile bookkeeping stays local (2 bytes per query + one run descriptor per tile and owner)
        barrier               (symmetric-memory signal pads; orders the peer stores)
        cc_find_routed_dev    the owner searches every inbox segment; each result (4-byte shard-local index) stays on the
                              owner, same segment and position
        barrier
        cc_gather_routed_dev  the origin pulls its tiles' runs from the owners (P2P reads), rebases by the owner's first
                              record index and writes out[] in query order -- all coalesced

    The host never reads a count, so the whole batch is asynchronous on the current stream.  Buffers are one symmetric
    allocation per rank (torch.distributed._symmetric_memory: plumbing only).  `cap` is the capacity of one
    (source, owner) segment; a rank can route at most `cap` queries to one owner per batch (worst case = its batch size).
    """

    def __init__(self, graph, splitters, rank: int, world: int, device, cap: int, k: int, shard_first=None, group=None, emulate=None,
                 max_batch: int | None = None, vsub: int = 1):
        """cap: keys one (source, owner) segment can hold.  max_batch: largest batch of this rank (default cap).  With
        max_batch <= cap (the default) a segment cannot overflow; with max_batch > cap (balanced key ranges, memory-saving)
        find_packed() checks the sent counts after the batch and raises if a segment overflowed.
        vsub: virtual shards per rank (see corticall_cuda.h): splitters then holds world * vsub - 1 keys (first key of every
        sub-range but the first; `virtual_splitters` builds them) and the search walks L2-sized sub-ranges."""
        import ctypes as C
        self.g, self.rank, self.world, self.cap, self.k = graph, rank, world, self.round_cap(cap), int(k)
        self.vsub = int(vsub)
        self.nv = world * self.vsub
        self.max_batch = int(max_batch) if max_batch is not None else int(cap)
        self.kw = (2 * self.k + 31) // 32
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.splitters = splitters.contiguous() if splitters is not None else None
        assert (self.splitters.shape[0] if self.splitters is not None else 0) == self.nv - 1, "need world * vsub - 1 splitters"
        if shard_first is None:                 # first global record index of every rank's shard
            mine = torch.tensor([int(graph.firstIndex)], dtype=torch.int64, device=device)
            if world > 1:
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine, group=group)
                shard_first = torch.cat(parts)
            else:
                shard_first = mine
        shard_first = torch.as_tensor(shard_first, dtype=torch.int64).to(device)
        assert shard_first.numel() == world
        self.shard_first = shard_first.repeat_interleave(self.vsub).contiguous()      # per virtual owner
        # layout of the symmetric block, in bytes (every part 16-byte aligned)
        self.off_inbox, self.off_ret, self.off_counts, total = self._layout(world, self.vsub, self.cap, self.kw)
        if emulate is None:
            import torch.distributed._symmetric_memory as symm
            self.block = symm.empty(total // 8, dtype=torch.int64, device=device)
            self.hdl = symm.rendezvous(self.block, (group or dist.group.WORLD).group_name)
            bases = [int(p) for p in self.hdl.buffer_ptrs]
        else:                                   # single-process emulation of all ranks on one device (tests)
            self.block = emulate[rank]
            self.hdl = None
            bases = [int(t.data_ptr()) for t in emulate]
        self.block.zero_()
        seg_inbox = world * self.cap * self.kw * 4          # one sub-range's [world][cap][kw] block
        self.p_inbox = (C.c_void_p * self.nv)(*[bases[v // self.vsub] + self.off_inbox + (v % self.vsub) * seg_inbox for v in range(self.nv)])
        self.p_counts = (C.c_void_p * self.nv)(*[bases[v // self.vsub] + self.off_counts + (v % self.vsub) * world * 8 for v in range(self.nv)])
        # results stay on the owner in the inbox's layout; entry v = this rank's result segment on virtual owner v
        self.p_res = (C.c_void_p * self.nv)(*[bases[v // self.vsub] + self.off_ret + ((v % self.vsub) * world + rank) * self.cap * 4
                                              for v in range(self.nv)])
        nbytes = C.c_uint64(0)
        N.check(N.lib().cc_route_state_bytes(self.max_batch, self.nv, C.byref(nbytes)))
        self.state = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
        self.sent = torch.zeros(max(self.nv, 8), dtype=torch.int64, device=device)
        self.nq = 0

    @staticmethod
    def round_cap(cap: int) -> int:
        """Segment capacities are multiples of 4 keys so that every segment starts 16-byte aligned."""
        return (int(cap) + 3) & ~3

    @staticmethod
    def _layout(world, vsub, cap, kw):
        al = lambda x: (x + 15) & ~15
        off_inbox = 0
        off_ret = al(vsub * world * cap * kw * 4)
        off_counts = off_ret + al(world * vsub * cap * 4)
        total = off_counts + 8 * max(world * vsub, 8)
        return off_inbox, off_ret, off_counts, (total + 7) & ~7

    @staticmethod
    def virtual_splitters(graph, rank: int, world: int, vsub: int, device, group=None, emulate_graphs=None):
        """[world * vsub - 1, s] int64: the first key of every sub-range except the very first.  Sub-range j of a rank's
        shard starts at local record n * j // vsub.  An empty shard contributes the largest key (nothing routes to it)."""
        def local(g):
            n, s = g.getNumRecords(), g.getKmerBits()
            rows = []
            for j in range(vsub):
                if n == 0:
                    rows.append(np.full(s, np.iinfo(np.int64).max, dtype=np.int64))
                else:
                    w, _, _ = g.decodeRecords(n * j // vsub, 1)
                    rows.append(w[0].view(np.int64).copy())
            return torch.from_numpy(np.stack(rows)).to(device)
        if emulate_graphs is not None:
            allk = torch.cat([local(g) for g in emulate_graphs])
        elif world > 1:
            mine = local(graph)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            allk = torch.cat(parts)
        else:
            allk = local(graph)
        return allk[1:].contiguous() if allk.shape[0] > 1 else None

    @staticmethod
    def block_elems(world: int, cap: int, k: int, vsub: int = 1) -> int:
        """int64 elements of one rank's symmetric block."""
        return RoutedLookup._layout(world, vsub, RoutedLookup.round_cap(cap), (2 * k + 31) // 32)[3] // 8

    def _barrier(self, channel: int = 0):
        """Cross-rank barrier on the current stream (symmetric-memory signal pads).  Barriers that can be in flight at the
        same time on different streams must use different channels."""
        if self.hdl is not None:
            self.hdl.barrier(channel=channel)

    def route(self, words, flags):
        if words.shape[0] > self.max_batch:
            raise ValueError("batch of %d queries exceeds the largest batch %d this lookup was sized for" % (words.shape[0], self.max_batch))
        st = torch.cuda.current_stream().cuda_stream
        self.nq = words.shape[0]
        N.check(N.lib().cc_route_queries_dev(self.index, words.data_ptr(), flags.data_ptr() if flags is not None else None,
                                             self.nq, self.k, self.splitters.data_ptr() if self.splitters is not None else None,
                                             self.nv, self.rank, self.cap, self.p_inbox, self.p_counts,
                                             self.state.data_ptr(), self.max_batch, self.sent.data_ptr(), st))

    def search(self):
        st = torch.cuda.current_stream().cuda_stream
        base = self.block.data_ptr()
        N.check(N.lib().cc_find_routed_dev(self.g._h, base + self.off_inbox, base + self.off_counts, self.world, self.vsub,
                                           self.cap, base + self.off_ret, st))

    def gather(self, out):
        st = torch.cuda.current_stream().cuda_stream
        N.check(N.lib().cc_gather_routed_dev(self.index, self.p_res, self.state.data_ptr(), self.max_batch, self.nq,
                                             self.shard_first.data_ptr(), self.nv, self.cap, out.data_ptr(), st))

    def find_packed(self, words: torch.Tensor, flags: torch.Tensor | None, out: torch.Tensor, profile: bool = False) -> torch.Tensor:
        marks = []

        def mark(name):
            if profile:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))

        mark("start")
        self.route(words, flags)
        mark("route")
        self._barrier()
        mark("barrier1")
        self.search()
        mark("search")
        self._barrier()
        mark("barrier2")
        self.gather(out)
        mark("gather")
        if marks:
            torch.cuda.synchronize()
            self.phase_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        if self.max_batch > self.cap:
            self.check_overflow()
        return out

    def check_overflow(self):
        """Only needed when max_batch > cap: a segment that received more than cap keys dropped the rest."""
        worst = int(self.sent[:self.nv].max().item())
        if worst > self.cap:
            raise OverflowError("a routed segment overflowed: %d keys for one owner, capacity %d" % (worst, self.cap))


class PipelinedRoutedLookup:
    """RoutedLookup over sub-batches with the three legs of DIFFERENT sub-batches running at the same time: while the owners
    search sub-batch i (bound by random DRAM access), sub-batch i+1 is routed (SM issue + NVLink stores) and the results of
    sub-batch i-1 are pulled back (NVLink reads) -- three resources, three streams.

    Two RoutedLookup buffer sets (parity = sub-batch & 1); streams R, S, G carry route_i, search_i, gather_i in order.
    Dependencies: route_i waits for gather_(i-2) (its set's bookkeeping is free, and every owner has finished searching that
    set: the gather waited for all of them); search_i waits for route_i of EVERY rank (barrier on channel 0 of stream R);
    gather_i waits for search_i of every rank (barrier on channel 1 of stream S).  Every rank issues the same sequence per
    stream.  The kernels are sized (cc_set_option route/search/gather blocks per SM) so that all three fit on an SM together.
    """

    def __init__(self, graph, splitters, rank: int, world: int, device, sub_batch: int, k: int, shard_first=None, group=None,
                 per_sm=(3, 1, 2)):
        self.sub = int(sub_batch)
        self.sets = [RoutedLookup(graph, splitters, rank, world, device, int(self.sub / world * 1.25) + 4096 if world > 1 else self.sub, k,
                                  shard_first=shard_first, group=group, max_batch=self.sub) for _ in range(2)]
        self.sr = torch.cuda.Stream(device=device, priority=-1)     # the NVLink legs first: they are short and hide behind the search
        self.sg = torch.cuda.Stream(device=device, priority=-1)
        self.ss = torch.cuda.Stream(device=device)
        self.world = world
        self.route_per_sm, self.search_per_sm, self.gather_per_sm = per_sm

    def find_packed(self, words: torch.Tensor, flags: torch.Tensor | None, out: torch.Tensor) -> torch.Tensor:
        nq = words.shape[0]
        nsub = max(1, (nq + self.sub - 1) // self.sub)
        if self.world > 1:
            # every rank must run the same number of sub-batches (the barriers are collective)
            t = torch.tensor([nsub], dtype=torch.int64, device=words.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            nsub = int(t.item())
        cur = torch.cuda.current_stream()
        for st in (self.sr, self.ss, self.sg):
            st.wait_stream(cur)
        routed = [torch.cuda.Event() for _ in range(nsub)]
        searched = [torch.cuda.Event() for _ in range(nsub)]
        gathered = [torch.cuda.Event() for _ in range(nsub)]
        N.set_option("route_blocks_per_sm", self.route_per_sm)
        N.set_option("routed_search_blocks_per_sm", self.search_per_sm)
        N.set_option("gather_blocks_per_sm", self.gather_per_sm if self.gather_per_sm > 0 else 16)

        def part(i):
            lo, hi = min(i * self.sub, nq), min((i + 1) * self.sub, nq)
            return words[lo:hi], (flags[lo:hi] if flags is not None else None), out[lo:hi]

        try:
            for i in range(nsub):
                w, f, o = part(i)
                rl = self.sets[i & 1]
                with torch.cuda.stream(self.sr):
                    if i >= 2:
                        self.sr.wait_event(gathered[i - 2])
                    rl.route(w, f)
                    rl._barrier(0)
                    routed[i].record(self.sr)
                with torch.cuda.stream(self.ss):
                    self.ss.wait_event(routed[i])
                    rl.search()
                    rl._barrier(1)
                    searched[i].record(self.ss)
                with torch.cuda.stream(self.sg):
                    self.sg.wait_event(searched[i])
                    rl.gather(o)
                    gathered[i].record(self.sg)
        finally:
            N.set_option("route_blocks_per_sm", 0)
            N.set_option("gather_blocks_per_sm", 16)
            N.set_option("routed_search_blocks_per_sm", 0)
        for st in (self.sr, self.ss, self.sg):
            cur.wait_stream(st)
        return out

    def check_overflow(self):
        for rl in self.sets:
            rl.check_overflow()

