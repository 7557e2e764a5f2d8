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
