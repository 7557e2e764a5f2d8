#!/usr/bin/env python3
"""bench.py -- the novel-k-mer scan (FindROIs) and batched k-mer lookups on B200, one JSON line.

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU algorithm (oracle port), host cores

Headline metric (BASELINE.json): novelty-scan records/s on configs[1] -- a synthetic Plasmodium-sized trio graph
(2.5e7 records, k=47, 4 colours: child + 2 parents + ref; 36-byte records, 0.9 GB, larger than the 126 MB L2 so
no flush is needed between steps).  A step = one pass of the fused decode+filter+compact kernel over one rank's
record array.  Under torchrun each rank holds its own shard of 2.5e7 records (records are independent: no
data-path collective; weak scaling) and `value` = records of all ranks / max-over-ranks time.
`e2e` = the same scan through the host-buffer C-ABI call (cc_find_novel_host): the record array lives in pinned
host memory and is streamed through the device, novel records come back to the host, all inside the timed region.
`lookup` = batched findRecord throughput (configs[2] shape, table 1e8 records k=47) as an extra object.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K, COLORS = 47, 4
S_WORDS = 2
REC_BYTES = 8 * S_WORDS + 5 * COLORS          # 36
OUT_BYTES = 8 * S_WORDS + 5                   # 21
SEED_SCAN, SEED_LOOKUP = 20261018, 20261019
METRIC = "novel_kmer_scan_records_per_s"
UNIT = "records/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--records", type=int, default=25_000_000, help="records per rank (configs[1] = 2.5e7)")
    ap.add_argument("--table", type=int, default=100_000_000, help="lookup table records (configs[2] = 1e8)")
    ap.add_argument("--queries", type=int, default=1_000_000_000, help="lookup queries per step, all ranks together (configs[2] = 1e9)")
    ap.add_argument("--no-lookup", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lookup-only", action="store_true", help="experiments: only the lookup object (not a bench line)")
    ap.add_argument("--pipeline", type=int, default=0, help="experiments: also time the sub-batch pipelined routed lookup with this many sub-batches")
    ap.add_argument("--no-compare", action="store_true", help="skip the replicated-table and NCCL comparison modes of the lookup leg")
    ap.add_argument("--spinup", type=float, default=1.0, help="seconds of untimed back-to-back scans before warm-up")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm = sorted(int(float(r[0])) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in rows if len(r) >= 3 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(rows[0][1])) if rows else None,
                "reasons": reasons, "samples": len(rows), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpu_scan_rate(body: np.ndarray, n: int, threads: int, faithful: bool, min_seconds: float = 4.0, max_passes: int = 50):
    """The oracle port of FindROIs' loop (CortexGraph.getNextRecord + isNovel + writer layout) over `body`,
    record ranges split across `threads` host threads (ctypes releases the GIL).  Returns (records/s, passes, novel)."""
    from oracle import orc
    orc.lib()
    bounds = [n * i // threads for i in range(threads + 1)]
    counts = [0] * threads

    def work(t):
        lo, hi = bounds[t], bounds[t + 1]
        cnt, _, _ = orc.find_rois_body(body[lo * REC_BYTES:hi * REC_BYTES], hi - lo, K, S_WORDS, COLORS, 0, [1, 2, 3],
                                       faithful=faithful, cap=max(1024, (hi - lo) // 16))
        counts[t] = cnt

    passes, t0 = 0, time.perf_counter()
    while True:
        th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        [x.start() for x in th]
        [x.join() for x in th]
        passes += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or passes >= max_passes:
            break
    return n * passes / el, passes, sum(counts)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    import torch
    from tools import synth
    threads = os.cpu_count() or 1
    n = args.records
    body, _ = synth.make_graph_body(SEED_SCAN, n, K, COLORS, device="cpu")
    flat = body.numpy().reshape(-1)
    for _ in range(min(args.warmup, 1)):
        cpu_scan_rate(flat, n, threads, True, min_seconds=0.0, max_passes=1)
    rates = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, _, novel = cpu_scan_rate(flat, n, threads, True, min_seconds=0.0, max_passes=1)
        rates.append(r)
        if time.perf_counter() - t_all > 120:
            break
    steps = len(rates)
    ms = 1000.0 * sum(n / r for r in rates) / steps
    value = n / (ms / 1000.0)
    sample = ("oracle port (C restatement of the Java loop incl. the per-record k-mer string decode the reference does for "
              "its LRU; JVM allocation not simulated; the Java reference itself is single-threaded and cannot be built "
              "here: no JDK) over all %d records per step, record ranges split over %d threads" % (n, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic Plasmodium-sized trio graph, %d records, k=47, 4 colours, novelty scan" % n,
                   "records": n, "kmer_size": K, "colors": COLORS, "record_bytes": REC_BYTES, "novel_records": int(novel)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    import corticall_b200 as cb
    from corticall_b200 import _native as N
    from tools import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = N.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if args.lookup_only:
        lookup = bench_lookup(args, rank, world, local_rank, dev, barrier, max_over_ranks)
        if rank == 0:
            print(json.dumps({"lookup_only": True, "n_gpus": world, "lookup": lookup}))
        return

    # ------------------------------------------------------------ the scan workload (configs[1]), one shard per rank
    n = args.records
    body, _ = synth.make_graph_body(SEED_SCAN + 1000 * rank, n, K, COLORS, device=dev)
    g = cb.CortexGraph.fromDevice(body.data_ptr(), K, COLORS, n, firstIndex=rank * n, device=local_rank, keepalive=body)
    parents = np.array([1, 2, 3], dtype=np.int32)
    cap = max(1 << 20, n // 16)
    out = torch.empty(cap * OUT_BYTES + 64, dtype=torch.uint8, device=dev)
    cnt_dev = torch.zeros(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def scan_step():
        N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), None, cap, cnt_dev.data_ptr(), stream))

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_load0 = time.perf_counter()
    # spin-up: about a second of back-to-back scans so the clock record reflects the loaded state
    scan_step()
    torch.cuda.synchronize()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < args.spinup:
        for _ in range(20):
            scan_step()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        scan_step()
    barrier()
    launches0 = cb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        scan_step()
        ev[i + 1].record()
    barrier()
    launches = cb.launch_count() - launches0
    t_load1 = time.perf_counter()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    novel = int(cnt_dev[0].item())
    ms_step = max_over_ranks(total_ms / args.steps)
    value = world * n / (ms_step / 1000.0)
    clocks = sampler.stop(t_load0, t_load1) if sampler else None

    peak, peak_src = peaks()
    alg_bytes = n * REC_BYTES + novel * OUT_BYTES
    kernel_ms = total_ms / args.steps            # a step is exactly one launch of scan_novel_kernel on this stream
    achieved = alg_bytes / (kernel_ms / 1000.0) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "scan_novel_fast_kernel (+ no-op redo_chunks_kernel launch)",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": kernel_ms, "kernel_ms_min": per_launch[0],
                "kernel_ms_median": per_launch[len(per_launch) // 2],
                "frac_of_nominal_8TBs": achieved / 8000.0}
    traffic_file = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ------------------------------------------------------------ e2e: host-resident records through the C ABI
    e2e_steps = max(3, min(args.steps, 10))
    host = torch.empty(n * REC_BYTES, dtype=torch.uint8, pin_memory=True)
    host.copy_(body.reshape(-1))
    torch.cuda.synchronize()
    hcap = max(1 << 20, n // 16)
    hout = np.empty(hcap * OUT_BYTES, dtype=np.uint8)
    hcnt = C.c_uint64()
    st = N.Stats()

    def e2e_step():
        N.check(L.cc_find_novel_host(local_rank, host.data_ptr(), K, S_WORDS, COLORS, n, 0, parents.ctypes.data, 3,
                                     hout.ctypes.data, None, hcap, C.byref(hcnt), C.byref(st)))

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                      # synchronous: returns when the novel records are in host memory
    t1 = time.perf_counter()
    e2e_ms = max_over_ranks((t1 - t0) * 1000.0 / e2e_steps)
    assert hcnt.value == novel
    e2e = {"value": world * n / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(st.h2d_bytes),
           "d2h_bytes_per_step": int(st.d2h_bytes), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "cc_find_novel_host (pinned host record array -> device in chunks -> novel records back to host)"}
    del host

    # ------------------------------------------------------------ CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        flat = body.reshape(-1).cpu().numpy()
        threads = os.cpu_count() or 1
        r1, p1, cn = cpu_scan_rate(flat, n, 1, True, min_seconds=3.0, max_passes=3)
        rmt, pm, _ = cpu_scan_rate(flat, n, threads, True, min_seconds=3.0, max_passes=40)
        # the same loop WITHOUT the reference's per-record k-mer string decode -- what a tuned CPU port would do
        rpk, _, cn2 = cpu_scan_rate(flat, n, threads, False, min_seconds=2.0, max_passes=40)
        assert cn == novel and cn2 == novel, "oracle and GPU disagree on the novel count"
        cpu = {"value": rmt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle port of the Java FindROIs loop (faithful: incl. per-record k-mer string decode), all %d records, "
                         "%d passes on %d threads; single thread: %.3g records/s (%d passes). The Java reference is single-threaded "
                         "and cannot be built here (no JDK)." % (n, pm, threads, r1, p1),
               "value_single_thread": r1, "value_all_threads_without_string_decode": rpk}
        del flat
    g.dispose()
    del body, out

    # ------------------------------------------------------------ lookups (configs[2] shape) -- extra object
    lookup = None
    if not args.no_lookup:
        lookup = bench_lookup(args, rank, world, local_rank, dev, barrier, max_over_ranks)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic Plasmodium-sized trio graph, %d records per GPU, k=47, 4 colours "
                                   "(child, mom, dad, ref), novelty scan child vs 3" % n,
                       "records_per_gpu": n, "kmer_size": K, "colors": COLORS, "record_bytes": REC_BYTES,
                       "novel_records": novel, "l2": "input %.2f GB per step > 126 MB L2, no flush needed" % (n * REC_BYTES / 1e9),
                       "sharding": "contiguous record ranges per GPU, no data-path collective"},
            "gb_per_s": value * (alg_bytes / n) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "lookup": lookup,
        }
        print(json.dumps(line))


def bench_lookup(args, rank, world, local_rank, dev, barrier, max_over_ranks):
    """Batched findRecord on the configs[2] shape.  N=1: the whole table on one GPU.  N>1: the sorted table is sharded
    by k-mer range, each rank owns Q/N queries, routes them to the owning shard (bucket kernel + all-to-all over
    NCCL), searches locally and returns the indices (reverse all-to-all + scatter)."""
    import torch

    import corticall_b200 as cb
    from corticall_b200 import _native as N
    from corticall_b200.host.sharded import ShardedLookup
    from tools import synth

    L = N.lib()
    nt, nq_total = args.table, args.queries
    words = synth.random_canonical_keys(SEED_LOOKUP, nt, K, dev)          # identical on every rank (counter-based)
    lo, hi = nt * rank // world, nt * (rank + 1) // world
    cov, edges = synth.coverage_and_edges(SEED_LOOKUP, hi - lo, COLORS, dev, offset=lo)
    body = synth.assemble_records([w[lo:hi] for w in words], cov, edges)
    del cov, edges
    splitters = torch.stack([torch.stack([w[nt * r // world] for w in words]) for r in range(1, world)]) if world > 1 else None
    g = cb.CortexGraph.fromDevice(body.data_ptr(), K, COLORS, hi - lo, firstIndex=lo, device=local_rank, keepalive=body)
    g.buildIndex()
    # queries owned by this rank: uniform mode (half hits, random strand, 0.1 % N, 0.1 % lower case)
    nq = nq_total // world
    chunk = 1 << 24
    qwords = torch.empty((nq, S_WORDS), dtype=torch.int64, device=dev)
    qflags = torch.empty(nq, dtype=torch.uint8, device=dev)
    n_ascii = min(nq, 1 << 26)
    qascii = torch.empty((n_ascii, K), dtype=torch.uint8, device=dev)
    for o in range(0, nq, chunk):
        m = min(chunk, nq - o)
        a, canon, valid = synth.make_queries(SEED_LOOKUP, words, K, m, offset=rank * nq + o)
        qwords[o:o + m, 0], qwords[o:o + m, 1] = canon[0], canon[1]
        qflags[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
        if o < n_ascii:
            qascii[o:min(o + m, n_ascii)] = a[:max(0, min(m, n_ascii - o))]
        del a, canon, valid
    del words
    torch.cuda.synchronize()
    res = torch.empty(nq, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    out = {"table_records": nt, "queries_per_step": nq * world, "kmer_size": K,
           "scaling": "strong (fixed table and batch, sharded by k-mer range)" if world > 1 else "single GPU"}
    steps, warm = 5, 2

    def timeit(fn):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)

    if world == 1:
        for name, algo in (("bucketed", cb.CC_ALGO_AUTO), ("bsearch", cb.CC_ALGO_BSEARCH)):
            ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qwords.data_ptr(), qflags.data_ptr(), nq, res.data_ptr(), algo, stream)))
            out["packed_%s_lookups_per_s" % name] = nq / (ms / 1000.0)
        hits = int((res >= 0).sum().item())
        out["hit_fraction"] = hits / nq
        ms = timeit(lambda: N.check(L.cc_find_ascii_dev(g._h, qascii.data_ptr(), n_ascii, res.data_ptr(), cb.CC_ALGO_AUTO, stream)))
        out["ascii_lookups_per_s"] = n_ascii / (ms / 1000.0)
        out["ascii_queries_per_step"] = n_ascii
        pw = torch.empty((n_ascii, S_WORDS), dtype=torch.int64, device=dev)
        pf = torch.empty(n_ascii, dtype=torch.uint8, device=dev)
        ms = timeit(lambda: N.check(L.cc_pack_kmers_dev(local_rank, qascii.data_ptr(), n_ascii, K, pw.data_ptr(), pf.data_ptr(), stream)))
        out["pack_rows_per_s"] = n_ascii / (ms / 1000.0)                 # K3 on independent k-byte rows
        out["pack_rows_algorithmic_gb_per_s"] = out["pack_rows_per_s"] * (K + 8 * S_WORDS + 1) / 1e9
        del pw, pf
        nm = min(nq, 1 << 26)
        ms = timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qwords.data_ptr(), qflags.data_ptr(), nm, res.data_ptr(), cb.CC_ALGO_MERGE, stream)))
        out["packed_sorted_merge_lookups_per_s"] = nm / (ms / 1000.0)
        out["lookups_per_s"] = out["packed_bucketed_lookups_per_s"]
        # algorithmic bytes per lookup: 8s query + 8 result + one pass over the key column amortised
        bpl = 8 * S_WORDS + 8 + nt * 8 * S_WORDS / nq
        out["algorithmic_bytes_per_lookup"] = bpl
        out["algorithmic_gb_per_s"] = out["lookups_per_s"] * bpl / 1e9
    else:
        from corticall_b200.host.sharded import RoutedLookup
        # segments sized for the balanced case (+25 %) instead of the worst case nq: 6x less peer-mapped address space per
        # rank (the worst-case layout cost 40 % on the search leg of most ranks: TLB reach); overflow is checked per batch
        rl = RoutedLookup(g, splitters, rank, world, dev, cap=int(nq / world * 1.25) + 4096, k=K, max_batch=nq)
        ms = timeit(lambda: rl.find_packed(qwords, qflags, res))
        out["lookups_per_s"] = nq * world / (ms / 1000.0)
        out["ms_per_step"] = ms
        out["exchange"] = ("peer memory over NVLink, one kernel per leg: route (owner + P2P store into the owner's inbox) -> barrier -> "
                           "search (P2P store of results into the origin) -> barrier -> gather")
        out["hit_fraction"] = float((res >= 0).sum().item()) / nq
        barrier()
        rl.find_packed(qwords, qflags, res, profile=True)
        out["phase_ms_rank0"] = rl.phase_ms
        import torch.distributed as dist
        allp = [None] * world
        dist.all_gather_object(allp, {k: round(v, 3) for k, v in rl.phase_ms.items()})
        out["phase_ms_all_ranks"] = allp
        del rl
        if args.pipeline > 0:
            from corticall_b200.host.sharded import PipelinedRoutedLookup
            for per in ((1, 1, 2), (1, 1, 3), (2, 2, 2)):
                pl = PipelinedRoutedLookup(g, splitters, rank, world, dev, sub_batch=(nq + args.pipeline - 1) // args.pipeline, k=K)
                pl.route_per_sm, pl.gather_per_sm, pl.search_per_sm = per
                res3 = torch.empty_like(res)
                msp = timeit(lambda: pl.find_packed(qwords, qflags, res3))
                key = "pipelined_%d_r%d_g%d_s%d" % ((args.pipeline,) + per)
                out[key + "_lookups_per_s"] = nq * world / (msp / 1000.0)
                out[key + "_agrees"] = bool(torch.equal(res, res3))
                del pl, res3
        if args.no_compare:
            g.dispose()
            return out
        # replicas: when the whole table fits every GPU (it does for configs[2]: 3.6 GB of records + 2.1 GB of index), each
        # rank can hold a full copy and look up its own queries with no exchange at all (SURVEY 8e "alternative mode")
        try:
            words_all = synth.random_canonical_keys(SEED_LOOKUP, nt, K, dev)
            covf, edgf = synth.coverage_and_edges(SEED_LOOKUP, nt, COLORS, dev)
            full_body = synth.assemble_records(words_all, covf, edgf)
            del covf, edgf, words_all
            gfull = cb.CortexGraph.fromDevice(full_body.data_ptr(), K, COLORS, nt, firstIndex=0, device=local_rank, keepalive=full_body)
            gfull.buildIndex()
            res4 = torch.empty_like(res)
            msr = timeit(lambda: N.check(L.cc_find_packed_dev(gfull._h, qwords.data_ptr(), qflags.data_ptr(), nq, res4.data_ptr(), cb.CC_ALGO_AUTO, stream)))
            out["replicated_table_lookups_per_s"] = nq * world / (msr / 1000.0)
            out["replicated_agrees"] = bool(torch.equal(res, res4))
            gfull.dispose()
            del full_body, res4
        except Exception as e:          # comparison mode only
            out["replicated_table_lookups_per_s"] = None
            out["replicated_error"] = str(e)[:200]
        # the NCCL all-to-all formulation of the same exchange, as the comparison and as a cross-check of the results
        sl = ShardedLookup(g, splitters, rank, world, dev)
        res2 = torch.empty_like(res)
        ms2 = timeit(lambda: sl.find_packed(qwords, qflags, res2))
        out["nccl_all_to_all_lookups_per_s"] = nq * world / (ms2 / 1000.0)
        out["paths_agree"] = bool(torch.equal(res, res2))
        sl.profile = True
        sl.find_packed(qwords, qflags, res2)
        out["nccl_phase_ms_rank0"] = sl.last.get("phase_ms")
    g.dispose()
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the k-mer hot path has no CPU fallback (use --impl reference for the CPU arm)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
