#!/usr/bin/env python3
"""bench.py -- the novel-k-mer scan (FindROIs) and batched k-mer lookups on B200, one JSON line.

  python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU algorithm (oracle port), host cores
  python bench.py --config 3|4 ...                           # only BASELINE.json configs[3] (human scale, needs torchrun
                                                             # with >= 2 ranks) or configs[4] (k=63, 21 colours, 1 GPU)

Headline metric (BASELINE.json): novelty-scan records/s on configs[1] -- a synthetic Plasmodium-sized trio graph
(2.5e7 records, k=47, 4 colours: child + 2 parents + ref; 36-byte records, 0.9 GB, larger than the 126 MB L2 so
no flush is needed between steps).  A step = one pass of the fused decode+filter+compact kernel over one rank's
record array.  Under torchrun each rank holds its own shard of 2.5e7 records (records are independent: no
data-path collective; weak scaling) and `value` = records of all ranks / max-over-ranks time.
`e2e` = the same scan through the host-buffer C-ABI call (cc_find_novel_host): the record array lives in pinned
host memory and is streamed through the device, novel records come back to the host, all inside the timed region.
Extra objects on the same line, each with its own parity check against the oracle inside the run:
  `lookup`   batched findRecord on configs[2] (table 1e8 records k=47, 1e9 queries per step; sharded by k-mer range and
             routed over NVLink at N > 1), with its own roofline / cpu_baseline / e2e;
  `configs4` the many-colour scan of configs[4] (k=63, 21 colours, 1e8 records of 121 bytes) at N = 1;
  `configs3` the human-scale graph of configs[3] (3.75e8 records of k=31 per GPU = 3e9 on 8 GPUs): scan + routed lookups, N > 1;
  `composite` open a .ctx file -> FindROIs -> contig lookups with the graph resident across the calls, against the CPU port.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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
SEED_SCAN, SEED_LOOKUP, SEED_HUMAN, SEED_MANY = 20261018, 20261019, 20261020, 20261021
METRIC = "novel_kmer_scan_records_per_s"
UNIT = "records/s"
WORKLOAD = ("configs[1]: synthetic Plasmodium-sized trio graph, 25000000 records per GPU, k=47, 4 colours "
            "(child, mom, dad, ref), novelty scan child vs 3")


def workload_string(n: int) -> str:
    return WORKLOAD.replace("25000000", str(n))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=-1, choices=[-1, 3, 4], help="only this BASELINE.json config (3 = human scale, 4 = many colours)")
    ap.add_argument("--records", type=int, default=25_000_000, help="records per rank (configs[1] = 2.5e7)")
    ap.add_argument("--table", type=int, default=100_000_000, help="lookup table records (configs[2] = 1e8)")
    ap.add_argument("--queries", type=int, default=1_000_000_000, help="lookup queries per step, all ranks together (configs[2] = 1e9)")
    ap.add_argument("--human-records", type=int, default=375_000_000, help="configs[3]: records per GPU (3.75e8 x 8 GPUs = 3e9)")
    ap.add_argument("--human-queries", type=int, default=125_000_000, help="configs[3]: lookup queries per GPU and step")
    ap.add_argument("--many-records", type=int, default=100_000_000, help="configs[4]: records (1e8 x 121 B = 12.1 GB)")
    ap.add_argument("--no-lookup", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip configs3 / configs4 / composite")
    ap.add_argument("--lookup-only", action="store_true", help="experiments: only the lookup object (not a bench line)")
    ap.add_argument("--no-compare", action="store_true", help="skip the replicated-table and NCCL comparison modes of the lookup leg")
    ap.add_argument("--spinup", type=float, default=1.0, help="seconds of untimed back-to-back scans before warm-up")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def source_stamp() -> str:
    """Hash of what determines the profiled kernels -- the kernel sources and the default tuning options: ncu-derived numbers under
    profiles/ carry it and are ignored once any of that has changed."""
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, "corticall_b200", "csrc")
    for name in ("device_utils.cuh", "lookup.cu", "scan.cu"):
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    with open(os.path.join(d, "cc_internal.hpp")) as f:
        m = re.search(r"struct Options \{.*?\n\};", f.read(), flags=re.S)
        h.update(m.group(0).encode() if m else b"?")
    return h.hexdigest()[:16]


def measured_traffic(kernel: str):
    """dram bytes per launch of `kernel` from profiles/traffic.json (written by tools/ncu_traffic.py from an ncu capture),
    or None when the file is absent or was taken on other sources."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t["kernels"][kernel]
        if t.get("source_stamp") != source_stamp():
            return None, "stale: profiles/traffic.json was captured on other kernel sources"
        return e, None
    except Exception:
        return None, "no ncu capture on file"


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
def cpu_scan_pass(body: np.ndarray, n: int, k: int, s: int, c: int, parents, threads: int, faithful: bool, keep_output: bool = False):
    """One pass of the oracle port of FindROIs' loop (CortexGraph.getNextRecord + isNovel + writer layout) over `body`, record
    ranges split across `threads` host threads (ctypes releases the GIL).  Returns (seconds, novel count, [(records, indices)] per range)."""
    from oracle import orc
    orc.lib()
    S = 8 * s + 5 * c
    bounds = [n * i // threads for i in range(threads + 1)]
    res = [None] * threads

    def work(t):
        lo, hi = bounds[t], bounds[t + 1]
        cnt, recs, idx = orc.find_rois_body(body[lo * S:hi * S], hi - lo, k, s, c, 0, parents, faithful=faithful,
                                            cap=(hi - lo) if keep_output else max(1024, (hi - lo) // 16))
        res[t] = (cnt, recs, idx + lo) if keep_output else (cnt, None, None)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [x.start() for x in th]
    [x.join() for x in th]
    return time.perf_counter() - t0, sum(r[0] for r in res), res


def cpu_scan_rate(body, n, threads, faithful, min_seconds=4.0, max_passes=50):
    passes, el, novel = 0, 0.0, 0
    while True:
        dt, novel, _ = cpu_scan_pass(body, n, K, S_WORDS, COLORS, [1, 2, 3], threads, faithful)
        passes += 1
        el += dt
        if el >= min_seconds or passes >= max_passes:
            break
    return n * passes / el, passes, novel


def cpu_find_rate(og, kmers: np.ndarray, threads: int):
    """Oracle findRecord (CortexGraph.java:272-317 restated) over `kmers`, split across threads.  Returns (lookups/s, indices)."""
    nq = kmers.shape[0]
    out = np.empty(nq, dtype=np.int64)
    bounds = [nq * i // threads for i in range(threads + 1)]

    def work(t):
        lo, hi = bounds[t], bounds[t + 1]
        if hi > lo:
            out[lo:hi] = og.find_batch(kmers[lo:hi])

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [x.start() for x in th]
    [x.join() for x in th]
    return nq / (time.perf_counter() - t0), out


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from tools import synth
    threads = os.cpu_count() or 1
    n = args.records
    body, _ = synth.make_graph_body(SEED_SCAN, n, K, COLORS, device="cpu")
    flat = body.numpy().reshape(-1)
    warm = max(args.warmup, 3)
    for _ in range(min(warm, 2)):                  # the CPU arm needs no more than page-in + thread start; the count is reported as run
        cpu_scan_pass(flat, n, K, S_WORDS, COLORS, [1, 2, 3], threads, True)
    secs = []
    novel = 0
    t_all = time.perf_counter()
    for _ in range(args.steps):
        dt, novel, _ = cpu_scan_pass(flat, n, K, S_WORDS, COLORS, [1, 2, 3], threads, True)
        secs.append(dt)
        if time.perf_counter() - t_all > 120:
            break
    steps = len(secs)
    ms = 1000.0 * sum(secs) / steps
    value = n / (ms / 1000.0)
    sample = ("oracle port (C restatement of the Java loop incl. the per-record k-mer string decode the reference does for "
              "its LRU; JVM allocation not simulated; the Java reference itself is single-threaded and cannot be built "
              "here: no JDK) over all %d records per step, record ranges split over %d threads" % (n, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "warmup_passes_run": min(warm, 2), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_string(n), "records_per_gpu": n, "kmer_size": K, "colors": COLORS, "record_bytes": REC_BYTES,
                   "novel_records": int(novel)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------- GPU arm
class Env:
    """Rank plumbing shared by the legs."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.torch, self.dist = torch, dist
        self.dev = torch.device("cuda", local_rank)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x: float, op) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def timeit(self, fn, steps=5, warm=3):
        """Device time per call (CUDA events on the current stream, barrier on both sides, max over ranks)."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max(e0.elapsed_time(e1) / steps)


def torch_expected_novel(torch, body, s: int, c: int, child: int, parents):
    """An independent formulation of FindROIs.isNovel + projection with torch ops on the record matrix [n, S]:
    (indices int64, records uint8 [m, 8s+5]).  Coverage is a SIGNED Java int (BinaryUtils.java:6-17)."""
    n = body.shape[0]
    cov = body[:, 8 * s:8 * s + 4 * c].contiguous().view(torch.int32).reshape(n, c)
    mask = cov[:, child] > 0
    for p in parents:
        mask &= cov[:, p] == 0
    idx = torch.nonzero(mask).reshape(-1)
    rec = torch.cat([body[idx, :8 * s], body[idx, 8 * s + 4 * child:8 * s + 4 * child + 4], body[idx, 8 * s + 4 * c + child:8 * s + 4 * c + child + 1]], dim=1)
    return idx, rec


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch

    import corticall_b200 as cb
    from corticall_b200 import _native as N
    from tools import synth

    env = Env(rank, world, local_rank)
    torch.cuda.set_device(local_rank)
    dev = env.dev
    L = N.lib()

    if args.lookup_only:
        lookup = bench_lookup(args, env)
        if rank == 0:
            print(json.dumps({"lookup_only": True, "n_gpus": world, "lookup": lookup}))
        return
    if args.config == 4:
        out = bench_many_colour(args, env)
        if rank == 0:
            print(json.dumps({"config": "configs[4]", "n_gpus": world, "configs4": out}))
        return
    if args.config == 3:
        out = bench_human_scale(args, env)
        if rank == 0:
            print(json.dumps({"config": "configs[3]", "n_gpus": world, "configs3": out}))
        return

    # ------------------------------------------------------------ the scan workload (configs[1]), one shard per rank
    n = args.records
    body, _ = synth.make_graph_body(SEED_SCAN + 1000 * rank, n, K, COLORS, device=dev)
    g = cb.CortexGraph.fromDevice(body.data_ptr(), K, COLORS, n, firstIndex=rank * n, device=local_rank, keepalive=body)
    parents = np.array([1, 2, 3], dtype=np.int32)
    cap = max(1 << 20, n // 16)
    out = torch.empty(cap * OUT_BYTES + 64, dtype=torch.uint8, device=dev)
    out_idx = torch.empty(cap, dtype=torch.int64, device=dev)
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
    warm = max(args.warmup, 3)
    for _ in range(warm):
        scan_step()
    env.barrier()
    launches0 = cb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        scan_step()
        ev[i + 1].record()
    env.barrier()
    launches = cb.launch_count() - launches0
    t_load1 = time.perf_counter()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    novel = int(cnt_dev[0].item())
    ms_step = env.max(total_ms / args.steps)
    value = world * n / (ms_step / 1000.0)
    clocks = sampler.stop(t_load0, t_load1) if sampler else None

    # ---- parity inside the run: the whole output, bit for bit, against an independent torch formulation of the predicate ...
    N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), out_idx.data_ptr(), cap, cnt_dev.data_ptr(), stream))
    torch.cuda.synchronize()
    want_idx, want_rec = torch_expected_novel(torch, body, S_WORDS, COLORS, 0, [1, 2, 3])
    scan_parity = {"novel_records": novel,
                   "torch_formulation_bit_exact": bool(novel == want_idx.numel() and torch.equal(out[:novel * OUT_BYTES].reshape(-1, OUT_BYTES), want_rec)
                                                       and torch.equal(out_idx[:novel], want_idx + rank * n))}
    # ---- ... and the globally ordered output of the sharded scan: counts all-gathered, exclusive offsets, every
    # rank's slice lands at its offset; an order-dependent 128-bit hash of (global position, record bytes) is summed over ranks
    counts = torch.tensor([novel], dtype=torch.int64, device=dev)
    if world > 1:
        allc = [torch.empty_like(counts) for _ in range(world)]
        env.dist.all_gather(allc, counts)
        offset = int(sum(int(x.item()) for x in allc[:rank]))
        total_novel = int(sum(int(x.item()) for x in allc))
    else:
        offset, total_novel = 0, novel
    scan_parity["global_output"] = {"offset_of_this_rank": offset, "total_novel_all_ranks": total_novel,
                                    "first_index_rank0": int(out_idx[0].item()) if novel else None}
    del want_idx, want_rec

    peak, peak_src = peaks()
    alg_bytes = n * REC_BYTES + novel * OUT_BYTES
    kernel_ms = total_ms / args.steps            # a step = scan_novel_fast_kernel + the (empty) redo_chunks_kernel behind it
    achieved = alg_bytes / (kernel_ms / 1000.0) / 1e9
    traffic, traffic_note = measured_traffic("scan_novel_fast_kernel")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_note": traffic_note or traffic.get("note"),
                "peak_source": peak_src, "kernel": "scan_novel_fast_kernel (+ the empty redo_chunks_kernel launched behind it with PDL)",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": kernel_ms, "kernel_ms_min": per_launch[0],
                "kernel_ms_median": per_launch[len(per_launch) // 2],
                "frac_of_nominal_8TBs": achieved / 8000.0}

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
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                      # synchronous: returns when the novel records are in host memory
    t1 = time.perf_counter()
    e2e_ms = env.max((t1 - t0) * 1000.0 / e2e_steps)
    assert hcnt.value == novel
    scan_parity["e2e_output_equals_device_output"] = bool(np.array_equal(hout[:novel * OUT_BYTES], out[:novel * OUT_BYTES].cpu().numpy()))
    e2e = {"value": world * n / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(st.h2d_bytes),
           "d2h_bytes_per_step": int(st.d2h_bytes), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_gb_per_s_per_gpu": st.h2d_bytes / (e2e_ms / 1000.0) / 1e9,
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
        # the oracle's whole output against the device's, bit for bit (records and indices)
        _, _, parts = cpu_scan_pass(flat, n, K, S_WORDS, COLORS, [1, 2, 3], threads, False, keep_output=True)
        o_rec = np.concatenate([p[1] for p in parts])
        o_idx = np.concatenate([p[2] for p in parts]).astype(np.int64)
        scan_parity["oracle_bit_exact"] = bool(np.array_equal(o_rec, out[:novel * OUT_BYTES].cpu().numpy()) and
                                               np.array_equal(o_idx, out_idx[:novel].cpu().numpy()))
        cpu = {"value": rmt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle port of the Java FindROIs loop (faithful: incl. per-record k-mer string decode), all %d records, "
                         "%d passes on %d threads; single thread: %.3g records/s (%d passes). The Java reference is single-threaded "
                         "and cannot be built here (no JDK)." % (n, pm, threads, r1, p1),
               "value_single_thread": r1, "value_all_threads_without_string_decode": rpk}
        del flat, parts, o_rec, o_idx
    g.dispose()
    del body, out, out_idx

    # ------------------------------------------------------------ extra objects
    def guarded(fn, *a):
        try:
            return fn(*a)
        except Exception as e:          # an extra leg must never take the headline line down with it
            import traceback
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300]), "trace": traceback.format_exc()[-600:]}
        finally:
            torch.cuda.empty_cache()

    lookup = None if args.no_lookup else guarded(bench_lookup, args, env)
    composite = configs4 = configs3 = None
    if not args.no_extra:
        if world == 1:
            composite = guarded(bench_composite, args, env)
            configs4 = guarded(bench_many_colour, args, env)
        else:
            configs3 = guarded(bench_human_scale, args, env)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": workload_string(n),
                       "records_per_gpu": n, "kmer_size": K, "colors": COLORS, "record_bytes": REC_BYTES,
                       "novel_records": novel, "l2": "input %.2f GB per step > 126 MB L2, no flush needed" % (n * REC_BYTES / 1e9),
                       "sharding": "contiguous record ranges per GPU, no data-path collective; counts all-gathered into exclusive offsets"},
            "gb_per_s": value * (alg_bytes / n) / 1e9,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity": scan_parity, "lookup": lookup, "composite": composite, "configs4": configs4, "configs3": configs3,
        }
        print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- lookups (configs[2])
def oracle_check_sharded(env, table_body, first_index, k, c, splitters, qwords, qflags, res, sample=1_000_000, threads=8):
    """Oracle parity of a sharded lookup on the real multi-GPU path: every rank contributes `sample` of its queries and
    their answers; each rank runs the oracle's findRecord over ITS OWN shard for the sampled queries it owns (owner by an
    independent torch formulation of the splitter rule) and compares with what the routed path answered.  With one rank
    this is the plain single-GPU check.  Returns (checked, mismatches) summed over ranks."""
    torch = env.torch
    from oracle import orc
    from tools import synth
    s = qwords.shape[1]
    m = min(sample, qwords.shape[0])
    qs, rs, fs = qwords[:m].contiguous(), res[:m].contiguous(), qflags[:m].contiguous()
    if env.world > 1:
        aq = [torch.empty_like(qs) for _ in range(env.world)]
        ar = [torch.empty_like(rs) for _ in range(env.world)]
        af = [torch.empty_like(fs) for _ in range(env.world)]
        env.dist.all_gather(aq, qs)
        env.dist.all_gather(ar, rs)
        env.dist.all_gather(af, fs)
        qs, rs, fs = torch.cat(aq), torch.cat(ar), torch.cat(af)
    # flagged queries (an N or a lower-case base in the ASCII k-mer) must answer -1 whatever their packed words say
    flagged_bad = int(((fs != 0) & (rs != -1)).sum().item())
    live = fs == 0
    qs, rs = qs[live], rs[live]
    words = [qs[:, w].contiguous() for w in range(s)]
    mine = synth.owner_of_keys(words, splitters) == env.rank
    wq = [w[mine] for w in words]
    got = rs[mine].cpu().numpy()
    ascii_q = synth.words_to_ascii(wq, k).cpu().numpy()
    n_local = table_body.shape[0]
    hdr = synth.header_bytes(k, c)
    image = np.empty(len(hdr) + n_local * table_body.shape[1], dtype=np.uint8)
    image[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
    torch.from_numpy(image[len(hdr):]).copy_(table_body.reshape(-1))
    og = orc.Graph(image)
    _, want_local = cpu_find_rate(og, ascii_q, threads)
    want = np.where(want_local >= 0, want_local + first_index, -1)
    bad = int((want != got).sum()) + (flagged_bad if env.rank == 0 else 0)
    # queries this rank does NOT own must not have been answered with an index inside this rank's shard
    other = rs[~mine]
    bad += int(((other >= first_index) & (other < first_index + n_local)).sum().item())
    del image, og
    return int(env.sum(float(ascii_q.shape[0]))), int(env.sum(float(bad)))


def bench_lookup(args, env):
    """Batched findRecord on the configs[2] shape.  N=1: the whole table on one GPU.  N>1: the sorted table is sharded
    by k-mer range, each rank owns Q/N queries, routes them to the owning shard over peer memory (NVLink), the owner
    searches its bucket-line index and the origin pulls the results back."""
    torch = env.torch
    import corticall_b200 as cb
    from corticall_b200 import _native as N
    from corticall_b200.host.sharded import ShardedLookup
    from tools import synth

    rank, world, local_rank, dev = env.rank, env.world, env.local_rank, env.dev
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
    bpl = 8 * S_WORDS + 8 + nt * 8 * S_WORDS / (nq * world)   # 8s query + 8 result + one pass over the key column amortised
    peak, peak_src = peaks()

    if world == 1:
        ms = env.timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qwords.data_ptr(), qflags.data_ptr(), nq, res.data_ptr(), cb.CC_ALGO_AUTO, stream)))
        out["lookups_per_s"] = out["packed_line_index_lookups_per_s"] = nq / (ms / 1000.0)
        out["ms_per_step"] = ms
        out["hit_fraction"] = float((res >= 0).sum().item()) / nq
        checked, bad = oracle_check_sharded(env, body, lo, K, COLORS, None, qwords, qflags, res, threads=os.cpu_count() or 1)
        out["parity"] = {"oracle_checked_queries": checked, "oracle_mismatches": bad}
        ref = torch.empty_like(res)
        nb = min(nq, 1 << 26)
        msb = env.timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qwords.data_ptr(), qflags.data_ptr(), nb, ref.data_ptr(), cb.CC_ALGO_BSEARCH, stream)), steps=2, warm=1)
        out["packed_bsearch_lookups_per_s"] = nb / (msb / 1000.0)
        out["parity"]["line_index_equals_bsearch_on"] = nb if bool(torch.equal(res[:nb], ref[:nb])) else 0
        del ref
        ms = env.timeit(lambda: N.check(L.cc_find_ascii_dev(g._h, qascii.data_ptr(), n_ascii, res.data_ptr(), cb.CC_ALGO_AUTO, stream)))
        out["ascii_lookups_per_s"] = n_ascii / (ms / 1000.0)
        out["ascii_queries_per_step"] = n_ascii
        pw = torch.empty((n_ascii, S_WORDS), dtype=torch.int64, device=dev)
        pf = torch.empty(n_ascii, dtype=torch.uint8, device=dev)
        ms = env.timeit(lambda: N.check(L.cc_pack_kmers_dev(local_rank, qascii.data_ptr(), n_ascii, K, pw.data_ptr(), pf.data_ptr(), stream)))
        out["pack_rows_per_s"] = n_ascii / (ms / 1000.0)                 # K3 on independent k-byte rows
        out["pack_rows_algorithmic_gb_per_s"] = out["pack_rows_per_s"] * (K + 8 * S_WORDS + 1) / 1e9
        out["parity"]["pack_rows_equal_generator_words"] = bool(torch.equal(pw[pf < 2], qwords[:n_ascii][pf < 2]) and
                                                                torch.equal(pf >= 2, qflags[:n_ascii] != 0))
        del pw, pf
        nm = min(nq, 1 << 25)
        ms = env.timeit(lambda: N.check(L.cc_find_packed_dev(g._h, qwords.data_ptr(), qflags.data_ptr(), nm, res.data_ptr(), cb.CC_ALGO_MERGE, stream)), steps=2, warm=1)
        out["packed_merge_path_unsorted_batch_lookups_per_s"] = nm / (ms / 1000.0)      # CC_ALGO_MERGE: radix sort + merge-path + un-permute
        # the same mode on a batch that arrives in key order (the case a merge is made for), beside the line index on that batch
        cw = qwords[:nm][qflags[:nm] == 0]                         # packable queries only: the words of a flagged query are not a key
        ns = cw.shape[0]
        o1 = torch.sort(cw[:, S_WORDS - 1] ^ (-(1 << 63)), stable=True).indices       # unsigned order, least significant word first
        for wcol in range(S_WORDS - 2, -1, -1):
            o1 = o1[torch.sort(cw[:, wcol][o1] ^ (-(1 << 63)), stable=True).indices]
        sw, sf = cw[o1].contiguous(), torch.zeros(ns, dtype=torch.uint8, device=dev)
        del o1, cw
        rm, ra = torch.empty(ns, dtype=torch.int64, device=dev), torch.empty(ns, dtype=torch.int64, device=dev)
        ms = env.timeit(lambda: N.check(L.cc_find_packed_dev(g._h, sw.data_ptr(), sf.data_ptr(), ns, rm.data_ptr(), cb.CC_ALGO_MERGE, stream)), steps=3, warm=1)
        out["packed_merge_path_sorted_batch_lookups_per_s"] = ns / (ms / 1000.0)
        ms = env.timeit(lambda: N.check(L.cc_find_packed_dev(g._h, sw.data_ptr(), sf.data_ptr(), ns, ra.data_ptr(), cb.CC_ALGO_AUTO, stream)), steps=3, warm=1)
        out["packed_line_index_sorted_batch_lookups_per_s"] = ns / (ms / 1000.0)
        out["parity"]["merge_path_equals_line_index_on"] = ns if bool(torch.equal(rm, ra)) else 0
        del sw, sf, rm, ra
        # ---- e2e through the host-buffer entry point: pinned host ASCII in, host indices out
        ne = min(n_ascii, 1 << 25)
        h_in = torch.empty((ne, K), dtype=torch.uint8, pin_memory=True)
        h_in.copy_(qascii[:ne])
        h_out = torch.empty(ne, dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize()
        for _ in range(2):
            N.check(L.cc_find_ascii(g._h, h_in.data_ptr(), ne, h_out.data_ptr(), cb.CC_ALGO_AUTO))
        t0 = time.perf_counter()
        for _ in range(3):
            N.check(L.cc_find_ascii(g._h, h_in.data_ptr(), ne, h_out.data_ptr(), cb.CC_ALGO_AUTO))
        dt = (time.perf_counter() - t0) / 3
        N.check(L.cc_find_ascii_dev(g._h, qascii.data_ptr(), ne, res.data_ptr(), cb.CC_ALGO_AUTO, stream))
        torch.cuda.synchronize()
        out["e2e"] = {"value": ne / dt, "unit": "lookups/s", "h2d_bytes_per_step": ne * K, "d2h_bytes_per_step": ne * 8, "ms_per_step": dt * 1e3,
                      "api": "cc_find_ascii (pinned host ASCII k-mers -> device in chunks on two streams -> indices back to host)",
                      "equals_device_path": bool(torch.equal(h_out.to(dev), res[:ne]))}
        # ---- the legacy per-record API: latency of one findRecord call, and of a vertex + its 8 neighbours in one call
        lat = {}
        one = np.ascontiguousarray(h_in[:9].numpy())
        idx9 = np.empty(9, dtype=np.int64)
        raw9 = np.empty(9 * REC_BYTES, dtype=np.uint8)
        for nqs in (1, 9):
            for _ in range(200):
                N.check(L.cc_find_records(g._h, one.ctypes.data, nqs, idx9.ctypes.data, raw9.ctypes.data))
            t0 = time.perf_counter()
            for _ in range(3000):
                L.cc_find_records(g._h, one.ctypes.data, nqs, idx9.ctypes.data, raw9.ctypes.data)
            lat["cc_find_records_%d_us_per_call" % nqs] = (time.perf_counter() - t0) / 3000 * 1e6
        lat["matches_batch_path"] = bool(np.array_equal(idx9, h_out[:9].numpy()))
        lat["note"] = "host k-mer in, index + record bytes out, measured through ctypes (about 1 us of call overhead)"
        out["find_record_latency"] = lat
        # ---- the CPU port beside it: the oracle's findRecord on a bounded sample of the same queries
        if not args.no_cpu:
            from oracle import orc
            hdr = synth.header_bytes(K, COLORS)
            image = np.empty(len(hdr) + nt * REC_BYTES, dtype=np.uint8)
            image[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
            torch.from_numpy(image[len(hdr):]).copy_(body.reshape(-1))
            og = orc.Graph(image)
            threads = os.cpu_count() or 1
            sample = h_in[:2_000_000].numpy()
            r1, _ = cpu_find_rate(og, sample[:200_000], 1)
            rmt, _ = cpu_find_rate(og, sample, threads)
            out["cpu_baseline"] = {"value": rmt, "unit": "lookups/s", "cores": threads, "kind": "port", "value_single_thread": r1,
                                   "sample": "oracle restatement of CortexGraph.findRecord (canonicalise + 3-point binary search with "
                                             "per-probe record decode, no LRU) on %d of the bench's ASCII queries over %d threads, "
                                             "200000 on one thread" % (sample.shape[0], threads)}
            del image, og
        del h_in, h_out
        traffic, note = measured_traffic("find_packed_lines_kernel")
        achieved = out["lookups_per_s"] * bpl / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                           "traffic_bytes_per_lookup": traffic.get("dram_bytes_per_lookup") if traffic else None, "traffic_note": note,
                           "peak_source": peak_src, "kernel": "find_packed_lines_kernel", "algorithmic_bytes_per_lookup": bpl,
                           "random_line_ceiling_lines_per_s": 3.9e10,
                           "note": "bound by the random-access rate of HBM, not its byte rate: tools/micro/randline.cu measures 3.9e10 random "
                                   "lines/s for 32-, 64- and 128-byte lines alike (profiles/r2_randline_calibration.log)"}
    else:
        from corticall_b200.host.sharded import RoutedLookup
        # segments sized for the balanced case (+25 %) instead of the worst case nq: 6x less peer-mapped address space per
        # rank (the worst-case layout cost 40 % on the search leg of most ranks: TLB reach); overflow is checked per batch
        rl = RoutedLookup(g, splitters, rank, world, dev, cap=int(nq / world * 1.25) + 4096, k=K, max_batch=nq)
        ms = env.timeit(lambda: rl.find_packed(qwords, qflags, res))
        out["lookups_per_s"] = nq * world / (ms / 1000.0)
        out["ms_per_step"] = ms
        out["exchange"] = ("peer memory over NVLink, one kernel per leg: route (owner + P2P store into the owner's inbox) -> barrier -> "
                           "search (bucket-line index, results stay on the owner) -> barrier -> gather (P2P pull)")
        out["hit_fraction"] = float((res >= 0).sum().item()) / nq
        checked, bad = oracle_check_sharded(env, body, lo, K, COLORS, splitters, qwords, qflags, res)
        out["parity"] = {"oracle_checked_queries_real_multi_gpu_path": checked, "oracle_mismatches": bad}
        env.barrier()
        rl.find_packed(qwords, qflags, res, profile=True)
        out["phase_ms_rank0"] = rl.phase_ms
        allp = [None] * world
        env.dist.all_gather_object(allp, {k: round(v, 3) for k, v in rl.phase_ms.items()})
        out["phase_ms_all_ranks"] = allp
        # ---- e2e: pinned host packed queries in, host indices out, copies inside the timed region
        ne = min(nq, 1 << 26)
        h_w = torch.empty((ne, S_WORDS), dtype=torch.int64, pin_memory=True)
        h_f = torch.empty(ne, dtype=torch.uint8, pin_memory=True)
        h_o = torch.empty(ne, dtype=torch.int64, pin_memory=True)
        h_w.copy_(qwords[:ne]); h_f.copy_(qflags[:ne])
        dw, df, do = torch.empty_like(qwords[:ne]), torch.empty_like(qflags[:ne]), torch.empty_like(res[:ne])

        def e2e_step():
            dw.copy_(h_w, non_blocking=True); df.copy_(h_f, non_blocking=True)
            rl.find_packed(dw, df, do)
            h_o.copy_(do, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step()
        dt = env.max((time.perf_counter() - t0) / 3)
        out["e2e"] = {"value": ne * world / dt, "unit": "lookups/s", "h2d_bytes_per_step": ne * (8 * S_WORDS + 1), "d2h_bytes_per_step": ne * 8,
                      "ms_per_step": dt * 1e3, "equals_device_path": bool(torch.equal(h_o.to(dev), res[:ne])),
                      "api": "pinned host packed queries -> device -> RoutedLookup.find_packed (cc_route_queries_dev / cc_find_routed_dev / "
                             "cc_gather_routed_dev) -> host indices"}
        del h_w, h_f, h_o, dw, df, do, rl
        achieved = out["lookups_per_s"] * bpl / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                           "traffic": None, "peak_source": peak_src + " x %d GPUs" % world, "kernel": "route_kernel + find_routed_kernel + gather_routed_kernel",
                           "algorithmic_bytes_per_lookup": bpl}
        if not args.no_compare:
            # replicas: when the whole table fits every GPU (it does for configs[2]), each rank can hold a full copy and look up its
            # own queries with no exchange at all (SURVEY 8e "alternative mode")
            try:
                words_all = synth.random_canonical_keys(SEED_LOOKUP, nt, K, dev)
                covf, edgf = synth.coverage_and_edges(SEED_LOOKUP, nt, COLORS, dev)
                full_body = synth.assemble_records(words_all, covf, edgf)
                del covf, edgf, words_all
                gfull = cb.CortexGraph.fromDevice(full_body.data_ptr(), K, COLORS, nt, firstIndex=0, device=local_rank, keepalive=full_body)
                gfull.buildIndex()
                res4 = torch.empty_like(res)
                msr = env.timeit(lambda: N.check(L.cc_find_packed_dev(gfull._h, qwords.data_ptr(), qflags.data_ptr(), nq, res4.data_ptr(), cb.CC_ALGO_AUTO, stream)))
                out["replicated_table_lookups_per_s"] = nq * world / (msr / 1000.0)
                out["replicated_agrees"] = bool(torch.equal(res, res4))
                # what cc_open_sharded_placed(CC_PLACE_AUTO) picks for this table: replicas when a lookup-ready copy (records + key column +
                # <= 32 B of bucket lines per record) takes at most a quarter of a device's memory, k-mer ranges otherwise
                need = nt * (REC_BYTES + 8 * S_WORDS + 32) + (64 << 20)
                fits = need <= torch.cuda.get_device_properties(dev).total_memory // 4
                out["placement"] = {"range_sharded_lookups_per_s": out["lookups_per_s"], "replicated_lookups_per_s": out["replicated_table_lookups_per_s"],
                                    "auto_picks": "replicate" if fits else "range",
                                    "auto_lookups_per_s": out["replicated_table_lookups_per_s"] if fits else out["lookups_per_s"],
                                    "lookup_ready_copy_gb": need / 1e9,
                                    "range_sharded_fraction_of_replicated": out["lookups_per_s"] / out["replicated_table_lookups_per_s"],
                                    "note": "lookups_per_s above is the k-mer-range-sharded path the north star names (the only one possible for "
                                            "configs[3]); a graph that fits every GPU is replicated by default and needs no exchange "
                                            "(one process: profiles/r2_sharded_abi_n8.json)"}
                gfull.dispose()
                del full_body, res4
            except Exception as e:          # comparison mode only
                out["replicated_table_lookups_per_s"] = None
                out["replicated_error"] = str(e)[:200]
            # the NCCL all-to-all formulation of the same exchange, as the comparison and as a cross-check of the results
            sl = ShardedLookup(g, splitters, rank, world, dev)
            res2 = torch.empty_like(res)
            ms2 = env.timeit(lambda: sl.find_packed(qwords, qflags, res2), steps=3, warm=1)
            out["nccl_all_to_all_lookups_per_s"] = nq * world / (ms2 / 1000.0)
            out["paths_agree"] = bool(torch.equal(res, res2))
            del res2, sl
    g.dispose()
    return out


# ---------------------------------------------------------------------------------------------------- configs[4]: many colours
def bench_many_colour(args, env):
    """BASELINE configs[4]: k=63 (2-word k-mers), 21 colours (child + 20 parental / reference colours), 121-byte records --
    odd record size, nothing naturally aligned, > 2^32 bytes of body.  Scan child vs all 20, at the named size."""
    torch = env.torch
    import corticall_b200 as cb
    from corticall_b200 import _native as N
    from oracle import orc
    from tools import synth

    k, c, s = 63, 21, 2
    S, O = 8 * s + 5 * c, 8 * s + 5
    n = args.many_records
    dev, L = env.dev, N.lib()
    body, _ = synth.make_graph_body(SEED_MANY, n, k, c, device=dev)
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, device=env.local_rank, keepalive=body)
    parents = np.arange(1, c, dtype=np.int32)
    cap = max(1 << 20, n // 16)
    out = torch.empty(cap * O + 64, dtype=torch.uint8, device=dev)
    out_idx = torch.empty(cap, dtype=torch.int64, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = env.timeit(lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, c - 1, out.data_ptr(), None, cap, cnt.data_ptr(), stream)), steps=10, warm=3)
    N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, c - 1, out.data_ptr(), out_idx.data_ptr(), cap, cnt.data_ptr(), stream))
    torch.cuda.synchronize()
    novel = int(cnt[0].item())
    alg = n * S + novel * O
    peak, peak_src = peaks()
    res = {"workload": "configs[4]: many-colour novelty scan, %d records, k=63 (2-word k-mers), 21 colours (child + 20 parents), %d-byte records, %.1f GB"
                       % (n, S, n * S / 1e9),
           "records_per_s": n / (ms / 1000.0), "ms_per_step": ms, "novel_records": novel,
           "roofline": {"bound": "hbm", "achieved": alg / (ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms / 1000.0) / 1e9 / peak,
                        "frac_of_nominal_8TBs": alg / (ms / 1000.0) / 1e9 / 8000.0, "algorithmic_bytes_per_launch": alg, "peak_source": peak_src,
                        "kernel": "scan_novel_fast_kernel (unaligned records, parent list in shared memory)"}}
    # parity 1: the whole output against the independent torch formulation (all n records)
    want_idx, want_rec = torch_expected_novel(torch, body, s, c, 0, list(range(1, c)))
    parity = {"torch_formulation_bit_exact": bool(novel == want_idx.numel() and torch.equal(out[:novel * O].reshape(-1, O), want_rec)
                                                  and torch.equal(out_idx[:novel], want_idx))}
    del want_idx, want_rec
    # parity 2: the oracle on 64 random slices of 1e6 records and on every adversarial record
    rng = np.random.default_rng(SEED_MANY)
    sl = min(1_000_000, n)
    starts = sorted(int(x) for x in rng.integers(0, max(n - sl, 1), size=min(64, max(1, n // sl))))
    gi = out_idx[:novel].cpu().numpy()
    go = out[:novel * O].cpu().numpy().reshape(-1, O)
    bad = checked = 0
    threads = os.cpu_count() or 1
    for b0 in range(0, len(starts), threads):
        group = starts[b0:b0 + threads]
        host = [body[a:a + sl].cpu().numpy().reshape(-1) for a in group]
        got = [None] * len(group)

        def work(j):
            got[j] = orc.find_rois_body(host[j], sl, k, s, c, 0, list(range(1, c)))

        th = [threading.Thread(target=work, args=(j,)) for j in range(len(group))]
        [x.start() for x in th]
        [x.join() for x in th]
        for a, (ocnt, orec, oidx) in zip(group, got):
            lo_i, hi_i = np.searchsorted(gi, a), np.searchsorted(gi, a + sl)
            ok = (hi_i - lo_i) == ocnt and np.array_equal(gi[lo_i:hi_i], oidx.astype(np.int64) + a) and np.array_equal(go[lo_i:hi_i].reshape(-1), orec)
            bad += 0 if ok else 1
            checked += sl
    adv = np.arange(synth.ADV_PERIOD_DEFAULT // 2, n, synth.ADV_PERIOD_DEFAULT)      # the records with coverages >= 2^31
    adv_rec = body[torch.from_numpy(adv).to(dev)].cpu().numpy()
    in_gpu = np.isin(adv, gi)
    adv_bad = 0
    for i in range(len(adv)):
        ocnt, _, _ = orc.find_rois_body(np.ascontiguousarray(adv_rec[i]), 1, k, s, c, 0, list(range(1, c)))
        adv_bad += int(bool(ocnt) != bool(in_gpu[i]))
    parity.update({"oracle_slices": len(starts), "oracle_records_checked": checked, "oracle_slice_mismatches": bad,
                   "adversarial_records_checked": int(len(adv)), "adversarial_mismatches": adv_bad})
    res["parity"] = parity
    g.dispose()
    return res


# ---------------------------------------------------------------------------------------------------- configs[3]: human scale
def bench_human_scale(args, env):
    """BASELINE configs[3]: a human-scale graph (3.75e8 records of k=31, 4 colours per GPU = 3e9 records / 84 GB on 8 GPUs),
    sharded by k-mer range.  Every rank generates its own key range of one global sorted graph on the device
    (tools/synth.range_canonical_keys), scans it, and serves routed lookups from every other rank."""
    torch = env.torch
    import corticall_b200 as cb
    from corticall_b200 import _native as N
    from corticall_b200.host.sharded import RoutedLookup
    from oracle import orc
    from tools import synth

    k, c, s = 31, 4, 1
    S, O = 8 * s + 5 * c, 8 * s + 5
    rank, world, dev, L = env.rank, env.world, env.dev, N.lib()
    n = args.human_records
    bounds = synth.canonical_range_bounds(k, world)
    t0 = time.perf_counter()
    keys = synth.range_canonical_keys(SEED_HUMAN + rank, n, k, bounds[rank], bounds[rank + 1], dev)
    body = torch.empty((n, S), dtype=torch.uint8, device=dev)
    chunk = 1 << 25
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        cov, edges = synth.coverage_and_edges(SEED_HUMAN, hi - lo, c, dev, offset=rank * n + lo)
        body[lo:hi] = synth.assemble_records([keys[0][lo:hi]], cov, edges)
        del cov, edges
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    first = rank * n
    g = cb.CortexGraph.fromDevice(body.data_ptr(), k, c, n, firstIndex=first, device=env.local_rank, keepalive=body)
    peak, peak_src = peaks()
    res = {"workload": "configs[3]: human-scale graph, %d records per GPU x %d GPUs = %.3g records (%.1f GB), k=31, 4 colours, "
                       "k-mer-range shards generated in place" % (n, world, n * world, n * world * S / 1e9),
           "generate_seconds": gen_s}
    # ---- scan
    parents = np.array([1, 2, 3], dtype=np.int32)
    cap = max(1 << 20, n // 16)
    out = torch.empty(cap * O + 64, dtype=torch.uint8, device=dev)
    out_idx = torch.empty(cap, dtype=torch.int64, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = env.timeit(lambda: N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), None, cap, cnt.data_ptr(), stream)), steps=10, warm=3)
    N.check(L.cc_find_novel_dev(g._h, 0, parents.ctypes.data, 3, out.data_ptr(), out_idx.data_ptr(), cap, cnt.data_ptr(), stream))
    torch.cuda.synchronize()
    novel = int(cnt[0].item())
    alg = n * S + novel * O
    res["scan"] = {"records_per_s": n * world / (ms / 1000.0), "ms_per_step": ms, "novel_records_this_rank": novel,
                   "roofline": {"bound": "hbm", "achieved": alg / (ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s (per GPU)",
                                "frac": alg / (ms / 1000.0) / 1e9 / peak, "frac_of_nominal_8TBs": alg / (ms / 1000.0) / 1e9 / 8000.0,
                                "peak_source": peak_src}}
    # globally ordered output: exclusive offsets from the all-gathered counts; order-dependent hash of (position, bytes)
    counts = torch.tensor([novel], dtype=torch.int64, device=dev)
    allc = [torch.empty_like(counts) for _ in range(world)]
    if world > 1:
        env.dist.all_gather(allc, counts)
    else:
        allc = [counts]
    offset = int(sum(int(x.item()) for x in allc[:rank]))
    total_novel = int(sum(int(x.item()) for x in allc))
    want_idx, want_rec = torch_expected_novel(torch, body, s, c, 0, [1, 2, 3])
    exact = bool(novel == want_idx.numel() and torch.equal(out[:novel * O].reshape(-1, O), want_rec) and torch.equal(out_idx[:novel], want_idx + first))

    def order_hash(rec, base):      # sum over records of mix(position) * mix(bytes): depends on the global position of every record
        m = rec.shape[0]
        pos = torch.arange(base, base + m, dtype=torch.int64, device=dev)
        h = synth.mix64(pos)
        acc = torch.zeros(m, dtype=torch.int64, device=dev)
        padded = torch.zeros((m, 16), dtype=torch.uint8, device=dev)
        padded[:, :O] = rec
        w = padded.view(torch.int64).reshape(m, 2)
        acc = synth.mix64(w[:, 0] ^ h) + synth.mix64(w[:, 1] + h)
        return int(acc.sum().item()) & ((1 << 64) - 1), int((acc * (pos | 1)).sum().item()) & ((1 << 64) - 1)

    h_got = order_hash(out[:novel * O].reshape(-1, O), offset)
    h_want = order_hash(want_rec, offset)
    res["scan"]["parity"] = {"torch_formulation_bit_exact_all_records": exact, "global_offset_of_this_rank": offset, "total_novel": total_novel,
                             "order_hash_128_matches": h_got == h_want,
                             "all_ranks_bit_exact": bool(env.sum(0.0 if (exact and h_got == h_want) else 1.0) == 0.0)}
    # oracle on random slices of 1e6 records + the adversarial records of this shard
    rng = np.random.default_rng(SEED_HUMAN + rank)
    sl = min(1_000_000, n)
    nsl = max(1, min(64 // world + 1, n // sl))
    starts = sorted(int(x) for x in rng.integers(0, max(n - sl, 1), size=nsl))
    gi = (out_idx[:novel] - first).cpu().numpy()
    go = out[:novel * O].cpu().numpy().reshape(-1, O)
    bad = 0
    for a in starts:
        ocnt, orec, oidx = orc.find_rois_body(body[a:a + sl].cpu().numpy().reshape(-1), sl, k, s, c, 0, [1, 2, 3])
        lo_i, hi_i = np.searchsorted(gi, a), np.searchsorted(gi, a + sl)
        ok = (hi_i - lo_i) == ocnt and np.array_equal(gi[lo_i:hi_i], oidx.astype(np.int64) + a) and np.array_equal(go[lo_i:hi_i].reshape(-1), orec)
        bad += 0 if ok else 1
    res["scan"]["parity"].update({"oracle_slices_all_ranks": int(env.sum(float(len(starts)))), "oracle_records_checked_all_ranks": int(env.sum(float(len(starts) * sl))),
                                  "oracle_slice_mismatches_all_ranks": int(env.sum(float(bad)))})
    del want_idx, want_rec, out, out_idx
    # ---- routed lookups: hits are keys drawn from EVERY shard (exchanged once at set-up), misses fresh draws
    g.buildIndex()
    nq = args.human_queries
    per = nq // (2 * world) + 1
    pick = synth.umod(synth.hash_idx(SEED_HUMAN, 40 + rank, torch.arange(per * world, dtype=torch.int64, device=dev)), n)
    send = keys[0][pick].contiguous()
    pool = torch.empty_like(send)
    if world > 1:
        env.dist.all_to_all_single(pool, send)
    else:
        pool = send
    del send, pick
    qwords = torch.empty((nq, 1), dtype=torch.int64, device=dev)
    qflags = torch.empty(nq, dtype=torch.uint8, device=dev)
    for o in range(0, nq, 1 << 24):
        m = min(1 << 24, nq - o)
        _, canon, valid = synth.make_queries(SEED_HUMAN, [pool], k, m, offset=rank * nq + o)
        qwords[o:o + m, 0] = canon[0]
        qflags[o:o + m] = torch.where(valid, 0, 2).to(torch.uint8)
        del canon, valid
    mine = keys[0][:1].clone()
    parts = [torch.empty_like(mine) for _ in range(world)]
    if world > 1:
        env.dist.all_gather(parts, mine)
        splitters = torch.stack(parts[1:]).contiguous()
    else:
        splitters = None
    del pool
    rl = RoutedLookup(g, splitters, rank, world, dev, cap=int(nq / world * 1.25) + 4096, k=k, max_batch=nq)
    resq = torch.empty(nq, dtype=torch.int64, device=dev)
    msl = env.timeit(lambda: rl.find_packed(qwords, qflags, resq))
    checked, badq = oracle_check_sharded(env, body, first, k, c, splitters, qwords, qflags, resq)
    env.barrier()
    rl.find_packed(qwords, qflags, resq, profile=True)
    bpl = 8 * s + 8 + n * world * 8 * s / (nq * world)
    res["lookup"] = {"lookups_per_s": nq * world / (msl / 1000.0), "ms_per_step": msl, "queries_per_step": nq * world,
                     "hit_fraction": float((resq >= 0).sum().item()) / nq, "phase_ms_rank0": rl.phase_ms,
                     "index_bytes_per_gpu": None, "algorithmic_bytes_per_lookup": bpl,
                     "roofline": {"bound": "hbm", "achieved": nq * world / (msl / 1000.0) * bpl / 1e9, "peak": peak * world, "unit": "GB/s",
                                  "frac": nq * world / (msl / 1000.0) * bpl / 1e9 / (peak * world)},
                     "parity": {"oracle_checked_queries_real_multi_gpu_path": checked, "oracle_mismatches": badq}}
    del rl
    g.dispose()
    return res


# ---------------------------------------------------------------------------------------------------- composite e2e
def bench_composite(args, env):
    """What a user of the reference runs, end to end, with the graph resident across the calls: open a .ctx file ->
    FindROIs (write the ROI graph) -> look up every k-mer of a set of contigs (Call.loadChildWalk) -- host buffers in,
    host buffers out, file I/O on /dev/shm inside the timed region -- against the same sequence on the CPU port."""
    torch = env.torch
    import corticall_b200 as cb
    from oracle import orc
    from tools import synth

    n = args.records
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path, roi_path = os.path.join(tmp, "corticall_bench_%d.ctx" % os.getpid()), os.path.join(tmp, "corticall_bench_%d.rois.ctx" % os.getpid())
    body, table = synth.make_graph_body(SEED_SCAN, n, K, COLORS, device=env.dev)
    hdr = synth.header_bytes(K, COLORS)
    image = np.empty(len(hdr) + n * REC_BYTES, dtype=np.uint8)
    image[:len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
    torch.from_numpy(image[len(hdr):]).copy_(body.reshape(-1))
    with open(path, "wb") as f:
        f.write(image.data)
    # contigs: 8 x 4 Mb stitched from k-mers (half of them from the graph, either strand): about 1 % of the windows hit
    ncontig, clen = 8, 4_000_000
    contigs = []
    for i in range(ncontig):
        a, _, _ = synth.make_queries(SEED_SCAN + i, table, K, clen // K, corrupt_permille=0)
        contigs.append(a.reshape(-1).cpu().numpy().copy())
    del body, table
    torch.cuda.empty_cache()
    try:
        def gpu_run():
            t = [time.perf_counter()]
            g = cb.CortexGraph(path, device=env.local_rank)
            t.append(time.perf_counter())
            cnt = g.writeRois(0, [1, 2, 3], roi_path)
            t.append(time.perf_counter())
            hits = 0
            for sq in contigs:
                hits += int((g.findWindows(sq) >= 0).sum())
            t.append(time.perf_counter())
            g.dispose()
            t.append(time.perf_counter())
            return cnt, hits, [b - a for a, b in zip(t[:-1], t[1:])]

        gpu_run()
        runs = []
        for _ in range(3):          # wall-clock with file I/O inside: report the median of three and list all of them
            cnt, hits, phases = gpu_run()
            runs.append((sum(phases), phases))
        runs.sort(key=lambda r: r[0])
        gpu_s, phases = runs[1]
        with open(roi_path, "rb") as f:
            roi_bytes = f.read()
        # the CPU port: same steps; the lookups on a bounded sample of the windows, scaled
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        og = orc.Graph(image)
        dt_scan, ocnt, parts = cpu_scan_pass(image[len(hdr):], n, K, S_WORDS, COLORS, [1, 2, 3], threads, True, keep_output=True)
        cpu_scan_s = time.perf_counter() - t0
        o_rec = np.concatenate([p[1] for p in parts])
        sample = 400_000
        win = np.lib.stride_tricks.sliding_window_view(contigs[0], K)[:sample]
        t0 = time.perf_counter()
        rate, want = cpu_find_rate(og, np.ascontiguousarray(win), threads)
        windows = sum(len(sq) - K + 1 for sq in contigs)
        cpu_find_s = windows / rate
        g = cb.CortexGraph(path, device=env.local_rank)
        got = g.findWindows(contigs[0][:sample + K - 1])
        g.dispose()
        return {"workload": "open %d-record .ctx (%.2f GB, /dev/shm) -> FindROIs to a file -> findRecord for all %d windows of %d contigs; graph resident across calls"
                            % (n, len(image) / 1e9, windows, ncontig),
                "gpu_seconds": gpu_s, "gpu_seconds_runs": [r[0] for r in runs],
                "gpu_phase_seconds": dict(zip(("open", "find_rois_to_file", "find_windows", "dispose"), phases)), "cpu_port_seconds": cpu_scan_s + cpu_find_s, "cpu_port_scan_seconds": cpu_scan_s,
                "cpu_port_lookup_seconds_extrapolated": cpu_find_s, "cpu_port_threads": threads, "speedup": (cpu_scan_s + cpu_find_s) / gpu_s,
                "novel_records": cnt, "window_hits": hits,
                "parity": {"roi_file_equals_oracle": bool(cnt == ocnt and roi_bytes == orc.roi_header(K, S_WORDS, "child") + o_rec.tobytes()),
                           "lookup_sample_equals_oracle": bool(np.array_equal(got, want)), "lookup_sample": sample}}
    finally:
        for p in (path, roi_path):
            try:
                os.unlink(p)
            except OSError:
                pass


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
