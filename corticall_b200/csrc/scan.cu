// scan.cu -- K1 (streaming .ctx record decode) and K1+K2 fused (the FindROIs novelty scan).
//
// Reference semantics reproduced (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
//   record layout / decode   S/utils/io/graph/cortex/CortexGraph.java:189-237, docs/ctx_spec.md table 5
//   novelty predicate        S/commands/discover/roi/FindROIs.java:72-82   (signed int compares)
//   output record layout     S/commands/discover/roi/FindROIs.java:54-59 + CortexGraphWriter.java:106-138
//   output order = input order (FindROIs.java:52-64)
//
// B200 design.  The record array is an unaligned array-of-structures (S = 8s+5c bytes, S odd whenever c
// is odd).  A persistent grid (ctas_per_sm x #SM CTAs) streams it through shared memory with 1-D bulk
// async copies (cp.async.bulk -> SASS UBLKCP, the TMA engine) into a ring of `stages` tiles guarded by
// full/empty mbarriers: one producer warp issues copies of 16-byte-aligned supersets of each tile, eight
// consumer warps decode records straight out of shared memory with funnel-shifted word reads (no field
// is naturally aligned).  Tiles are handed out by a global ticket so a tile's predecessors are always
// already running; novel records are compacted in input order with warp ballots, an intra-tile scan and
// a decoupled look-back over per-tile descriptors (single pass: the array is read exactly once and only
// novel records are written).  Algorithmic bytes per record: S + f*(8s+5), f = novel fraction.
#include <string.h>

#include <algorithm>

#include "cc_internal.hpp"
#include "device_utils.cuh"

namespace cc {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kScanWarp = kConsumerWarps;            // warp 8: intra-tile scan + decoupled look-back
constexpr int kProducerWarp = kConsumerWarps + 1;    // warp 9: tile tickets + bulk copies
constexpr int kThreads = kConsumerThreads + 64;
constexpr int kMaxIters = 8;                        // tile_records <= kMaxIters * kConsumerThreads
constexpr int kMaxStages = 8;
constexpr int kSlots = kMaxIters * kConsumerWarps;  // (iteration, warp) count slots per tile
constexpr uint32_t kCtrlBytes = 2048;
constexpr int kFastParents = 4;                     // parent offsets kept in registers up to this many

constexpr uint64_t kFlagAgg = 1ull << 62;
constexpr uint64_t kFlagPrefix = 2ull << 62;
constexpr uint64_t kFlagMask = 3ull << 62;
constexpr uint64_t kValueMask = (1ull << 40) - 1;
constexpr uint32_t kEpochMask = (1u << 22) - 1;

struct TileGeom {
    uint32_t S;             // record bytes
    uint32_t tile_records;  // multiple of 32
    uint32_t num_tiles;
    uint32_t stages;
    uint32_t stage_bytes;   // multiple of 128, >= tile_records*S + 32
    uint32_t iters;         // ceil(tile_records / kConsumerThreads)
};

struct ScanKParams {
    const uint8_t *body;
    uint64_t n, index_base;
    uint32_t s, c, cov_off, edge_off, O;
    int32_t child, nparents;
    uint32_t child_off;                 // cov_off + 4*child
    uint32_t parent_off[kFastParents];  // cov_off + 4*parent (first kFastParents parents)
    const int32_t *parents;
    uint8_t *out;
    uint64_t *out_index;
    uint64_t cap;
    const uint64_t *total_in;
    uint64_t *total_out;
    uint64_t *tile_state;
    uint32_t *tile_counter;
    uint32_t ticket_base;
    uint32_t epoch;
    uint32_t debug;
    int *err;
    // general kernel only: run iff *run_if == run_expect (run_if == null: always).  When skipped, block 0 still
    // draws the launch's tickets so the host's mirror of the ticket counter stays exact.
    const uint32_t *run_if;
    uint32_t run_expect, skip_tickets;
    // fast kernel only
    uint32_t chunk_log2, num_chunks, stg_stride, stg_cap, Ow;
    uint8_t *stg_scratch;               // global (L2-resident) staging lists: [grid][2][kConsumerWarps][stg_cap] entries
    uint32_t *overflow;
    // chunks whose staging list overflowed in the fast kernel: {chunk index, exclusive output prefix} pairs, rewritten
    // by redo_chunks_kernel.  dirty_ctl[0] = entries, [1] = ticket, [2] = CTAs done (reset by the last CTA to leave).
    uint64_t *dirty_list;
    uint32_t *dirty_ctl;
    TileGeom g;
};

struct DecodeKParams {
    const uint8_t *body;
    uint64_t n;
    uint32_t s, c, cov_off, edge_off;
    uint64_t *words;
    int32_t *cov;
    uint8_t *edges;
    uint32_t *tile_counter;
    uint32_t ticket_base;
    int *err;
    TileGeom g;
};

// Shared-memory control block (first kCtrlBytes of dynamic smem).
struct Ctrl {
    uint64_t full[kMaxStages];      // producer -> everyone: tile bytes have landed (tx-count barrier)
    uint64_t empty[kMaxStages];     // consumer warps -> producer: stage may be overwritten
    uint64_t counted[2];            // consumer warps -> scan warp: per-warp novel counts of the tile are in cnt[slot]
    uint64_t based[2];              // scan warp -> consumer warps: pre[slot] and tile_base[slot] are valid
    int32_t tile_id[kMaxStages];
    uint32_t cnt[2][kSlots];
    uint32_t pre[2][kSlots];
    uint64_t tile_base[2];
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block too large");

// ------------------------------------------------------------------ producer: the TMA side of the ring
// One thread.  Tiles are handed out by a global ticket (so every predecessor of a tile is already owned by a
// running CTA: the look-back can never wait on an unscheduled CTA); the ticket of the NEXT tile is requested
// before waiting for a free stage so the atomic's round trip overlaps the wait and the copy issue.
__device__ __forceinline__ void producer_loop(uint64_t *full, uint64_t *empty, int32_t *tile_id, uint8_t *stage0,
                                              const uint8_t *body, uint64_t n, const TileGeom &g, uint32_t *tile_counter,
                                              uint32_t ticket_base, int *err) {
    const uint64_t policy = make_evict_first_policy();
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(body) & 15u);
    const uint8_t *abase = body - mis;
    uint32_t t_next = atomicAdd(tile_counter, 1u) - ticket_base;
    for (uint32_t it = 0;; ++it) {
        const uint32_t st = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1u;
        const uint32_t t = t_next;
        if (t < g.num_tiles) t_next = atomicAdd(tile_counter, 1u) - ticket_base;   // exactly one ticket >= num_tiles per CTA
        mbar_wait(&empty[st], ph ^ 1u, err, DEV_TIMEOUT_EMPTY);
        if (t >= g.num_tiles) {
            tile_id[st] = -1;
            mbar_arrive(&full[st]);                 // sentinel: consumers see tile_id < 0 and leave
            break;
        }
        tile_id[st] = (int32_t)t;
        const uint64_t rec0 = (uint64_t)t * g.tile_records;
        const uint64_t left = n - rec0;
        const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
        const uint32_t bytes = (nrec * g.S + mis + 15u) & ~15u;
        mbar_arrive_expect_tx(&full[st], bytes);
        bulk_g2s(stage0 + (size_t)st * g.stage_bytes, abase + rec0 * g.S, bytes, &full[st], policy);
    }
}

// ------------------------------------------------------------------ decoupled look-back (one warp)
__device__ __forceinline__ uint64_t pack_desc(uint64_t flag, uint32_t epoch, uint64_t value) {
    return flag | ((uint64_t)(epoch & kEpochMask) << 40) | (value & kValueMask);
}

// Publishes this tile's aggregate, returns its exclusive prefix (novel records in all earlier tiles + base) and
// publishes its inclusive prefix.  Called by all 32 lanes of one warp.
__device__ __forceinline__ uint64_t lookback(uint64_t *state, uint32_t tile, uint64_t agg, uint32_t epoch,
                                             const uint64_t *total_in, int *err) {
    const uint32_t lane = threadIdx.x & 31u;
    if (tile == 0) {
        uint64_t base = total_in ? *total_in : 0ull;
        if (lane == 0) st_relaxed_gpu(&state[0], pack_desc(kFlagPrefix, epoch, base + agg));
        return base;
    }
    if (lane == 0) st_relaxed_gpu(&state[tile], pack_desc(kFlagAgg, epoch, agg));
    uint64_t excl = 0;
    int64_t idx = (int64_t)tile - 1;
    const uint64_t want_epoch = (uint64_t)(epoch & kEpochMask);
    uint64_t t0 = 0;
    while (true) {
        const int64_t my = idx - (int64_t)lane;
        uint64_t d = kFlagAgg;                       // virtual tiles below 0 contribute 0
        if (my >= 0) {
            uint32_t spins = 0;
            while (true) {
                d = ld_relaxed_gpu(&state[my]);
                if ((d & kFlagMask) != 0 && ((d >> 40) & kEpochMask) == want_epoch) break;
                if ((++spins & 0x3ff) == 0) {
                    uint64_t now = globaltimer_ns();
                    if (t0 == 0) t0 = now;
                    if (now - t0 > kWatchdogNs) watchdog_fail(err, DEV_TIMEOUT_LOOKBACK);
                }
            }
        }
        const uint32_t is_prefix = __ballot_sync(0xffffffffu, (d & kFlagMask) == kFlagPrefix);
        uint64_t v = d & kValueMask;
        if (my < 0) v = 0;
        if (is_prefix) {
            const uint32_t first = __ffs(is_prefix) - 1;       // nearest predecessor holding a prefix
            if (lane > first) v = 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            break;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        idx -= 32;
    }
    if (lane == 0) st_relaxed_gpu(&state[tile], pack_desc(kFlagPrefix, epoch, excl + agg));
    return excl;
}

// Two-phase form of the look-back for a warp that serves several chunks at once.
//   lookback_publish  announces the aggregate (chunk 0 announces its inclusive prefix straight away)
//   lookback_try      one polling pass over the predecessors; false if any descriptor in reach is not there yet
__device__ __forceinline__ void lookback_publish(uint64_t *state, uint32_t unit, uint64_t agg, uint32_t epoch, const uint64_t *total_in) {
    if ((threadIdx.x & 31u) != 0) return;
    if (unit == 0) st_relaxed_gpu(&state[0], pack_desc(kFlagPrefix, epoch, (total_in ? *total_in : 0ull) + agg));
    else st_relaxed_gpu(&state[unit], pack_desc(kFlagAgg, epoch, agg));
}
__device__ __forceinline__ bool lookback_try(uint64_t *state, uint32_t unit, uint64_t agg, uint32_t epoch, const uint64_t *total_in,
                                             uint64_t &excl_out) {
    const uint32_t lane = threadIdx.x & 31u;
    if (unit == 0) { excl_out = total_in ? *total_in : 0ull; return true; }
    const uint64_t want_epoch = (uint64_t)(epoch & kEpochMask);
    uint64_t excl = 0;
    int64_t idx = (int64_t)unit - 1;
    while (true) {
        const int64_t my = idx - (int64_t)lane;
        uint64_t d = kFlagAgg;
        bool ready = true;
        if (my >= 0) {
            d = ld_relaxed_gpu(&state[my]);
            ready = (d & kFlagMask) != 0 && ((d >> 40) & kEpochMask) == want_epoch;
        }
        const uint32_t is_prefix = __ballot_sync(0xffffffffu, ready && (d & kFlagMask) == kFlagPrefix);
        const uint32_t not_ready = __ballot_sync(0xffffffffu, !ready);
        uint64_t v = (my >= 0 && ready) ? (d & kValueMask) : 0ull;
        if (is_prefix) {
            const uint32_t first = __ffs(is_prefix) - 1;       // nearest predecessor holding a prefix
            if (not_ready & ((2u << first) - 1u)) return false;   // someone nearer than that prefix has not published yet
            if (lane > first) v = 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            break;
        }
        if (not_ready) return false;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        idx -= 32;
    }
    if (lane == 0) st_relaxed_gpu(&state[unit], pack_desc(kFlagPrefix, epoch, excl + agg));
    excl_out = excl;
    return true;
}

// ------------------------------------------------------------------ K1+K2: novelty scan
// Warp roles: 0..7 consumers, 8 scan, 9 producer.  Per tile:
//   consumers  wait full[stage]; evaluate FindROIs.isNovel for their records (ballot masks stay in registers);
//              post per-(iteration,warp) counts; arrive counted[slot];
//   scan warp  wait counted[slot]; exclusive scan of the counts; PUBLISH THE TILE AGGREGATE at once; look back
//              for the exclusive prefix; arrive based[slot];
//   consumers  (after having evaluated the NEXT tile) wait based[slot]; copy their novel records to
//              out[(base + rank) * O]; release the stage.
// The aggregate of a tile is therefore published as soon as its bytes have been seen, independent of how long
// earlier tiles of the same CTA wait for their own prefixes -- that keeps the look-back chain short.
template <bool ALIGNED4, int NP>   // NP = number of parents when <= kFastParents, else -1 (list in shared memory)
__global__ void __launch_bounds__(kThreads, 2) scan_novel_kernel(const __grid_constant__ ScanKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (p.run_if != nullptr && *reinterpret_cast<const volatile uint32_t *>(p.run_if) != p.run_expect) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.tile_counter, p.skip_tickets);
        return;
    }
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem);
    uint32_t *parents_s = reinterpret_cast<uint32_t *>(smem + kCtrlBytes);     // byte offsets cov_off + 4*parent
    const uint32_t parents_bytes = NP >= 0 ? 0u : (((uint32_t)p.nparents * 4u + 127u) & ~127u);
    uint8_t *stage0 = smem + kCtrlBytes + parents_bytes;

    const TileGeom &g = p.g;
    if (NP < 0)
        for (int i = threadIdx.x; i < p.nparents; i += kThreads) parents_s[i] = p.cov_off + 4u * (uint32_t)p.parents[i];
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < g.stages; ++i) {
            mbar_init(&ctrl->full[i], 1);
            mbar_init(&ctrl->empty[i], kConsumerWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&ctrl->counted[i], kConsumerWarps);
            mbar_init(&ctrl->based[i], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (warp == kProducerWarp) {
        if (lane == 0)
            producer_loop(ctrl->full, ctrl->empty, ctrl->tile_id, stage0, p.body, p.n, g, p.tile_counter, p.ticket_base, p.err);
        return;
    }
    const uint32_t nslots = g.iters * kConsumerWarps;

    if (warp == kScanWarp) {
        for (uint32_t it = 0;; ++it) {
            const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u, slot = it & 1u, sph = (it >> 1) & 1u;
            mbar_wait(&ctrl->full[st], ph, p.err, DEV_TIMEOUT_FULL);
            const int32_t tile = ctrl->tile_id[st];
            if (tile < 0) break;
            mbar_wait(&ctrl->counted[slot], sph, p.err, DEV_TIMEOUT_FULL);
            const uint32_t a = lane < nslots ? ctrl->cnt[slot][lane] : 0u;
            const uint32_t b = lane + 32 < nslots ? ctrl->cnt[slot][lane + 32] : 0u;
            uint32_t ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o);
                const uint32_t tb = __shfl_up_sync(0xffffffffu, ib, o);
                if ((int)lane >= o) { ia += ta; ib += tb; }
            }
            const uint32_t tot_a = __shfl_sync(0xffffffffu, ia, 31);
            const uint32_t tot_b = __shfl_sync(0xffffffffu, ib, 31);
            if (lane < nslots) ctrl->pre[slot][lane] = ia - a;
            if (lane + 32 < nslots) ctrl->pre[slot][lane + 32] = tot_a + ib - b;
            const uint64_t agg = (uint64_t)tot_a + tot_b;
            const uint64_t excl = (p.debug & 1u) ? 0ull : lookback(p.tile_state, (uint32_t)tile, agg, p.epoch, p.total_in, p.err);
            if (lane == 0) {
                ctrl->tile_base[slot] = excl;
                if ((uint32_t)tile == g.num_tiles - 1) *p.total_out = excl + agg;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->based[slot]);
        }
        return;
    }

    // ---- consumer warps
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.body) & 15u);
    const uint32_t S = g.S;
    uint32_t poff[kFastParents > 0 ? kFastParents : 1];
#pragma unroll
    for (int i = 0; i < kFastParents; ++i) poff[i] = p.parent_off[i];
    const uint32_t lane_lt = (1u << lane) - 1u;

    uint32_t masks_prev[kMaxIters];
#pragma unroll
    for (int j = 0; j < kMaxIters; ++j) masks_prev[j] = 0;
    int32_t tile_prev = -1;
    uint32_t st_prev = 0;

    for (uint32_t it = 0;; ++it) {
        const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u, slot = it & 1u;
        mbar_wait(&ctrl->full[st], ph, p.err, DEV_TIMEOUT_FULL);
        const int32_t tile = ctrl->tile_id[st];
        uint32_t masks_cur[kMaxIters];
#pragma unroll
        for (int j = 0; j < kMaxIters; ++j) masks_cur[j] = 0;
        if (tile >= 0) {
            // ---- predicate (FindROIs.isNovel :72-82): coverage[child] > 0 (signed) && every listed parent == 0
            const uint64_t rec0 = (uint64_t)tile * g.tile_records;
            const uint64_t left = p.n - rec0;
            const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
            const uint8_t *tile_s = stage0 + (size_t)st * g.stage_bytes + mis;
#pragma unroll
            for (int j = 0; j < kMaxIters; ++j) {
                if (j < (int)g.iters) {
                    const uint32_t r = (uint32_t)j * kConsumerThreads + threadIdx.x;
                    bool novel = false;
                    if (r < nrec && !(p.debug & 2u)) {
                        const uint8_t *rec = tile_s + r * S;
                        uint32_t any_parent = 0;
                        if (NP >= 0) {
#pragma unroll
                            for (int i = 0; i < (NP > 0 ? NP : 0); ++i) any_parent |= lds_u32<ALIGNED4>(rec + poff[i]);
                        } else {
                            for (int i = 0; i < p.nparents; ++i) any_parent |= lds_u32<ALIGNED4>(rec + parents_s[i]);
                        }
                        const int32_t child_cov = (int32_t)lds_u32<ALIGNED4>(rec + p.child_off);
                        novel = (child_cov > 0) && (any_parent == 0);
                    }
                    masks_cur[j] = __ballot_sync(0xffffffffu, novel);
                    if (lane == 0) ctrl->cnt[slot][j * kConsumerWarps + warp] = __popc(masks_cur[j]);
                }
            }
            if (lane == 0) mbar_arrive(&ctrl->counted[slot]);       // release: orders the cnt stores of this lane
        }
        if (tile_prev >= 0) {
            // ---- ordered write-out of the PREVIOUS tile (its prefix has had a whole tile's time to resolve)
            const uint32_t pslot = slot ^ 1u, psph = ((it - 1) >> 1) & 1u;
            if (!(p.debug & 8u)) mbar_wait(&ctrl->based[pslot], psph, p.err, DEV_TIMEOUT_LOOKBACK);
            const uint64_t tile_base = ctrl->tile_base[pslot];
            const uint64_t rec0 = (uint64_t)tile_prev * g.tile_records;
            const uint8_t *tile_s = stage0 + (size_t)st_prev * g.stage_bytes + mis;
#pragma unroll
            for (int j = 0; j < kMaxIters; ++j) {
                const uint32_t m = masks_prev[j];
                if (j < (int)g.iters && m != 0) {
                    if ((m >> lane) & 1u) {
                        const uint64_t pos = tile_base + ctrl->pre[pslot][j * kConsumerWarps + warp] + __popc(m & lane_lt);
                        if (pos < p.cap) {
                            const uint32_t r = (uint32_t)j * kConsumerThreads + threadIdx.x;
                            const uint8_t *rec = tile_s + r * S;
                            uint8_t *dst = p.out + pos * p.O;
                            // CortexGraphWriter.addRecord :106-138: s words verbatim, child coverage (LE), child edge byte
                            for (uint32_t b = 0; b < p.cov_off; ++b) dst[b] = rec[b];
#pragma unroll
                            for (uint32_t b = 0; b < 4; ++b) dst[p.cov_off + b] = rec[p.child_off + b];
                            dst[p.cov_off + 4] = rec[p.edge_off + p.child];
                            if (p.out_index) p.out_index[pos] = p.index_base + rec0 + r;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->empty[st_prev]);
        }
        if (tile < 0) break;
#pragma unroll
        for (int j = 0; j < kMaxIters; ++j) masks_prev[j] = masks_cur[j];
        tile_prev = tile;
        st_prev = st;
    }
}

// ------------------------------------------------------------------ K1+K2, fast path: chunked scan, deferred look-back
// The per-tile look-back above costs every tile a global round trip with all consumer warps parked behind it.
// Novel k-mers are rare (well under 1 % of a trio graph), so this kernel defers everything that needs another
// CTA: a CTA claims CHUNKS of 2^chunk_log2 consecutive tiles; each consumer warp owns a contiguous slice of every
// tile and appends its novel records to a private shared-memory list, posting one count per (tile, warp) -- the
// eight consumer warps never synchronise with each other or with any other CTA while they stream.  At the end of
// a chunk the scan warp sums the counts, publishes the CHUNK aggregate, looks back over chunk descriptors for the
// exclusive prefix and copies the staged runs to out[] in (tile, warp) = input order, while the consumers are
// already filling the other staging buffer with the next chunk.  A chunk whose staging list overflows (dense
// novelty) raises *overflow = epoch; the general kernel is launched right behind this one and runs only then.
constexpr int kMaxChunkTiles = 16;

// Consumers may run up to kSegRing chunks ahead of the scan warp: the look-back of a chunk has to wait for the slowest of
// the concurrently running predecessor chunks, and with only two slots that wait stalled the streaming (measured: 10 %).
constexpr int kSegRing = 8;
constexpr uint32_t kFastCtrlBytes = 13312;
constexpr uint32_t kMaxListed = 2048;          // staged records per chunk the scan warp can list (8 warps x stg_cap)
struct FastCtrl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t seg_done[kSegRing];    // consumer warps -> scan warp: segment (chunk) fully evaluated and staged
    uint64_t seg_free[kSegRing];    // scan warp -> consumer warps: staging lists / counts of that slot are free
    int32_t tile_id[kMaxStages];
    int32_t seg_chunk[kSegRing];    // chunk index of the segment, -1 = no more work
    uint32_t ovfw[kSegRing][kConsumerWarps];
    uint32_t cnt[kSegRing][kMaxChunkTiles * kConsumerWarps];
    uint32_t list[kMaxListed];      // scan warp: the chunk's staged records in output order, (warp << 24) | index
};
static_assert(sizeof(FastCtrl) <= kFastCtrlBytes, "control block too large");

__device__ __forceinline__ void producer_loop_chunks(FastCtrl *ctrl, uint8_t *stage0, const ScanKParams &p) {
    const TileGeom &g = p.g;
    const uint64_t policy = make_evict_first_policy();
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.body) & 15u);
    const uint8_t *abase = p.body - mis;
    uint32_t c_next = atomicAdd(p.tile_counter, 1u) - p.ticket_base;
    uint32_t it = 0;
    while (true) {
        const uint32_t c = c_next;
        if (c >= p.num_chunks) {                     // exactly one ticket >= num_chunks per CTA
            const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u;
            mbar_wait(&ctrl->empty[st], ph ^ 1u, p.err, DEV_TIMEOUT_EMPTY);
            ctrl->tile_id[st] = -1;
            mbar_arrive(&ctrl->full[st]);
            break;
        }
        c_next = atomicAdd(p.tile_counter, 1u) - p.ticket_base;      // round trip overlaps the whole chunk
        const uint32_t t0 = c << p.chunk_log2;
        const uint32_t t1 = min(t0 + (1u << p.chunk_log2), g.num_tiles);
        for (uint32_t t = t0; t < t1; ++t, ++it) {
            const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u;
            mbar_wait(&ctrl->empty[st], ph ^ 1u, p.err, DEV_TIMEOUT_EMPTY);
            ctrl->tile_id[st] = (int32_t)t;
            const uint64_t rec0 = (uint64_t)t * g.tile_records;
            const uint64_t left = p.n - rec0;
            const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
            const uint32_t bytes = (nrec * g.S + mis + 15u) & ~15u;
            mbar_arrive_expect_tx(&ctrl->full[st], bytes);
            bulk_g2s(stage0 + (size_t)st * g.stage_bytes, abase + rec0 * g.S, bytes, &ctrl->full[st], policy);
        }
    }
}

template <bool ALIGNED4, int NP>
__global__ void __launch_bounds__(kThreads, 3) scan_novel_fast_kernel(const __grid_constant__ ScanKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    FastCtrl *ctrl = reinterpret_cast<FastCtrl *>(smem);
    uint32_t *parents_s = reinterpret_cast<uint32_t *>(smem + kFastCtrlBytes);
    const uint32_t parents_bytes = NP >= 0 ? 0u : (((uint32_t)p.nparents * 4u + 127u) & ~127u);
    // Staging lists live in global memory (they are written at the novelty rate, well under 1 % of the traffic, and stay
    // in L2), which leaves all of shared memory to the tile ring: more bytes in flight per SM, and room for long lists.
    const uint32_t stg_warp_bytes = p.stg_cap * p.stg_stride;
    uint8_t *stg0 = p.stg_scratch + (size_t)blockIdx.x * kSegRing * kConsumerWarps * stg_warp_bytes;   // [kSegRing][kConsumerWarps][stg_cap]
    uint8_t *stage0 = smem + kFastCtrlBytes + parents_bytes;

    const TileGeom &g = p.g;
    if (NP < 0)
        for (int i = threadIdx.x; i < p.nparents; i += kThreads) parents_s[i] = p.cov_off + 4u * (uint32_t)p.parents[i];
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < g.stages; ++i) {
            mbar_init(&ctrl->full[i], 1);
            mbar_init(&ctrl->empty[i], kConsumerWarps);
        }
        for (int i = 0; i < kSegRing; ++i) {
            mbar_init(&ctrl->seg_done[i], kConsumerWarps);
            mbar_init(&ctrl->seg_free[i], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (warp == kProducerWarp) {
        if (lane == 0) producer_loop_chunks(ctrl, stage0, p);
        return;
    }
    const uint32_t T = 1u << p.chunk_log2;

    if (warp == kScanWarp) {
        // Two cursors over the ring of chunk slots: `pub` announces the aggregate of every chunk the consumers have
        // finished AS SOON AS it is finished; `seg` resolves prefixes and copies out, in order.  Announcing must never
        // queue behind a look-back that is itself waiting for other CTAs' announcements -- that feedback serialised
        // the whole grid when both were done by one blocking loop.
        uint32_t pub = 0, seg = 0;
        bool end_seen = false;
        auto announce = [&]() {
            while (!end_seen && pub < seg + kSegRing) {
                const uint32_t ppar = pub % kSegRing, pk = pub / kSegRing;
                if (pub == seg) mbar_wait(&ctrl->seg_done[ppar], pk & 1u, p.err, DEV_TIMEOUT_FULL);
                else if (!mbar_try_wait(&ctrl->seg_done[ppar], pk & 1u)) break;
                const int32_t pchunk = ctrl->seg_chunk[ppar];
                if (pchunk < 0) { end_seen = true; ++pub; break; }
                const uint32_t pt0 = (uint32_t)pchunk << p.chunk_log2;
                const uint32_t pnent = min(T, g.num_tiles - pt0) * kConsumerWarps;
                uint32_t psum = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t e = (uint32_t)i * 32u + lane;
                    psum += e < pnent ? ctrl->cnt[ppar][e] : 0u;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
                if (!(p.debug & 1u)) lookback_publish(p.tile_state, (uint32_t)pchunk, psum, p.epoch, p.total_in);
                ++pub;
            }
        };
        // one staged entry -> registers (independent word loads: one L2 round trip for the whole warp)
        auto fetch = [&](uint32_t par, uint32_t r, uint32_t nrec, uint32_t (&wv)[12]) {
#pragma unroll
            for (int i = 0; i < 12; ++i) wv[i] = 0u;
            if (r < nrec) {
                const uint32_t dsc = ctrl->list[r];
                const uint32_t *src = reinterpret_cast<const uint32_t *>(stg0 + (size_t)(par * kConsumerWarps + (dsc >> 24)) * stg_warp_bytes +
                                                                         (size_t)(dsc & 0xffffffu) * p.stg_stride);
                const uint32_t nw = p.stg_stride >> 2;
#pragma unroll
                for (int i = 0; i < 12; ++i) if ((uint32_t)i < nw) wv[i] = src[i];
            }
        };
        auto emit = [&](uint64_t pos, uint64_t chunk_rec0, const uint32_t (&wv)[12]) {
            if (pos >= p.cap) return;
            uint8_t *dst = p.out + pos * p.O;
#pragma unroll
            for (int b = 0; b < 37; ++b)
                if ((uint32_t)b < p.O) dst[b] = (uint8_t)(wv[b >> 2] >> (8 * (b & 3)));
            if (p.out_index) {
                uint32_t rel = 0;
#pragma unroll
                for (int i = 0; i < 12; ++i) if ((uint32_t)i == (p.Ow >> 2)) rel = wv[i];
                p.out_index[pos] = p.index_base + chunk_rec0 + rel;
            }
        };
        while (true) {
            announce();
            const uint32_t par = seg % kSegRing;
            const int32_t chunk = ctrl->seg_chunk[par];
            if (chunk < 0) break;                      // (pub > seg here, so slot `par` has been waited for)
            const uint32_t t0 = (uint32_t)chunk << p.chunk_log2;
            const uint32_t nent = min(T, g.num_tiles - t0) * kConsumerWarps;
            const uint64_t chunk_rec0 = (uint64_t)t0 * g.tile_records;
            uint32_t cv[4];
            uint32_t sum = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t e = (uint32_t)i * 32u + lane;
                cv[i] = e < nent ? ctrl->cnt[par][e] : 0u;
                sum += cv[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const bool ovf = __any_sync(0xffffffffu, lane < kConsumerWarps && ctrl->ovfw[par][lane & 7u] != 0u);
            const bool narrow = p.stg_stride <= 48u;                 // entry fits the register path (k <= 128)
            const bool copy = sum != 0 && !ovf && !(p.debug & 4u);
            uint32_t wa[12], wb[12];
            if (copy) {
                // ---- the chunk's staged records in (tile, warp) = input order: entry e = t*8 + w is held by lane e%32 in
                // cv[e/32]; list[] gets (consumer warp, index in its staging list) for every record, in output order
                uint32_t row_carry = 0;             // records in entries of earlier i (row-major prefix)
                uint32_t col_carry = 0;             // records of warp (lane&7) in tiles of earlier i
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t v = cv[i];
                    uint32_t inc = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                        if ((int)lane >= o) inc += t;
                    }
                    const uint32_t row_total = __shfl_sync(0xffffffffu, inc, 31);
                    const uint32_t u8 = __shfl_up_sync(0xffffffffu, v, 8);
                    const uint32_t u16 = __shfl_up_sync(0xffffffffu, v, 16);
                    const uint32_t u24 = __shfl_up_sync(0xffffffffu, v, 24);
                    const uint32_t col_pre = (lane >= 8 ? u8 : 0u) + (lane >= 16 ? u16 : 0u) + (lane >= 24 ? u24 : 0u);
                    uint32_t col_total = v + __shfl_xor_sync(0xffffffffu, v, 8);
                    col_total += __shfl_xor_sync(0xffffffffu, col_total, 16);
                    const uint32_t first = row_carry + (inc - v), start = col_carry + col_pre;
                    for (uint32_t q = 0; q < v; ++q) ctrl->list[first + q] = ((lane & (kConsumerWarps - 1u)) << 24) | (start + q);
                    row_carry += row_total;
                    col_carry += col_total;
                }
                __syncwarp();
                // the entries do not depend on the prefix: fetch the first 64 now, so their L2 round trip overlaps the look-back
                if (narrow) { fetch(par, lane, sum, wa); fetch(par, lane + 32, sum, wb); }
            }
            uint64_t excl = 0;
            if (!(p.debug & 1u)) {
                uint64_t t_start = 0;
                while (!lookback_try(p.tile_state, (uint32_t)chunk, sum, p.epoch, p.total_in, excl)) {
                    announce();                        // keep announcing while predecessors are still running
                    if (t_start == 0) t_start = globaltimer_ns();
                    else if (globaltimer_ns() - t_start > kWatchdogNs) watchdog_fail(p.err, DEV_TIMEOUT_LOOKBACK);
                }
            }
            if (ovf) {
                if (lane == 0) {          // counts are exact, only the staged bytes were dropped: queue the chunk for a rewrite
                    const uint32_t at = atomicAdd(&p.dirty_ctl[0], 1u);
                    p.dirty_list[2ull * at] = (uint64_t)(uint32_t)chunk;
                    p.dirty_list[2ull * at + 1] = excl;
                }
            } else if (copy) {
                if (narrow) {
                    if (lane < sum) emit(excl + lane, chunk_rec0, wa);
                    if (lane + 32 < sum) emit(excl + lane + 32, chunk_rec0, wb);
                    for (uint32_t r = 64 + lane; r < sum; r += 32) {
                        fetch(par, r, sum, wa);
                        emit(excl + r, chunk_rec0, wa);
                    }
                } else {                            // very wide k-mers (k > 128): byte by byte
                    for (uint32_t r = lane; r < sum; r += 32) {
                        const uint64_t pos = excl + r;
                        if (pos >= p.cap) continue;
                        const uint32_t dsc = ctrl->list[r];
                        const uint8_t *src = stg0 + (size_t)(par * kConsumerWarps + (dsc >> 24)) * stg_warp_bytes + (size_t)(dsc & 0xffffffu) * p.stg_stride;
                        uint8_t *dst = p.out + pos * p.O;
                        for (uint32_t b = 0; b < p.O; ++b) dst[b] = src[b];
                        if (p.out_index) p.out_index[pos] = p.index_base + chunk_rec0 + *reinterpret_cast<const uint32_t *>(src + p.Ow);
                    }
                }
            }
            if ((uint32_t)chunk == p.num_chunks - 1 && lane == 0) *p.total_out = excl + sum;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->seg_free[par]);
            ++seg;
        }
        return;
    }

    // ---- consumer warps: warp w owns records [w*rpw, (w+1)*rpw) of every tile
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.body) & 15u);
    const uint32_t S = g.S;
    const uint32_t rpw = g.tile_records / kConsumerWarps;            // multiple of 32
    uint32_t poff[kFastParents > 0 ? kFastParents : 1];
#pragma unroll
    for (int i = 0; i < kFastParents; ++i) poff[i] = p.parent_off[i];
    const uint32_t lane_lt = (1u << lane) - 1u;
    const uint32_t words4 = p.cov_off >> 2;                          // 4-byte words of the k-mer (2 per 64-bit word)

    uint32_t seg = 0, fill = 0, ovf = 0;
    for (uint32_t it = 0;; ++it) {
        const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u;
        mbar_wait(&ctrl->full[st], ph, p.err, DEV_TIMEOUT_FULL);
        const int32_t tile = ctrl->tile_id[st];
        const uint32_t par = seg % kSegRing, k = seg / kSegRing;
        if (tile < 0) {
            if (k >= 1) mbar_wait(&ctrl->seg_free[par], (k - 1u) & 1u, p.err, DEV_TIMEOUT_LOOKBACK);
            if (warp == 0 && lane == 0) ctrl->seg_chunk[par] = -1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->seg_done[par]);
            break;
        }
        const uint32_t tin = (uint32_t)tile & (T - 1u);
        if (tin == 0) {
            // a new chunk: its staging buffer was last used two chunks ago; wait until the scan warp drained it
            if (k >= 1) mbar_wait(&ctrl->seg_free[par], (k - 1u) & 1u, p.err, DEV_TIMEOUT_LOOKBACK);
            fill = 0;
            ovf = 0;
            if (warp == 0 && lane == 0) ctrl->seg_chunk[par] = tile >> p.chunk_log2;
        }
        const uint64_t rec0 = (uint64_t)tile * g.tile_records;
        const uint64_t left = p.n - rec0;
        const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
        const uint8_t *tile_s = stage0 + (size_t)st * g.stage_bytes + mis;
        uint8_t *stg = stg0 + (size_t)(par * kConsumerWarps + warp) * stg_warp_bytes;
        uint32_t tile_cnt = 0;
#pragma unroll
        for (int j = 0; j < kMaxIters; ++j) {
            if (j * 32 < (int)rpw) {
                const uint32_t r = warp * rpw + (uint32_t)j * 32u + lane;
                const uint8_t *rec = tile_s + r * S;
                bool novel = false;
                if (r < nrec && !(p.debug & 2u)) {
                    // FindROIs.isNovel :72-82: coverage[child] > 0 (signed) && every listed parent == 0
                    uint32_t any_parent = 0;
                    if (NP >= 0) {
#pragma unroll
                        for (int i = 0; i < (NP > 0 ? NP : 0); ++i) any_parent |= lds_u32<ALIGNED4>(rec + poff[i]);
                    } else {
                        for (int i = 0; i < p.nparents; ++i) any_parent |= lds_u32<ALIGNED4>(rec + parents_s[i]);
                    }
                    const int32_t child_cov = (int32_t)lds_u32<ALIGNED4>(rec + p.child_off);
                    novel = (child_cov > 0) && (any_parent == 0);
                    if (p.debug & 32u) novel = novel && (child_cov == 0x7fffffff);      // diagnosis: loads kept, nothing novel
                }
                const uint32_t m = __ballot_sync(0xffffffffu, novel);
                if (m != 0) {
                    const uint32_t c = __popc(m);
                    if (fill + c <= p.stg_cap) {
                        if (novel && !(p.debug & 8u)) {
                            // staged entry = the output record (s words verbatim, child coverage LE, child edge byte:
                            // CortexGraphWriter.addRecord :106-138), padded to Ow, then the chunk-relative record number
                            uint8_t *dst = stg + (size_t)(fill + __popc(m & lane_lt)) * p.stg_stride;
                            for (uint32_t w4 = 0; w4 < words4; ++w4)
                                *reinterpret_cast<uint32_t *>(dst + 4u * w4) = lds_u32<ALIGNED4>(rec + 4u * w4);
                            *reinterpret_cast<uint32_t *>(dst + p.cov_off) = lds_u32<ALIGNED4>(rec + p.child_off);
                            dst[p.cov_off + 4u] = rec[p.edge_off + p.child];
                            *reinterpret_cast<uint32_t *>(dst + p.Ow) = tin * g.tile_records + r;
                        }
                    } else {
                        ovf = 1;
                    }
                    fill += c;
                    tile_cnt += c;
                }
            }
        }
        if (lane == 0) ctrl->cnt[par][tin * kConsumerWarps + warp] = tile_cnt;
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(&ctrl->empty[st]);      // the stage was only read; staging stores may still be in flight
        if (tin == T - 1u || (uint32_t)tile == g.num_tiles - 1u) {
            if (lane == 0) {
                ctrl->ovfw[par][warp] = ovf;
                mbar_arrive(&ctrl->seg_done[par]);       // release: orders this warp's staging + count stores (after __syncwarp)
            }
            ++seg;
        }
    }
}

// ------------------------------------------------------------------ dense chunks: rewrite with a known base
// Runs right behind the fast kernel over the chunks it queued (none on a sparse graph: every CTA draws one ticket
// and leaves).  The exclusive prefix of a queued chunk is already known, so a CTA streams the chunk's tiles again and
// writes in order with a block-local scan only: per tile the novel records are laid out in shared memory as the
// byte image of their output run (same 16-byte phase as the destination) and copied out with aligned 16-byte stores.
struct RedoCtrl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    int32_t tile_id[kMaxStages];
    uint64_t tile_base[kMaxStages];        // exclusive output prefix of the chunk the tile belongs to
    uint32_t wcnt[kConsumerWarps];
};
static_assert(sizeof(RedoCtrl) <= kCtrlBytes, "control block too large");

template <bool ALIGNED4, int NP>
__global__ void __launch_bounds__(kThreads, 2) redo_chunks_kernel(const __grid_constant__ ScanKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    RedoCtrl *ctrl = reinterpret_cast<RedoCtrl *>(smem);
    uint32_t *parents_s = reinterpret_cast<uint32_t *>(smem + kCtrlBytes);
    const uint32_t parents_bytes = NP >= 0 ? 0u : (((uint32_t)p.nparents * 4u + 127u) & ~127u);
    const TileGeom &g = p.g;
    uint8_t *image = smem + kCtrlBytes + parents_bytes;                              // tile_records*O + 32 bytes
    const uint32_t image_bytes = (g.tile_records * p.O + 32u + 127u) & ~127u;
    uint8_t *stage0 = image + image_bytes;
    // Launched with programmatic stream serialisation right behind the fast scan: the CTAs of this (almost always empty)
    // kernel are scheduled while the scan drains and wait here for its memory to be visible.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t ndirty = *reinterpret_cast<const volatile uint32_t *>(&p.dirty_ctl[0]);

    if (ndirty != 0) {
        if (NP < 0)
            for (int i = threadIdx.x; i < p.nparents; i += kThreads) parents_s[i] = p.cov_off + 4u * (uint32_t)p.parents[i];
        if (threadIdx.x == 0) {
            for (uint32_t i = 0; i < g.stages; ++i) {
                mbar_init(&ctrl->full[i], 1);
                mbar_init(&ctrl->empty[i], kConsumerWarps);
            }
            mbar_fence_init();
        }
        __syncthreads();
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.body) & 15u);
        if (warp == kProducerWarp) {
            if (lane == 0) {
                const uint64_t policy = make_evict_first_policy();
                const uint8_t *abase = p.body - mis;
                uint32_t it = 0;
                while (true) {
                    const uint32_t e = atomicAdd(&p.dirty_ctl[1], 1u);
                    if (e >= ndirty) {
                        const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u;
                        mbar_wait(&ctrl->empty[st], ph ^ 1u, p.err, DEV_TIMEOUT_EMPTY);
                        ctrl->tile_id[st] = -1;
                        mbar_arrive(&ctrl->full[st]);
                        break;
                    }
                    const uint32_t chunk = (uint32_t)p.dirty_list[2ull * e];
                    const uint64_t base = p.dirty_list[2ull * e + 1];
                    const uint32_t t0 = chunk << p.chunk_log2;
                    const uint32_t t1 = min(t0 + (1u << p.chunk_log2), g.num_tiles);
                    for (uint32_t t = t0; t < t1; ++t, ++it) {
                        const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u;
                        mbar_wait(&ctrl->empty[st], ph ^ 1u, p.err, DEV_TIMEOUT_EMPTY);
                        ctrl->tile_id[st] = (int32_t)t;
                        ctrl->tile_base[st] = base;
                        const uint64_t rec0 = (uint64_t)t * g.tile_records;
                        const uint64_t left = p.n - rec0;
                        const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
                        const uint32_t bytes = (nrec * g.S + mis + 15u) & ~15u;
                        mbar_arrive_expect_tx(&ctrl->full[st], bytes);
                        bulk_g2s(stage0 + (size_t)st * g.stage_bytes, abase + rec0 * g.S, bytes, &ctrl->full[st], policy);
                    }
                }
            }
        } else if (warp < kConsumerWarps) {
            const uint32_t S = g.S;
            const uint32_t rpw = g.tile_records / kConsumerWarps;
            const uint32_t T = 1u << p.chunk_log2;
            uint32_t poff[kFastParents > 0 ? kFastParents : 1];
#pragma unroll
            for (int i = 0; i < kFastParents; ++i) poff[i] = p.parent_off[i];
            const uint32_t lane_lt = (1u << lane) - 1u;
            uint64_t running = 0;                       // novel records of the chunk's earlier tiles
            for (uint32_t it = 0;; ++it) {
                const uint32_t st = it % g.stages, ph = (it / g.stages) & 1u;
                mbar_wait(&ctrl->full[st], ph, p.err, DEV_TIMEOUT_FULL);
                const int32_t tile = ctrl->tile_id[st];
                if (tile < 0) break;
                if (((uint32_t)tile & (T - 1u)) == 0) running = 0;
                const uint64_t out_pos0 = ctrl->tile_base[st] + running;       // output position of the tile's first novel record
                const uint64_t rec0 = (uint64_t)tile * g.tile_records;
                const uint64_t left = p.n - rec0;
                const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
                const uint8_t *tile_s = stage0 + (size_t)st * g.stage_bytes + mis;
                uint32_t masks[kMaxIters];
                uint32_t wtotal = 0;
#pragma unroll
                for (int j = 0; j < kMaxIters; ++j) {
                    masks[j] = 0;
                    if (j * 32 < (int)rpw) {
                        const uint32_t r = warp * rpw + (uint32_t)j * 32u + lane;
                        const uint8_t *rec = tile_s + r * S;
                        bool novel = false;
                        if (r < nrec) {
                            uint32_t any_parent = 0;
                            if (NP >= 0) {
#pragma unroll
                                for (int i = 0; i < (NP > 0 ? NP : 0); ++i) any_parent |= lds_u32<ALIGNED4>(rec + poff[i]);
                            } else {
                                for (int i = 0; i < p.nparents; ++i) any_parent |= lds_u32<ALIGNED4>(rec + parents_s[i]);
                            }
                            const int32_t child_cov = (int32_t)lds_u32<ALIGNED4>(rec + p.child_off);
                            novel = (child_cov > 0) && (any_parent == 0);
                        }
                        masks[j] = __ballot_sync(0xffffffffu, novel);
                        wtotal += __popc(masks[j]);
                    }
                }
                if (lane == 0) ctrl->wcnt[warp] = wtotal;
                named_bar_sync(1, kConsumerThreads);                            // counts visible; image free (see the barrier below)
                uint32_t before = 0, tile_total = 0;
#pragma unroll
                for (int w = 0; w < kConsumerWarps; ++w) {
                    const uint32_t cw = ctrl->wcnt[w];
                    if (w < (int)warp) before += cw;
                    tile_total += cw;
                }
                // byte image of the tile's output run, phase-aligned with its destination
                uint8_t *dst0 = p.out + out_pos0 * p.O;
                const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(dst0) & 15u);
                uint32_t at = before;
#pragma unroll
                for (int j = 0; j < kMaxIters; ++j) {
                    const uint32_t m = masks[j];
                    if (j * 32 < (int)rpw && m != 0) {
                        if ((m >> lane) & 1u) {
                            const uint32_t r = warp * rpw + (uint32_t)j * 32u + lane;
                            const uint8_t *rec = tile_s + r * S;
                            const uint32_t q = at + __popc(m & lane_lt);
                            uint8_t *d = image + shift + q * p.O;
                            for (uint32_t b = 0; b < p.cov_off; ++b) d[b] = rec[b];
#pragma unroll
                            for (uint32_t b = 0; b < 4; ++b) d[p.cov_off + b] = rec[p.child_off + b];
                            d[p.cov_off + 4u] = rec[p.edge_off + p.child];
                            if (p.out_index && out_pos0 + q < p.cap) p.out_index[out_pos0 + q] = p.index_base + rec0 + r;
                        }
                        at += __popc(m);
                    }
                }
                named_bar_sync(1, kConsumerThreads);                            // image complete
                // records beyond cap are dropped
                uint64_t keep = 0;
                if (out_pos0 < p.cap) keep = (p.cap - out_pos0 < tile_total) ? (p.cap - out_pos0) : tile_total;
                const uint32_t nbytes = (uint32_t)keep * p.O;
                const uint32_t head = min(nbytes, (16u - shift) & 15u);
                const uint32_t ctid = threadIdx.x;                               // 0..255
                if (ctid < head) dst0[ctid] = image[shift + ctid];
                const uint32_t body16 = (nbytes - head) >> 4;
                const uint4 *src4 = reinterpret_cast<const uint4 *>(image + shift + head);
                uint4 *dst4 = reinterpret_cast<uint4 *>(dst0 + head);
                for (uint32_t i = ctid; i < body16; i += kConsumerThreads) dst4[i] = src4[i];
                const uint32_t done = head + (body16 << 4);
                if (ctid < nbytes - done) dst0[done + ctid] = image[shift + done + ctid];
                running += tile_total;
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctrl->empty[st]);                   // (the next tile's first barrier also guards the image)
            }
        }
    }
    // ---- the last CTA to leave resets the queue for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(&p.dirty_ctl[2], 1u);
        if (done == gridDim.x - 1) {
            p.dirty_ctl[0] = 0; p.dirty_ctl[1] = 0; p.dirty_ctl[2] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------ K1: decode into columns (keys / coverage / edges)
template <bool ALIGNED4>
__global__ void __launch_bounds__(kThreads, 2) decode_columns_kernel(const __grid_constant__ DecodeKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem);
    uint8_t *stage0 = smem + kCtrlBytes;
    const TileGeom &g = p.g;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < g.stages; ++i) {
            mbar_init(&ctrl->full[i], 1);
            mbar_init(&ctrl->empty[i], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (warp == kProducerWarp) {
        if (lane == 0)
            producer_loop(ctrl->full, ctrl->empty, ctrl->tile_id, stage0, p.body, p.n, g, p.tile_counter, p.ticket_base, p.err);
        return;
    }
    if (warp == kScanWarp) return;                  // no compaction in the column decode
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.body) & 15u);
    const uint32_t S = g.S;

    for (uint32_t it = 0;; ++it) {
        const uint32_t st = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1u;
        mbar_wait(&ctrl->full[st], ph, p.err, DEV_TIMEOUT_FULL);
        const int32_t tile = ctrl->tile_id[st];
        if (tile < 0) break;
        const uint64_t rec0 = (uint64_t)tile * g.tile_records;
        const uint64_t left = p.n - rec0;
        const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
        const uint8_t *tile_s = stage0 + (size_t)st * g.stage_bytes + mis;

        if (p.words) {      // element e = (record, word): consecutive threads write consecutive 8-byte words
            const uint32_t total = nrec * p.s;
            uint64_t *dst = p.words + rec0 * p.s;
            for (uint32_t e = threadIdx.x; e < total; e += kConsumerThreads) {
                const uint32_t r = e / p.s, w = e - r * p.s;
                dst[e] = lds_u64<ALIGNED4>(tile_s + r * S + 8u * w);
            }
        }
        if (p.cov) {
            const uint32_t total = nrec * p.c;
            int32_t *dst = p.cov + rec0 * p.c;
            for (uint32_t e = threadIdx.x; e < total; e += kConsumerThreads) {
                const uint32_t r = e / p.c, j = e - r * p.c;
                dst[e] = (int32_t)lds_u32<ALIGNED4>(tile_s + r * S + p.cov_off + 4u * j);
            }
        }
        if (p.edges) {
            const uint32_t total = nrec * p.c;
            uint8_t *dst = p.edges + rec0 * p.c;
            for (uint32_t e = threadIdx.x; e < total; e += kConsumerThreads) {
                const uint32_t r = e / p.c, j = e - r * p.c;
                dst[e] = tile_s[r * S + p.edge_off + j];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->empty[st]);
    }
}

// ------------------------------------------------------------------ host side
int pick_geometry(uint64_t n, uint32_t S, uint32_t extra_smem, int max_smem_optin, int ctas_per_sm, TileGeom &g) {
    const Options &o = options();
    if (S == 0) return fail(CC_ERR_ARG, "record size is zero");
    uint32_t target = (uint32_t)std::max(4096, o.scan_tile_bytes);
    uint32_t R = (target / S) & ~31u;
    if (R < 32) R = 32;
    if (R > (uint32_t)(kMaxIters * kConsumerThreads)) R = kMaxIters * kConsumerThreads;
    uint32_t stage_bytes = (R * S + 32u + 127u) & ~127u;
    int stages = std::min(std::max(o.scan_stages, 2), kMaxStages);
    const int64_t budget = (int64_t)max_smem_optin / std::max(ctas_per_sm, 1) - 1024 /*per-CTA reservation*/ - kCtrlBytes - extra_smem;
    while (stages > 2 && (int64_t)stages * stage_bytes > budget) --stages;
    if ((int64_t)stages * stage_bytes > budget)
        return fail(CC_ERR_UNSUPPORTED, "record size %u bytes does not fit the shared-memory tile ring", S);
    if (((uint64_t)R * S + 31u) >= (1u << 20))
        return fail(CC_ERR_UNSUPPORTED, "tile exceeds the mbarrier transaction limit");
    uint64_t ntiles = (n + R - 1) / R;
    if (ntiles >= (1ull << 31)) return fail(CC_ERR_UNSUPPORTED, "too many tiles");
    g.S = S;
    g.tile_records = R;
    g.num_tiles = (uint32_t)ntiles;
    g.stages = (uint32_t)stages;
    g.stage_bytes = stage_bytes;
    g.iters = (R + kConsumerThreads - 1) / kConsumerThreads;
    return CC_OK;
}

struct DeviceLimits {
    int smem_optin = 0;
    int sm_count = 0;
};
int device_limits(DeviceLimits &l) {
    int dev = 0;
    CC_CUDA(cudaGetDevice(&dev));
    CC_CUDA(cudaDeviceGetAttribute(&l.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CC_CUDA(cudaDeviceGetAttribute(&l.sm_count, cudaDevAttrMultiProcessorCount, dev));
    return CC_OK;
}

}  // namespace

uint64_t scan_tiles_for(uint64_t n, uint32_t s, uint32_t c) {
    const uint32_t S = 8u * s + 5u * c;
    if (S == 0 || n == 0) return 1;
    uint32_t R = ((uint32_t)std::max(4096, options().scan_tile_bytes) / S) & ~31u;
    if (R < 32) R = 32;
    if (R > (uint32_t)(kMaxIters * kConsumerThreads)) R = kMaxIters * kConsumerThreads;
    return (n + R - 1) / R;
}

int ScanWorkspace::ensure(uint64_t ntiles, uint32_t nparents) {
    // cudaMemset runs on the legacy default stream, which the handle's non-blocking stream does not wait for:
    // settle it before anything on another stream may touch the fresh allocation.
    bool fresh = false;
    if (!tile_counter) {
        CC_CUDA(cudaMalloc(&tile_counter, 256));
        CC_CUDA(cudaMemset(tile_counter, 0, 256));
        CC_CUDA(cudaMalloc(&totals, 256));
        CC_CUDA(cudaMemset(totals, 0, 256));
        // the watchdog word lives in mapped host memory: the host can still read it after a kernel trapped
        if (cudaHostAlloc(reinterpret_cast<void **>(&host_error), 256, cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer(reinterpret_cast<void **>(&dev_error), host_error, 0) == cudaSuccess) {
            memset(host_error, 0, 256);
        } else {
            cudaGetLastError();
            if (host_error) { cudaFreeHost(host_error); host_error = nullptr; }
            CC_CUDA(cudaMalloc(&dev_error, 256));
            CC_CUDA(cudaMemset(dev_error, 0, 256));
        }
        fresh = true;
    }
    if (poisoned) {
        // an earlier launch failed half-way: the device-side ticket counter, the queue of dense chunks and the host
        // mirrors may disagree.  Start over from zero (legacy stream + synchronise: this is the error path).
        CC_CUDA(cudaDeviceSynchronize());
        CC_CUDA(cudaMemset(tile_counter, 0, 256));
        CC_CUDA(cudaMemset(totals, 0, 256));
        if (tile_state) CC_CUDA(cudaMemset(tile_state, 0, tile_state_cap * sizeof(uint64_t)));
        ticket_base = 0;
        epoch = 0;
        poisoned = false;
        fresh = true;
    }
    if (ntiles > tile_state_cap) {
        if (tile_state) CC_CUDA(cudaFree(tile_state));
        uint64_t cap = std::max<uint64_t>(ntiles, 1024);
        CC_CUDA(cudaMalloc(&tile_state, cap * sizeof(uint64_t)));
        CC_CUDA(cudaMemset(tile_state, 0, cap * sizeof(uint64_t)));
        if (dirty_list) CC_CUDA(cudaFree(dirty_list));
        CC_CUDA(cudaMalloc(&dirty_list, cap * 2 * sizeof(uint64_t)));
        tile_state_cap = cap;
        epoch = 0;
        fresh = true;
    }
    if (nparents > parents_cap) {
        if (parents) CC_CUDA(cudaFree(parents));
        uint32_t cap = std::max<uint32_t>(nparents, 64);
        CC_CUDA(cudaMalloc(&parents, cap * sizeof(int32_t)));
        parents_cap = cap;
    }
    if (fresh) CC_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    return CC_OK;
}

int ScanWorkspace::ensure_scratch(size_t bytes) {
    if (bytes > scratch_bytes) {
        if (scratch) CC_CUDA(cudaFree(scratch));
        scratch = nullptr;
        scratch_bytes = 0;
        CC_CUDA(cudaMalloc(&scratch, bytes + 256));
        scratch_bytes = bytes;
    }
    return CC_OK;
}

void ScanWorkspace::release() {
    if (scratch) cudaFree(scratch);
    if (tile_state) cudaFree(tile_state);
    if (dirty_list) cudaFree(dirty_list);
    if (tile_counter) cudaFree(tile_counter);
    if (totals) cudaFree(totals);
    if (parents) cudaFree(parents);
    if (host_error) cudaFreeHost(host_error);
    else if (dev_error) cudaFree(dev_error);
    *this = ScanWorkspace();
}

namespace {

struct FastGeom {
    TileGeom g;
    uint32_t chunk_log2 = 0, num_chunks = 0, stg_stride = 0, stg_cap = 0, Ow = 0, stg_bytes = 0;
    bool ok = false;
};

// Geometry of the fast kernel: tiles of a multiple of 256 records (8 warps x 32 lanes), a staging area, >= 2 stages.
FastGeom pick_fast_geometry(uint64_t n, uint32_t S, uint32_t O, uint32_t extra_smem, int max_smem_optin, int ctas_per_sm) {
    FastGeom f;
    const Options &o = options();
    if (!o.scan_fast || S == 0) return f;
    uint32_t R = ((uint32_t)std::max(4096, o.scan_tile_bytes) / S) & ~255u;
    if (R < 256) R = 256;
    if (R > (uint32_t)(kMaxIters * kConsumerThreads)) R = kMaxIters * kConsumerThreads;
    const uint32_t stage_bytes = (R * S + 32u + 127u) & ~127u;
    f.Ow = (O + 3u) & ~3u;
    f.stg_stride = f.Ow + 4u;
    f.stg_cap = std::min<uint32_t>((uint32_t)std::max(256, o.scan_stage_buf_bytes) / f.stg_stride, kMaxListed / kConsumerWarps);
    if (f.stg_cap < 4) return f;
    f.stg_bytes = (uint32_t)kSegRing * kConsumerWarps * f.stg_cap * f.stg_stride;          // per CTA, in the global scratch
    int stages = std::min(std::max(o.scan_stages, 2), kMaxStages);
    const int64_t budget = (int64_t)max_smem_optin / std::max(ctas_per_sm, 1) - 1024 - kFastCtrlBytes - extra_smem;
    while (stages > 2 && (int64_t)stages * stage_bytes > budget) --stages;
    if ((int64_t)stages * stage_bytes > budget) return f;
    if (((uint64_t)R * S + 31u) >= (1u << 20)) return f;
    const uint64_t ntiles = (n + R - 1) / R;
    if (ntiles >= (1ull << 31)) return f;
    int lg = 0;
    while ((1 << (lg + 1)) <= std::min(std::max(o.scan_chunk_tiles, 1), kMaxChunkTiles)) ++lg;
    f.chunk_log2 = (uint32_t)lg;
    f.num_chunks = (uint32_t)((ntiles + (1ull << lg) - 1) >> lg);
    f.g.S = S; f.g.tile_records = R; f.g.num_tiles = (uint32_t)ntiles; f.g.stages = (uint32_t)stages;
    f.g.stage_bytes = stage_bytes; f.g.iters = R / kConsumerThreads;
    f.ok = true;
    return f;
}

int next_epoch(ScanWorkspace &ws, cudaStream_t st) {
    // A fresh epoch makes every descriptor of earlier launches read as "invalid" without a memset.
    ws.epoch = (ws.epoch + 1) & kEpochMask;
    if (ws.epoch == 0) {
        CC_CUDA(cudaMemsetAsync(ws.tile_state, 0, ws.tile_state_cap * sizeof(uint64_t), st));
        CC_CUDA(cudaMemsetAsync(ws.totals + 16, 0, 8, st));     // the overflow flag holds an epoch too
        ws.epoch = 1;
    }
    return CC_OK;
}

}  // namespace

// ws.parents must already hold the parent list; ws.ensure() must have been called for this n.
int launch_scan_novel(const ScanArgs &a, ScanWorkspace &ws, int sm_count, cudaStream_t st) {
    if (a.n == 0) {
        // nothing to scan: total_out = total_in
        if (a.total_in) CC_CUDA(cudaMemcpyAsync(a.total_out, a.total_in, 8, cudaMemcpyDeviceToDevice, st));
        else CC_CUDA(cudaMemsetAsync(a.total_out, 0, 8, st));
        return CC_OK;
    }
    DeviceLimits lim;
    if (int rc = device_limits(lim)) return rc;
    const Options &o = options();
    const uint32_t S = 8u * a.s + 5u * a.c;
    const bool fastp = a.nparents <= kFastParents;
    const uint32_t parents_bytes = fastp ? 0u : (((uint32_t)a.nparents * 4u + 127u) & ~127u);
    const int ctas = std::max(1, std::min(o.scan_ctas_per_sm, 4));
    ScanKParams p{};
    p.body = a.body; p.n = a.n; p.index_base = a.index_base;
    p.s = a.s; p.c = a.c; p.cov_off = 8u * a.s; p.edge_off = 8u * a.s + 4u * a.c; p.O = 8u * a.s + 5u;
    p.child = a.child; p.nparents = a.nparents; p.parents = ws.parents;
    p.child_off = p.cov_off + 4u * (uint32_t)a.child;
    for (int i = 0; i < kFastParents; ++i)
        p.parent_off[i] = (fastp && i < a.nparents) ? p.cov_off + 4u * (uint32_t)a.parent_list[i] : p.child_off;
    p.out = a.out_records; p.out_index = a.out_index; p.cap = a.cap;
    p.total_in = a.total_in; p.total_out = a.total_out;
    p.tile_state = ws.tile_state; p.tile_counter = ws.tile_counter;
    p.err = ws.dev_error; p.debug = (uint32_t)o.scan_debug;
    p.overflow = reinterpret_cast<uint32_t *>(ws.totals + 16);
    p.dirty_list = ws.dirty_list;
    p.dirty_ctl = reinterpret_cast<uint32_t *>(ws.totals + 20);
    const bool aligned4 = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.body) & 3u) == 0);
    const int np_slot = fastp ? a.nparents : kFastParents + 1;
    using Kern = void (*)(const ScanKParams);

    // ---- fast kernel (chunked, deferred look-back)
    // Every step that can fail without having touched the workspace state (allocation, attribute calls, the "does the
    // rewrite kernel fit" test) comes first; the host mirrors of the device counters (ticket_base, epoch) advance only
    // with a successful launch, and a failure after that point poisons the workspace so that the next call resets it.
    const FastGeom f = pick_fast_geometry(a.n, S, p.O, parents_bytes, lim.smem_optin, ctas);
    if (f.ok && f.num_chunks <= ws.tile_state_cap) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sm_count * ctas, f.num_chunks);
        const size_t smem = kFastCtrlBytes + parents_bytes + (size_t)f.g.stages * f.g.stage_bytes;
        const uint32_t image_bytes = (f.g.tile_records * p.O + 32u + 127u) & ~127u;
        int rstages = (int)f.g.stages;
        const int64_t rbudget = (int64_t)lim.smem_optin / ctas - 1024 - kCtrlBytes - parents_bytes - image_bytes;
        while (rstages > 2 && (int64_t)rstages * f.g.stage_bytes > rbudget) --rstages;
        if ((int64_t)rstages * f.g.stage_bytes > rbudget) return fail(CC_ERR_UNSUPPORTED, "record size %u bytes does not fit the rewrite kernel", S);
        const size_t rsmem = kCtrlBytes + parents_bytes + image_bytes + (size_t)rstages * f.g.stage_bytes;
        static const Kern ftable[2][kFastParents + 2] = {
            {scan_novel_fast_kernel<false, 0>, scan_novel_fast_kernel<false, 1>, scan_novel_fast_kernel<false, 2>,
             scan_novel_fast_kernel<false, 3>, scan_novel_fast_kernel<false, 4>, scan_novel_fast_kernel<false, -1>},
            {scan_novel_fast_kernel<true, 0>, scan_novel_fast_kernel<true, 1>, scan_novel_fast_kernel<true, 2>,
             scan_novel_fast_kernel<true, 3>, scan_novel_fast_kernel<true, 4>, scan_novel_fast_kernel<true, -1>}};
        static const Kern rtable[2][kFastParents + 2] = {
            {redo_chunks_kernel<false, 0>, redo_chunks_kernel<false, 1>, redo_chunks_kernel<false, 2>, redo_chunks_kernel<false, 3>,
             redo_chunks_kernel<false, 4>, redo_chunks_kernel<false, -1>},
            {redo_chunks_kernel<true, 0>, redo_chunks_kernel<true, 1>, redo_chunks_kernel<true, 2>, redo_chunks_kernel<true, 3>,
             redo_chunks_kernel<true, 4>, redo_chunks_kernel<true, -1>}};
        Kern kern = ftable[aligned4 ? 1 : 0][np_slot];
        Kern rkern = rtable[aligned4 ? 1 : 0][np_slot];
        if (int rc = ws.ensure_scratch((size_t)grid * f.stg_bytes)) return rc;
        CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CC_CUDA(cudaFuncSetAttribute(rkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        if (int rc = next_epoch(ws, st)) { ws.poisoned = true; return rc; }
        ScanKParams q = p;
        q.g = f.g; q.epoch = ws.epoch;
        q.chunk_log2 = f.chunk_log2; q.num_chunks = f.num_chunks; q.stg_stride = f.stg_stride; q.stg_cap = f.stg_cap; q.Ow = f.Ow;
        q.ticket_base = ws.ticket_base;
        q.stg_scratch = ws.scratch;
        kern<<<grid, kThreads, smem, st>>>(q);
        if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) return cuda_fail(e, "scan_novel_fast_kernel launch", __FILE__, __LINE__);   // nothing ran: state unchanged
        count_launch();
        ws.ticket_base += f.num_chunks + grid;      // every CTA draws its chunks plus one terminal ticket
        // ---- rewrite of the chunks whose staging overflowed (dense novelty); a no-op launch on sparse graphs
        ScanKParams r = q;
        r.g.stages = (uint32_t)rstages;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = rsmem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = options().scan_pdl ? 1 : 0;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaError_t e = cudaLaunchKernelEx(&cfg, rkern, r); e != cudaSuccess) {
            cudaGetLastError();
            ws.poisoned = true;                     // the fast kernel may have queued chunks nobody will drain
            return cuda_fail(e, "redo_chunks_kernel launch", __FILE__, __LINE__);
        }
        count_launch();
        return CC_OK;
    }

    // ---- general kernel (per-tile look-back, any density): alone, or behind the fast kernel and run only on overflow
    if (int rc = pick_geometry(a.n, S, parents_bytes, lim.smem_optin, ctas, p.g)) return rc;
    if (p.g.num_tiles > ws.tile_state_cap) return fail(CC_ERR_ARG, "scan workspace too small");
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sm_count * ctas, p.g.num_tiles);
    p.ticket_base = ws.ticket_base;
    p.run_if = nullptr;
    p.run_expect = 0;
    p.skip_tickets = p.g.num_tiles + grid;
    const size_t smem = kCtrlBytes + parents_bytes + (size_t)p.g.stages * p.g.stage_bytes;
    static const Kern table[2][kFastParents + 2] = {
        {scan_novel_kernel<false, 0>, scan_novel_kernel<false, 1>, scan_novel_kernel<false, 2>, scan_novel_kernel<false, 3>,
         scan_novel_kernel<false, 4>, scan_novel_kernel<false, -1>},
        {scan_novel_kernel<true, 0>, scan_novel_kernel<true, 1>, scan_novel_kernel<true, 2>, scan_novel_kernel<true, 3>,
         scan_novel_kernel<true, 4>, scan_novel_kernel<true, -1>}};
    Kern kern = table[aligned4 ? 1 : 0][np_slot];
    CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int rc = next_epoch(ws, st)) { ws.poisoned = true; return rc; }
    p.epoch = ws.epoch;
    kern<<<grid, kThreads, smem, st>>>(p);
    CC_CUDA(cudaGetLastError());                    // a failed launch ran nothing: the mirror stays where it was
    count_launch();
    ws.ticket_base += p.g.num_tiles + grid;         // drawn by the CTAs, or added by block 0 when the launch is skipped
    return CC_OK;
}

int launch_decode_columns(const uint8_t *dev_body, uint64_t n, uint32_t s, uint32_t c,
                          uint64_t *dev_words, int32_t *dev_cov, uint8_t *dev_edges, ScanWorkspace &ws, int sm_count,
                          cudaStream_t st) {
    if (n == 0) return CC_OK;
    DeviceLimits lim;
    if (int rc = device_limits(lim)) return rc;
    if (int rc = ws.ensure(0, 0)) return rc;
    const uint32_t S = 8u * s + 5u * c;
    DecodeKParams p{};
    if (int rc = pick_geometry(n, S, 0, lim.smem_optin, 2, p.g)) return rc;
    p.body = dev_body; p.n = n; p.s = s; p.c = c; p.cov_off = 8u * s; p.edge_off = 8u * s + 4u * c;
    p.words = dev_words; p.cov = dev_cov; p.edges = dev_edges;
    p.tile_counter = ws.tile_counter; p.err = ws.dev_error;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sm_count * 2, p.g.num_tiles);
    p.ticket_base = ws.ticket_base;
    const size_t smem = kCtrlBytes + (size_t)p.g.stages * p.g.stage_bytes;
    const bool aligned4 = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(dev_body) & 3u) == 0);
    auto kern = aligned4 ? decode_columns_kernel<true> : decode_columns_kernel<false>;
    CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(p);
    CC_CUDA(cudaGetLastError());
    count_launch();
    ws.ticket_base += p.g.num_tiles + grid;
    return CC_OK;
}

}  // namespace cc
