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
#include <algorithm>

#include "cc_internal.hpp"
#include "device_utils.cuh"

namespace cc {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;     // + 1 producer warp
constexpr int kMaxIters = 8;                        // tile_records <= kMaxIters * kConsumerThreads
constexpr int kMaxStages = 8;
constexpr uint32_t kCtrlBytes = 1024;

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
    int *err;
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
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    int32_t tile_id[kMaxStages];
    uint32_t cnt[kMaxIters * kConsumerWarps];
    uint32_t pre[kMaxIters * kConsumerWarps];
    uint64_t tile_base;
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block too large");

// ------------------------------------------------------------------ producer: the TMA side of the ring
__device__ __forceinline__ void producer_loop(Ctrl *ctrl, uint8_t *stage0, const uint8_t *body, uint64_t n,
                                              const TileGeom &g, uint32_t *tile_counter, uint32_t ticket_base, int *err) {
    const uint64_t policy = make_evict_first_policy();
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(body) & 15u);
    const uint8_t *abase = body - mis;
    for (uint32_t it = 0;; ++it) {
        const uint32_t st = it % g.stages;
        const uint32_t ph = (it / g.stages) & 1u;
        mbar_wait(&ctrl->empty[st], ph ^ 1u, err, DEV_TIMEOUT_EMPTY);
        const uint32_t t = atomicAdd(tile_counter, 1u) - ticket_base;
        if (t >= g.num_tiles) {
            ctrl->tile_id[st] = -1;
            mbar_arrive(&ctrl->full[st]);           // sentinel: consumers see tile_id < 0 and leave
            break;
        }
        ctrl->tile_id[st] = (int32_t)t;
        const uint64_t rec0 = (uint64_t)t * g.tile_records;
        const uint64_t left = n - rec0;
        const uint32_t nrec = left < g.tile_records ? (uint32_t)left : g.tile_records;
        const uint32_t bytes = (nrec * g.S + mis + 15u) & ~15u;
        mbar_arrive_expect_tx(&ctrl->full[st], bytes);
        bulk_g2s(stage0 + (size_t)st * g.stage_bytes, abase + rec0 * g.S, bytes, &ctrl->full[st], policy);
    }
}

__device__ __forceinline__ void ring_init(Ctrl *ctrl, uint32_t stages) {
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < stages; ++i) {
            mbar_init(&ctrl->full[i], 1);
            mbar_init(&ctrl->empty[i], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// ------------------------------------------------------------------ decoupled look-back (one warp)
__device__ __forceinline__ uint64_t pack_desc(uint64_t flag, uint32_t epoch, uint64_t value) {
    return flag | ((uint64_t)(epoch & kEpochMask) << 40) | (value & kValueMask);
}

// Returns the exclusive prefix of `tile` (novel records in all earlier tiles + base) and publishes this
// tile's inclusive prefix.  Called by all 32 lanes of one warp.
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

// ------------------------------------------------------------------ K1+K2: novelty scan
template <bool ALIGNED4>
__global__ void __launch_bounds__(kThreads, 2) scan_novel_kernel(const __grid_constant__ ScanKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem);
    int32_t *parents_s = reinterpret_cast<int32_t *>(smem + kCtrlBytes);
    const uint32_t parents_bytes = ((uint32_t)p.nparents * 4u + 127u) & ~127u;
    uint8_t *stage0 = smem + kCtrlBytes + parents_bytes;

    const TileGeom &g = p.g;
    for (int i = threadIdx.x; i < p.nparents; i += kThreads) parents_s[i] = p.parents[i];
    ring_init(ctrl, g.stages);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (warp == kConsumerWarps) {
        if (lane == 0) producer_loop(ctrl, stage0, p.body, p.n, g, p.tile_counter, p.ticket_base, p.err);
        return;
    }

    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p.body) & 15u);
    const uint32_t S = g.S;
    const uint32_t nslots = g.iters * kConsumerWarps;

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

        // ---- predicate (FindROIs.isNovel): coverage[child] > 0 (signed) && all listed parents == 0
        uint32_t masks[kMaxIters];
#pragma unroll
        for (int j = 0; j < kMaxIters; ++j) {
            masks[j] = 0;
            if (j < (int)g.iters) {
                const uint32_t r = (uint32_t)j * kConsumerThreads + threadIdx.x;
                bool novel = false;
                if (r < nrec) {
                    const uint8_t *cov = tile_s + r * S + p.cov_off;
                    uint32_t any_parent = 0;
                    for (int i = 0; i < p.nparents; ++i) any_parent |= lds_u32<ALIGNED4>(cov + 4 * parents_s[i]);
                    const int32_t child_cov = (int32_t)lds_u32<ALIGNED4>(cov + 4 * p.child);
                    novel = (child_cov > 0) && (any_parent == 0);
                }
                masks[j] = __ballot_sync(0xffffffffu, novel);
                if (lane == 0) ctrl->cnt[j * kConsumerWarps + warp] = __popc(masks[j]);
            }
        }
        named_bar_sync(1, kConsumerThreads);

        // ---- intra-tile exclusive scan of the (iteration, warp) counts + look-back, by warp 0
        if (warp == 0) {
            const uint32_t a = lane < nslots ? ctrl->cnt[lane] : 0u;
            const uint32_t b = lane + 32 < nslots ? ctrl->cnt[lane + 32] : 0u;
            uint32_t ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o);
                uint32_t tb = __shfl_up_sync(0xffffffffu, ib, o);
                if ((int)lane >= o) { ia += ta; ib += tb; }
            }
            const uint32_t tot_a = __shfl_sync(0xffffffffu, ia, 31);
            const uint32_t tot_b = __shfl_sync(0xffffffffu, ib, 31);
            if (lane < nslots) ctrl->pre[lane] = ia - a;
            if (lane + 32 < nslots) ctrl->pre[lane + 32] = tot_a + ib - b;
            const uint64_t agg = (uint64_t)tot_a + tot_b;
            const uint64_t excl = lookback(p.tile_state, (uint32_t)tile, agg, p.epoch, p.total_in, p.err);
            if (lane == 0) {
                ctrl->tile_base = excl;
                if ((uint32_t)tile == g.num_tiles - 1) *p.total_out = excl + agg;
            }
        }
        named_bar_sync(1, kConsumerThreads);

        // ---- ordered write-out: each warp copies its novel records' output bytes (8s words verbatim,
        //      child coverage, child edge) as one contiguous byte run, 32 consecutive bytes per instruction.
        const uint64_t tile_base = ctrl->tile_base;
#pragma unroll
        for (int j = 0; j < kMaxIters; ++j) {
            if (j < (int)g.iters && masks[j] != 0) {
                const uint32_t m = masks[j];
                const uint64_t wbase = tile_base + ctrl->pre[j * kConsumerWarps + warp];
                const uint32_t cnt = __popc(m);
                const uint32_t rbase = (uint32_t)j * kConsumerThreads + warp * 32u;
                const uint32_t total = cnt * p.O;
                for (uint32_t b = lane; b < total; b += 32) {
                    const uint32_t q = b / p.O;
                    const uint32_t ob = b - q * p.O;
                    if (wbase + q < p.cap) {
                        const uint32_t src_lane = __fns(m, 0, q + 1);
                        const uint8_t *rec = tile_s + (rbase + src_lane) * S;
                        uint32_t off;
                        if (ob < p.cov_off) off = ob;
                        else if (ob < p.cov_off + 4) off = p.cov_off + 4 * p.child + (ob - p.cov_off);
                        else off = p.edge_off + p.child;
                        p.out[(wbase + q) * p.O + ob] = rec[off];
                    }
                }
                if (p.out_index && ((m >> lane) & 1u)) {
                    const uint64_t pos = wbase + __popc(m & ((1u << lane) - 1u));
                    if (pos < p.cap) p.out_index[pos] = p.index_base + rec0 + rbase + lane;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->empty[st]);
    }
}

// ------------------------------------------------------------------ K1: decode into columns (keys / coverage / edges)
template <bool ALIGNED4>
__global__ void __launch_bounds__(kThreads, 2) decode_columns_kernel(const __grid_constant__ DecodeKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    Ctrl *ctrl = reinterpret_cast<Ctrl *>(smem);
    uint8_t *stage0 = smem + kCtrlBytes;
    const TileGeom &g = p.g;
    ring_init(ctrl, g.stages);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (warp == kConsumerWarps) {
        if (lane == 0) producer_loop(ctrl, stage0, p.body, p.n, g, p.tile_counter, p.ticket_base, p.err);
        return;
    }
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
    if (!tile_counter) {
        CC_CUDA(cudaMalloc(&tile_counter, 256));
        CC_CUDA(cudaMemset(tile_counter, 0, 256));
        CC_CUDA(cudaMalloc(&totals, 256));
        CC_CUDA(cudaMemset(totals, 0, 256));
        CC_CUDA(cudaMalloc(&dev_error, 256));
        CC_CUDA(cudaMemset(dev_error, 0, 256));
    }
    if (ntiles > tile_state_cap) {
        if (tile_state) CC_CUDA(cudaFree(tile_state));
        uint64_t cap = std::max<uint64_t>(ntiles, 1024);
        CC_CUDA(cudaMalloc(&tile_state, cap * sizeof(uint64_t)));
        CC_CUDA(cudaMemset(tile_state, 0, cap * sizeof(uint64_t)));
        tile_state_cap = cap;
        epoch = 0;
    }
    if (nparents > parents_cap) {
        if (parents) CC_CUDA(cudaFree(parents));
        uint32_t cap = std::max<uint32_t>(nparents, 64);
        CC_CUDA(cudaMalloc(&parents, cap * sizeof(int32_t)));
        parents_cap = cap;
    }
    return CC_OK;
}

void ScanWorkspace::release() {
    if (tile_state) cudaFree(tile_state);
    if (tile_counter) cudaFree(tile_counter);
    if (totals) cudaFree(totals);
    if (parents) cudaFree(parents);
    if (dev_error) cudaFree(dev_error);
    *this = ScanWorkspace();
}

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
    const uint32_t parents_bytes = ((uint32_t)a.nparents * 4u + 127u) & ~127u;
    int ctas = std::max(1, std::min(o.scan_ctas_per_sm, 4));
    ScanKParams p{};
    if (int rc = pick_geometry(a.n, S, parents_bytes, lim.smem_optin, ctas, p.g)) return rc;
    if (p.g.num_tiles > ws.tile_state_cap) return fail(CC_ERR_ARG, "scan workspace too small");
    // A fresh epoch makes every descriptor of earlier launches read as "invalid" without a memset.
    ws.epoch = (ws.epoch + 1) & kEpochMask;
    if (ws.epoch == 0) {
        CC_CUDA(cudaMemsetAsync(ws.tile_state, 0, ws.tile_state_cap * sizeof(uint64_t), st));
        ws.epoch = 1;
    }
    p.body = a.body; p.n = a.n; p.index_base = a.index_base;
    p.s = a.s; p.c = a.c; p.cov_off = 8u * a.s; p.edge_off = 8u * a.s + 4u * a.c; p.O = 8u * a.s + 5u;
    p.child = a.child; p.nparents = a.nparents; p.parents = ws.parents;
    p.out = a.out_records; p.out_index = a.out_index; p.cap = a.cap;
    p.total_in = a.total_in; p.total_out = a.total_out;
    p.tile_state = ws.tile_state; p.tile_counter = ws.tile_counter;
    p.epoch = ws.epoch; p.err = ws.dev_error;

    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sm_count * ctas, p.g.num_tiles);
    // every launch draws exactly num_tiles + grid tickets from the workspace's counter
    p.ticket_base = ws.ticket_base;
    ws.ticket_base += p.g.num_tiles + grid;

    const size_t smem = kCtrlBytes + parents_bytes + (size_t)p.g.stages * p.g.stage_bytes;
    const bool aligned4 = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.body) & 3u) == 0);
    auto kern = aligned4 ? scan_novel_kernel<true> : scan_novel_kernel<false>;
    CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(p);
    count_launch();
    CC_CUDA(cudaGetLastError());
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
    ws.ticket_base += p.g.num_tiles + grid;
    const size_t smem = kCtrlBytes + (size_t)p.g.stages * p.g.stage_bytes;
    const bool aligned4 = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(dev_body) & 3u) == 0);
    auto kern = aligned4 ? decode_columns_kernel<true> : decode_columns_kernel<false>;
    CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(p);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
