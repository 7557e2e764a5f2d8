// sharded.cu -- one graph over several GPUs of ONE process, behind the C ABI (cc_open_sharded and friends).
//
// The reference's host is a single JVM that constructs one CortexGraph per file (S/utils/arguments/ArgumentHandler.java:271-274)
// and answers findRecord (S/utils/io/graph/cortex/CortexGraph.java:272-317) and the FindROIs scan
// (S/commands/discover/roi/FindROIs.java:52-64) from it.  Here the sorted record array is cut into k-mer-range shards
// (contiguous record slices, one per device); a cc_sharded handle owns the shards and hides them:
//   lookups   every device takes an equal slice of the batch and runs the three routed legs of lookup.cu -- route (owner +
//             P2P store into the owner's inbox), search (bucket-line index of the shard), gather (P2P pull of the results) --
//             over plain cudaMalloc memory made visible with cudaDeviceEnablePeerAccess.  The cross-device barriers between
//             the legs are CUDA events (every device's stream waits for the leg of all others); no collective library.
//   scan      records are independent: every device scans its own slice; the per-shard counts become exclusive offsets and
//             the shards' novel records land at those offsets, so the output is the globally sorted list the single-GPU
//             scan produces (the ROI file must stay sorted: S/utils/stoppingrules/BubbleOpeningStopper.java:17).
// The same device id may be listed more than once (several shards on one GPU): that is how the logic is tested on a
// one-GPU box.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "cc_internal.hpp"

struct cc_sharded {
    int world = 0;
    std::vector<int> dev;
    std::vector<cc_graph *> shard;
    std::vector<void *> owned_body;           // device allocations made by cc_open_sharded (null when wrapping caller memory)
    std::vector<uint64_t> first;              // first global record index of every shard
    cc::Header h;                             // the whole graph's header (num_records = all shards)
    uint32_t kw = 0;
    bool boundaries_checked = false;
    bool replicated = false;                  // every device holds the WHOLE graph (shard[r] = full copy, first = 0); lookups need no exchange
    std::vector<cc_graph *> scan_view;        // replicated only: device r's view of records [n r / world, n (r+1) / world) of its copy, for the scan
    void *map_base = nullptr;
    uint64_t map_len = 0;
    // ---- routed-lookup state, per rank (allocated on first use, grown on demand)
    struct Rank {
        void *inbox = nullptr, *res = nullptr, *route_state = nullptr;
        uint64_t *counts_in = nullptr, *sent = nullptr, *splitters = nullptr, *shard_first = nullptr;
        uint64_t *q_words = nullptr;          // staging of the host-buffer entry points
        uint8_t *q_flags = nullptr, *q_ascii = nullptr;
        int64_t *q_out = nullptr;
        uint64_t stage_cap = 0, ascii_cap = 0;
        cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // route / search / gather of the current chunk
        cudaEvent_t t[4] = {nullptr, nullptr, nullptr, nullptr};
        uint64_t *scan_count = nullptr;
    };
    std::vector<Rank> rank;
    uint64_t cap = 0, max_batch = 0;          // segment capacity / largest chunk per rank the buffers were sized for
    std::vector<std::vector<void *>> p_inbox, p_counts, p_res;   // per rank: the peer pointer tables of the legs
    cc_sharded_stats stats{};
};

namespace cc {
namespace {

struct DevGuard {
    int prev = -1;
    DevGuard() { cudaGetDevice(&prev); }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void free_rank(cc_sharded::Rank &r) {
    void *ptrs[] = {r.inbox, r.res, r.route_state, r.counts_in, r.sent, r.splitters, r.shard_first, r.q_words, r.q_flags, r.q_ascii, r.q_out, r.scan_count};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t &e : r.ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t &e : r.t) if (e) cudaEventDestroy(e);
    r = cc_sharded::Rank();
}

void destroy_sharded(cc_sharded *sh) {
    if (!sh) return;
    DevGuard guard;
    for (int r = 0; r < (int)sh->rank.size(); ++r) {
        cudaSetDevice(sh->dev[r]);
        if (r < (int)sh->shard.size() && sh->shard[r] && sh->shard[r]->stream) cudaStreamSynchronize(sh->shard[r]->stream);
        free_rank(sh->rank[r]);
    }
    for (cc_graph *g : sh->scan_view) if (g) cc_dispose(g);
    for (cc_graph *g : sh->shard) if (g) cc_dispose(g);
    for (int r = 0; r < (int)sh->owned_body.size(); ++r)
        if (sh->owned_body[r]) { cudaSetDevice(sh->dev[r]); cudaFree(sh->owned_body[r]); }
    if (sh->map_base) munmap(sh->map_base, sh->map_len);
    cudaGetLastError();
    delete sh;
}

int enable_peers(const std::vector<int> &dev) {
    for (size_t i = 0; i < dev.size(); ++i) {
        for (size_t j = 0; j < dev.size(); ++j) {
            if (dev[i] == dev[j]) continue;
            int can = 0;
            CC_CUDA(cudaDeviceCanAccessPeer(&can, dev[i], dev[j]));
            if (!can) return fail(CC_ERR_CUDA, "device %d cannot access device %d (the sharded lookup needs peer access over NVLink)", dev[i], dev[j]);
            CC_CUDA(cudaSetDevice(dev[i]));
            const cudaError_t e = cudaDeviceEnablePeerAccess(dev[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
    }
    return CC_OK;
}

int check_devices(const int *devices, int ndev) {
    if (!devices || ndev < 1 || ndev > 64) return fail(CC_ERR_ARG, "device list must hold 1..64 entries");
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CC_ERR_CUDA, "no CUDA device available (%s); libcorticall_cuda has no CPU fallback", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    for (int i = 0; i < ndev; ++i)
        if (devices[i] < 0 || devices[i] >= n) return fail(CC_ERR_ARG, "device %d out of range (0..%d)", devices[i], n - 1);
    return CC_OK;
}

// Wraps per-device slices into shard handles, builds their lookup indices lazily (first lookup).
int finish_sharded(cc_sharded *sh, const void *const *bodies, const uint64_t *counts) {
    const Header &h = sh->h;
    sh->kw = (2 * h.k + 31) / 32;
    if (int rc = enable_peers(sh->dev)) return rc;
    uint64_t at = 0;
    sh->shard.assign(sh->world, nullptr);
    sh->first.assign(sh->world, 0);
    sh->rank.assign(sh->world, cc_sharded::Rank());
    if (sh->replicated) {
        // bodies[r] = device r's copy of all records; counts[r] = the slice of it device r scans
        sh->scan_view.assign(sh->world, nullptr);
        for (int r = 0; r < sh->world; ++r) {
            if (int rc = cc_open_device(bodies[r], h.k, h.s, h.c, h.num_records, 0, sh->dev[r], &sh->shard[r])) return rc;
            if (int rc = cc_open_device(static_cast<const uint8_t *>(bodies[r]) + at * h.record_size, h.k, h.s, h.c, counts[r], at, sh->dev[r], &sh->scan_view[r])) return rc;
            sh->shard[r]->h.colors = h.colors.size() == h.c ? h.colors : sh->shard[r]->h.colors;
            sh->shard[r]->h.version = 6;
            at += counts[r];
        }
        return CC_OK;
    }
    for (int r = 0; r < sh->world; ++r) {
        sh->first[r] = at;
        if (int rc = cc_open_device(bodies[r], h.k, h.s, h.c, counts[r], at, sh->dev[r], &sh->shard[r])) return rc;
        // the shard answers header queries like the whole graph would (colour names, cleaning flags); only its record count is its own
        sh->shard[r]->h.colors = h.colors.size() == h.c ? h.colors : sh->shard[r]->h.colors;
        sh->shard[r]->h.version = 6;
        at += counts[r];
    }
    return CC_OK;
}

int ensure_indices(cc_sharded *sh) {
    for (int r = 0; r < sh->world; ++r) {
        cc_graph *g = sh->shard[r];
        if (!g->index.built) {
            CC_CUDA(cudaSetDevice(sh->dev[r]));
            if (int rc = build_index(g, 0)) return rc;
        }
        if (!g->index.sorted)
            return fail(CC_ERR_UNSORTED, "Records are not sorted (record %llu sorts before its predecessor)",
                        (unsigned long long)(sh->first[r] + g->index.unsorted_at));
    }
    // the order must also hold ACROSS the shard boundaries (checked once; replicas hold the whole array)
    if (!sh->boundaries_checked && !sh->replicated) {
        const uint32_t s = sh->h.s;
        std::vector<uint64_t> prev_last(s), cur(s);
        bool have_prev = false;
        for (int r = 0; r < sh->world; ++r) {
            cc_graph *g = sh->shard[r];
            const uint64_t n = g->h.num_records;
            if (n == 0) continue;
            CC_CUDA(cudaSetDevice(sh->dev[r]));
            CC_CUDA(cudaMemcpy(cur.data(), g->index.keys, s * 8, cudaMemcpyDeviceToHost));
            if (have_prev && std::lexicographical_compare(cur.begin(), cur.end(), prev_last.begin(), prev_last.end()))
                return fail(CC_ERR_UNSORTED, "Records are not sorted (record %llu sorts before its predecessor)", (unsigned long long)sh->first[r]);
            CC_CUDA(cudaMemcpy(prev_last.data(), g->index.keys + (n - 1) * s, s * 8, cudaMemcpyDeviceToHost));
            have_prev = true;
        }
        sh->boundaries_checked = true;
    }
    return CC_OK;
}

// Sizes the exchange buffers for chunks of up to `batch` queries per rank.  Segments hold the balanced share plus a quarter
// (a worst-case layout costs the search leg TLB reach); a chunk that overflows one is re-run in pieces.
int ensure_exchange(cc_sharded *sh, uint64_t batch) {
    if (batch <= sh->max_batch && sh->max_batch) return CC_OK;
    const int world = sh->world;
    const uint32_t s = sh->h.s, kw = sh->kw;
    batch = std::max<uint64_t>(batch, 1024);
    uint64_t cap = world == 1 ? batch : batch / world + batch / (4 * world) + 4096;
    cap = (std::min(cap, batch) + 3) & ~3ull;
    // splitters: the first key of shards 1.. (an empty shard gets the largest key: nothing routes to it)
    std::vector<uint64_t> spl((size_t)std::max(world - 1, 1) * s, ~0ull);
    for (int r = 1; r < world; ++r) {
        if (sh->shard[r]->h.num_records == 0) continue;
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        CC_CUDA(cudaMemcpy(&spl[(size_t)(r - 1) * s], sh->shard[r]->index.keys, s * 8, cudaMemcpyDeviceToHost));
    }
    // an empty shard in the middle must not break the ascending order of the splitters: give it the next shard's first key
    for (int r = world - 2; r >= 1; --r) {
        if (sh->shard[r]->h.num_records == 0) memcpy(&spl[(size_t)(r - 1) * s], &spl[(size_t)r * s], s * 8);
    }
    uint64_t state_bytes = route_state_size(batch, world);
    for (int r = 0; r < world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        cc_sharded::Rank &k = sh->rank[r];
        CC_CUDA(cudaStreamSynchronize(sh->shard[r]->stream));
        void **grow[] = {&k.inbox, &k.res, &k.route_state};
        for (void **p : grow) if (*p) { cudaFree(*p); *p = nullptr; }
        CC_CUDA(cudaMalloc(&k.inbox, (uint64_t)world * cap * kw * 4 + 64));
        CC_CUDA(cudaMalloc(&k.res, (uint64_t)world * cap * 4 + 64));
        CC_CUDA(cudaMalloc(&k.route_state, state_bytes + 64));
        if (!k.counts_in) {
            CC_CUDA(cudaMalloc(&k.counts_in, 8 * 64));
            CC_CUDA(cudaMalloc(&k.sent, 8 * 64));
            CC_CUDA(cudaMalloc(&k.splitters, spl.size() * 8));
            CC_CUDA(cudaMalloc(&k.shard_first, 8 * 64));
            CC_CUDA(cudaMalloc(&k.scan_count, 64));
            for (cudaEvent_t &e : k.ev) CC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            for (cudaEvent_t &e : k.t) CC_CUDA(cudaEventCreate(&e));
        }
        CC_CUDA(cudaMemset(k.counts_in, 0, 8 * 64));
        CC_CUDA(cudaMemcpy(k.splitters, spl.data(), spl.size() * 8, cudaMemcpyHostToDevice));
        CC_CUDA(cudaMemcpy(k.shard_first, sh->first.data(), 8 * world, cudaMemcpyHostToDevice));
    }
    sh->p_inbox.assign(world, std::vector<void *>(world));
    sh->p_counts.assign(world, std::vector<void *>(world));
    sh->p_res.assign(world, std::vector<void *>(world));
    for (int r = 0; r < world; ++r) {
        for (int o = 0; o < world; ++o) {
            sh->p_inbox[r][o] = sh->rank[o].inbox;                                                    // owner o's [world][cap][kw] block
            sh->p_counts[r][o] = sh->rank[o].counts_in;
            sh->p_res[r][o] = static_cast<uint8_t *>(sh->rank[o].res) + (uint64_t)r * cap * 4;        // segment (owner o, source r)
        }
    }
    sh->cap = cap;
    sh->max_batch = batch;
    return CC_OK;
}

// Every rank's stream waits for event `which` of all ranks (the cross-device barrier between two legs).
int barrier_on(cc_sharded *sh, int which) {
    for (int r = 0; r < sh->world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        for (int o = 0; o < sh->world; ++o)
            if (o != r) CC_CUDA(cudaStreamWaitEvent(sh->shard[r]->stream, sh->rank[o].ev[which], 0));
    }
    return CC_OK;
}

// One chunk: nq[r] <= max_batch canonical packed queries resident on every rank's device -> out[r].  Asynchronous on the
// shard streams; the caller synchronises.  `timed`: record the leg boundaries for cc_sharded_stats.
int routed_chunk(cc_sharded *sh, const uint64_t *const *words, const uint8_t *const *flags, const uint64_t *nq, int64_t *const *out, bool timed) {
    const int world = sh->world;
    const uint32_t k = sh->h.k;
    for (int r = 0; r < world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        cc_sharded::Rank &rk = sh->rank[r];
        cudaStream_t st = sh->shard[r]->stream;
        if (timed) CC_CUDA(cudaEventRecord(rk.t[0], st));
        if (int rc = launch_route(words[r], flags ? flags[r] : nullptr, nq[r], k, rk.splitters, world, r, sh->cap, sh->p_inbox[r].data(),
                                  sh->p_counts[r].data(), rk.route_state, sh->max_batch, rk.sent, st)) return rc;
        CC_CUDA(cudaEventRecord(rk.ev[0], st));
        if (timed) CC_CUDA(cudaEventRecord(rk.t[1], st));
    }
    if (int rc = barrier_on(sh, 0)) return rc;
    for (int r = 0; r < world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        cc_sharded::Rank &rk = sh->rank[r];
        cudaStream_t st = sh->shard[r]->stream;
        if (int rc = launch_find_routed(sh->shard[r], rk.inbox, rk.counts_in, world, 1, sh->cap, rk.res, st)) return rc;
        CC_CUDA(cudaEventRecord(rk.ev[1], st));
        if (timed) CC_CUDA(cudaEventRecord(rk.t[2], st));
    }
    if (int rc = barrier_on(sh, 1)) return rc;
    for (int r = 0; r < world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        cc_sharded::Rank &rk = sh->rank[r];
        cudaStream_t st = sh->shard[r]->stream;
        if (int rc = launch_gather_routed(sh->p_res[r].data(), rk.route_state, sh->max_batch, nq[r], rk.shard_first, world, sh->cap, out[r], st)) return rc;
        CC_CUDA(cudaEventRecord(rk.ev[2], st));
        if (timed) CC_CUDA(cudaEventRecord(rk.t[3], st));
    }
    // No barrier behind the gather: the next chunk's route follows this rank's gather in stream order, and that gather
    // waited for the search of every owner (nobody still reads an inbox); the next search waits for the next route of every
    // rank, hence for every rank's gather (nobody still pulls results).
    return CC_OK;
}

int sync_all(cc_sharded *sh) {
    for (int r = 0; r < sh->world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        const cudaError_t e = cudaStreamSynchronize(sh->shard[r]->stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
    }
    return CC_OK;
}

// Did a segment overflow in the chunk just finished?  (sent[o] on rank r = keys r routed to o.)
int overflowed(cc_sharded *sh, bool &over) {
    over = false;
    if (sh->cap >= sh->max_batch) return CC_OK;
    std::vector<uint64_t> sent(sh->world);
    for (int r = 0; r < sh->world; ++r) {
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        CC_CUDA(cudaMemcpy(sent.data(), sh->rank[r].sent, 8 * sh->world, cudaMemcpyDeviceToHost));
        for (uint64_t v : sent) over |= v > sh->cap;
    }
    return CC_OK;
}

// All chunks of a device-resident batch.  A chunk that overflowed a segment (a batch far from balanced) is repeated in
// pieces of `cap` queries, which cannot overflow.
int routed_all(cc_sharded *sh, const uint64_t *const *words, const uint8_t *const *flags, const uint64_t *nq, int64_t *const *out, uint64_t piece) {
    const int world = sh->world;
    const uint32_t s = sh->h.s;
    uint64_t longest = 0;
    for (int r = 0; r < world; ++r) longest = std::max(longest, nq[r]);
    std::vector<const uint64_t *> w(world);
    std::vector<const uint8_t *> f(world);
    std::vector<int64_t *> o(world);
    std::vector<uint64_t> m(world);
    for (uint64_t at = 0; at < longest; at += piece) {
        for (int r = 0; r < world; ++r) {
            const uint64_t lo = std::min(at, nq[r]), hi = std::min(at + piece, nq[r]);
            w[r] = words[r] + lo * s;
            f[r] = flags && flags[r] ? flags[r] + lo : nullptr;
            o[r] = out[r] + lo;
            m[r] = hi - lo;
        }
        const bool timed = at == 0 && piece == sh->max_batch;
        if (int rc = routed_chunk(sh, w.data(), flags ? f.data() : nullptr, m.data(), o.data(), timed)) return rc;
        if (sh->cap < piece) {
            if (int rc = sync_all(sh)) return rc;
            bool over = false;
            if (int rc = overflowed(sh, over)) return rc;
            if (over) {
                sh->stats.overflow_retries++;
                if (int rc = routed_all(sh, w.data(), flags ? f.data() : nullptr, m.data(), o.data(), sh->cap)) return rc;
            }
        }
    }
    return CC_OK;
}

// Replicated placement: every device answers its own queries from its own copy -- no exchange, global indices directly.
int local_all(cc_sharded *sh, const uint64_t *const *words, const uint8_t *const *flags, const uint64_t *nq, int64_t *const *out) {
    for (int r = 0; r < sh->world; ++r) {
        if (!nq[r]) continue;
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        cc_graph *g = sh->shard[r];
        if (int rc = launch_find_packed(g, words[r], flags && flags[r] ? flags[r] : nullptr, nq[r], out[r], CC_ALGO_AUTO, g->stream)) return rc;
    }
    return CC_OK;
}

void collect_stats(cc_sharded *sh) {
    if (sh->replicated) return;
    float route = 0, search = 0, gather = 0, total = 0;
    for (int r = 0; r < sh->world; ++r) {
        cudaSetDevice(sh->dev[r]);
        float a = 0, b = 0, c = 0, d = 0;
        cudaEventElapsedTime(&a, sh->rank[r].t[0], sh->rank[r].t[1]);
        cudaEventElapsedTime(&b, sh->rank[r].t[1], sh->rank[r].t[2]);
        cudaEventElapsedTime(&c, sh->rank[r].t[2], sh->rank[r].t[3]);
        cudaEventElapsedTime(&d, sh->rank[r].t[0], sh->rank[r].t[3]);
        route = std::max(route, a); search = std::max(search, b); gather = std::max(gather, c); total = std::max(total, d);
    }
    cudaGetLastError();
    sh->stats.route_ms = route; sh->stats.search_ms = search; sh->stats.gather_ms = gather; sh->stats.chunk_ms = total;
}

uint64_t pick_batch(const cc_sharded *sh, uint64_t per_rank) {
    const uint64_t limit = 1ull << 27;            // 134 M queries per rank and chunk: 2 GB of packed queries
    return std::max<uint64_t>(std::min(per_rank, limit), 1);
}

}  // namespace
}  // namespace cc

using namespace cc;

extern "C" {

int cc_open_sharded_device(const void *const *dev_bodies, const uint64_t *counts, uint32_t k, uint32_t s, uint32_t c, const int *devices,
                           int ndev, cc_sharded **out) {
    if (!out || !dev_bodies || !counts) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    if (int rc = check_devices(devices, ndev)) return rc;
    if (k == 0 || s != (k + 31) / 32) return fail(CC_ERR_ARG, "kmer_bits %u does not match kmer_size %u", s, k);
    DevGuard guard;
    std::unique_ptr<cc_sharded, void (*)(cc_sharded *)> sh(new cc_sharded(), destroy_sharded);
    sh->world = ndev;
    sh->dev.assign(devices, devices + ndev);
    sh->h.version = 6; sh->h.k = k; sh->h.s = s; sh->h.c = c;
    sh->h.record_size = 8ull * s + 5ull * c;
    sh->h.num_records = 0;
    for (int r = 0; r < ndev; ++r) sh->h.num_records += counts[r];
    sh->owned_body.assign(ndev, nullptr);
    if (int rc = finish_sharded(sh.get(), dev_bodies, counts)) return rc;
    *out = sh.release();
    return CC_OK;
}

static int open_sharded_image(std::unique_ptr<cc_sharded, void (*)(cc_sharded *)> &sh, const uint8_t *image, uint64_t size, const char *path,
                              const int *devices, int ndev, int placement, cc_sharded **out) {
    if (placement < CC_PLACE_RANGE || placement > CC_PLACE_AUTO) return fail(CC_ERR_ARG, "unknown placement %d", placement);
    if (int rc = parse_header(image, size, size, path, sh->h)) return rc;
    if (sh->h.s != (sh->h.k + 31) / 32) return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': kmer_bits %u does not match kmer_size %u", path, sh->h.s, sh->h.k);
    if (int rc = check_devices(devices, ndev)) return rc;
    DevGuard guard;
    sh->world = ndev;
    sh->dev.assign(devices, devices + ndev);
    sh->owned_body.assign(ndev, nullptr);
    const uint64_t n = sh->h.num_records, S = sh->h.record_size;
    std::vector<const void *> bodies(ndev);
    std::vector<uint64_t> counts(ndev);
    if (placement == CC_PLACE_AUTO) {
        // replicas when a lookup-ready copy (records + key column + bucket lines, <= 32 bytes per record) takes at most a quarter of
        // every device's free memory: lookups then need no exchange and scale with the device count; k-mer ranges otherwise
        const uint64_t need = n * (S + 8ull * sh->h.s + 32) + (64ull << 20);
        placement = CC_PLACE_REPLICATE;
        for (int r = 0; r < ndev && placement == CC_PLACE_REPLICATE; ++r) {
            size_t free_b = 0, total_b = 0;
            CC_CUDA(cudaSetDevice(devices[r]));
            CC_CUDA(cudaMemGetInfo(&free_b, &total_b));
            if (need > free_b / 4) placement = CC_PLACE_RANGE;
        }
    }
    sh->replicated = placement == CC_PLACE_REPLICATE && ndev > 1;
    if (sh->replicated) {
        for (int r = 0; r < ndev; ++r) {
            counts[r] = n * (r + 1) / ndev - n * r / ndev;
            CC_CUDA(cudaSetDevice(devices[r]));
            CC_CUDA(cudaMalloc(&sh->owned_body[r], n * S + 256));
            CC_CUDA(cudaMemcpyAsync(sh->owned_body[r], image + sh->h.data_offset, n * S, cudaMemcpyHostToDevice, nullptr));
            CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(sh->owned_body[r]) + n * S, 0, 256, nullptr));
            bodies[r] = sh->owned_body[r];
        }
        for (int r = 0; r < ndev; ++r) {
            CC_CUDA(cudaSetDevice(devices[r]));
            CC_CUDA(cudaDeviceSynchronize());
        }
        if (int rc = finish_sharded(sh.get(), bodies.data(), counts.data())) return rc;
        *out = sh.release();
        return CC_OK;
    }
    for (int r = 0; r < ndev; ++r) {
        const uint64_t lo = n * r / ndev, hi = n * (r + 1) / ndev;
        counts[r] = hi - lo;
        CC_CUDA(cudaSetDevice(devices[r]));
        CC_CUDA(cudaMalloc(&sh->owned_body[r], (hi - lo) * S + 256));
        CC_CUDA(cudaMemcpy(sh->owned_body[r], image + sh->h.data_offset + lo * S, (hi - lo) * S, cudaMemcpyHostToDevice));
        CC_CUDA(cudaMemset(static_cast<uint8_t *>(sh->owned_body[r]) + (hi - lo) * S, 0, 256));
        bodies[r] = sh->owned_body[r];
    }
    if (int rc = finish_sharded(sh.get(), bodies.data(), counts.data())) return rc;
    *out = sh.release();
    return CC_OK;
}

int cc_open_sharded(const char *path, const int *devices, int ndev, cc_sharded **out) {
    return cc_open_sharded_placed(path, devices, ndev, CC_PLACE_RANGE, out);
}

int cc_open_sharded_placed(const char *path, const int *devices, int ndev, int placement, cc_sharded **out) {
    if (!path || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<cc_sharded, void (*)(cc_sharded *)> sh(new cc_sharded(), destroy_sharded);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(CC_ERR_IO, "Cortex graph file '%s' not found: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': %s", path, strerror(errno)); }
    const uint64_t size = (uint64_t)sb.st_size;
    if (size == 0) { close(fd); return fail(CC_ERR_NOT_CORTEX, "The file '%s' does not appear to be a Cortex graph", path); }
    void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': mmap: %s", path, strerror(errno));
    sh->map_base = m;
    sh->map_len = size;
    return open_sharded_image(sh, static_cast<const uint8_t *>(m), size, path, devices, ndev, placement, out);
}

int cc_open_sharded_memory(const void *file_image, uint64_t size, const int *devices, int ndev, cc_sharded **out) {
    return cc_open_sharded_memory_placed(file_image, size, devices, ndev, CC_PLACE_RANGE, out);
}

int cc_open_sharded_memory_placed(const void *file_image, uint64_t size, const int *devices, int ndev, int placement, cc_sharded **out) {
    if (!file_image || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<cc_sharded, void (*)(cc_sharded *)> sh(new cc_sharded(), destroy_sharded);
    return open_sharded_image(sh, static_cast<const uint8_t *>(file_image), size, "<memory>", devices, ndev, placement, out);
}

void cc_dispose_sharded(cc_sharded *sh) { destroy_sharded(sh); }

int cc_sharded_info(const cc_sharded *sh, int *ndev, uint64_t *num_records, uint32_t *kmer_size, uint32_t *num_colors) {
    if (!sh) return fail(CC_ERR_ARG, "null graph");
    if (ndev) *ndev = sh->world;
    if (num_records) *num_records = sh->h.num_records;
    if (kmer_size) *kmer_size = sh->h.k;
    if (num_colors) *num_colors = sh->h.c;
    return CC_OK;
}

int cc_sharded_placement(const cc_sharded *sh, int *placement) {
    if (!sh || !placement) return fail(CC_ERR_ARG, "null argument");
    *placement = sh->replicated ? CC_PLACE_REPLICATE : CC_PLACE_RANGE;
    return CC_OK;
}

int cc_sharded_shard(const cc_sharded *sh, int rank, cc_graph **shard, int *device, uint64_t *first_index) {
    if (!sh || rank < 0 || rank >= sh->world) return fail(CC_ERR_ARG, "shard %d out of range", rank);
    if (shard) *shard = sh->shard[rank];
    if (device) *device = sh->dev[rank];
    if (first_index) *first_index = sh->first[rank];
    return CC_OK;
}

int cc_sharded_last_stats(const cc_sharded *sh, cc_sharded_stats *out) {
    if (!sh || !out) return fail(CC_ERR_ARG, "null argument");
    *out = sh->stats;
    return CC_OK;
}

int cc_find_packed_sharded_dev(cc_sharded *sh, const uint64_t *const *dev_words, const uint8_t *const *dev_flags, const uint64_t *nq,
                               int64_t *const *dev_out) {
    if (!sh || !dev_words || !nq || !dev_out) return fail(CC_ERR_ARG, "null argument");
    DevGuard guard;
    uint64_t longest = 0;
    for (int r = 0; r < sh->world; ++r) {
        if (nq[r] && (!dev_words[r] || !dev_out[r])) return fail(CC_ERR_ARG, "null query or result buffer for device slot %d", r);
        longest = std::max(longest, nq[r]);
    }
    if (longest == 0) return CC_OK;
    if (int rc = ensure_indices(sh)) return rc;
    if (!sh->replicated)
        if (int rc = ensure_exchange(sh, pick_batch(sh, longest))) return rc;
    sh->stats = cc_sharded_stats{};
    const uint64_t launches0 = g_launches.load();
    if (sh->replicated) { if (int rc = local_all(sh, dev_words, dev_flags, nq, dev_out)) return rc; }
    else if (int rc = routed_all(sh, dev_words, dev_flags, nq, dev_out, sh->max_batch)) return rc;
    if (int rc = sync_all(sh)) return rc;
    collect_stats(sh);
    sh->stats.launches = (uint32_t)(g_launches.load() - launches0);
    return CC_OK;
}

}  // extern "C"

namespace {

// Host-buffer lookups: rank r takes queries [nq*r/world, nq*(r+1)/world); the slices are copied to their devices, packed
// there when they arrive as ASCII (K3), looked up through the routed legs, and the indices copied back.
enum class HostIn { Packed, Ascii, Windows };

int ensure_staging(cc_sharded *sh, uint64_t per_rank, uint64_t ascii_bytes) {
    for (int r = 0; r < sh->world; ++r) {
        cc_sharded::Rank &k = sh->rank[r];
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        if (per_rank > k.stage_cap) {
            CC_CUDA(cudaStreamSynchronize(sh->shard[r]->stream));
            void **ps[] = {reinterpret_cast<void **>(&k.q_words), reinterpret_cast<void **>(&k.q_flags), reinterpret_cast<void **>(&k.q_out)};
            for (void **p : ps) if (*p) { cudaFree(*p); *p = nullptr; }
            k.stage_cap = 0;
            CC_CUDA(cudaMalloc(&k.q_words, per_rank * sh->h.s * 8 + 64));
            CC_CUDA(cudaMalloc(&k.q_flags, per_rank + 64));
            CC_CUDA(cudaMalloc(&k.q_out, per_rank * 8 + 64));
            k.stage_cap = per_rank;
        }
        if (ascii_bytes > k.ascii_cap) {
            CC_CUDA(cudaStreamSynchronize(sh->shard[r]->stream));
            if (k.q_ascii) { cudaFree(k.q_ascii); k.q_ascii = nullptr; }
            k.ascii_cap = 0;
            CC_CUDA(cudaMalloc(&k.q_ascii, ascii_bytes + 64));
            k.ascii_cap = ascii_bytes;
        }
    }
    return CC_OK;
}

int host_lookup(cc_sharded *sh, HostIn kind, const void *in, const uint8_t *flags, uint64_t nq, int64_t *out) {
    cc::DevGuard guard;
    const int world = sh->world;
    const uint32_t s = sh->h.s, k = sh->h.k;
    if (int rc = ensure_indices(sh)) return rc;
    // super-chunks bound the staging memory; inside one, routed_all cuts the exchange chunks
    const uint64_t super = (uint64_t)world << 26;
    sh->stats = cc_sharded_stats{};
    const uint64_t launches0 = g_launches.load();
    for (uint64_t base = 0; base < nq; base += super) {
        const uint64_t m = std::min(super, nq - base);
        const uint64_t per = (m + world - 1) / world;
        const uint64_t ascii_bytes = kind == HostIn::Ascii ? per * k : kind == HostIn::Windows ? per + k : 0;
        if (int rc = ensure_staging(sh, per, ascii_bytes)) return rc;
        if (!sh->replicated)
            if (int rc = ensure_exchange(sh, pick_batch(sh, per))) return rc;
        std::vector<const uint64_t *> w(world);
        std::vector<const uint8_t *> f(world);
        std::vector<int64_t *> o(world);
        std::vector<uint64_t> cnt(world), lo(world);
        for (int r = 0; r < world; ++r) {
            lo[r] = base + m * r / world;
            cnt[r] = base + m * (r + 1) / world - lo[r];
            cc_sharded::Rank &rk = sh->rank[r];
            CC_CUDA(cudaSetDevice(sh->dev[r]));
            cudaStream_t st = sh->shard[r]->stream;
            if (kind == HostIn::Ascii) {
                CC_CUDA(cudaMemcpyAsync(rk.q_ascii, static_cast<const uint8_t *>(in) + lo[r] * k, cnt[r] * k, cudaMemcpyHostToDevice, st));
                if (cnt[r])
                    if (int rc = launch_pack_windows(rk.q_ascii, cnt[r] * k, k, rk.q_words, rk.q_flags, k, cnt[r], st)) return rc;
            } else if (kind == HostIn::Windows) {       // windows [lo, lo + cnt) of the sequence need bytes [lo, lo + cnt + k - 1)
                if (cnt[r]) {
                    CC_CUDA(cudaMemcpyAsync(rk.q_ascii, static_cast<const uint8_t *>(in) + lo[r], cnt[r] + k - 1, cudaMemcpyHostToDevice, st));
                    if (int rc = launch_pack_windows(rk.q_ascii, cnt[r] + k - 1, k, rk.q_words, rk.q_flags, 1, cnt[r], st)) return rc;
                }
            } else {
                CC_CUDA(cudaMemcpyAsync(rk.q_words, static_cast<const uint64_t *>(in) + lo[r] * s, cnt[r] * s * 8, cudaMemcpyHostToDevice, st));
                if (flags) CC_CUDA(cudaMemcpyAsync(rk.q_flags, flags + lo[r], cnt[r], cudaMemcpyHostToDevice, st));
            }
            sh->stats.h2d_bytes += kind == HostIn::Ascii ? cnt[r] * k : kind == HostIn::Windows ? (cnt[r] ? cnt[r] + k - 1 : 0) : cnt[r] * s * 8 + (flags ? cnt[r] : 0);
            w[r] = rk.q_words;
            f[r] = (kind != HostIn::Packed || flags) ? rk.q_flags : nullptr;
            o[r] = rk.q_out;
        }
        if (sh->replicated) { if (int rc = local_all(sh, w.data(), (kind != HostIn::Packed || flags) ? f.data() : nullptr, cnt.data(), o.data())) return rc; }
        else if (int rc = routed_all(sh, w.data(), (kind != HostIn::Packed || flags) ? f.data() : nullptr, cnt.data(), o.data(), sh->max_batch)) return rc;
        for (int r = 0; r < world; ++r) {
            CC_CUDA(cudaSetDevice(sh->dev[r]));
            CC_CUDA(cudaMemcpyAsync(out + lo[r], sh->rank[r].q_out, cnt[r] * 8, cudaMemcpyDeviceToHost, sh->shard[r]->stream));
            sh->stats.d2h_bytes += cnt[r] * 8;
        }
        if (int rc = sync_all(sh)) return rc;
        if (base == 0) collect_stats(sh);
    }
    sh->stats.launches = (uint32_t)(g_launches.load() - launches0);
    return CC_OK;
}

}  // namespace

extern "C" {

int cc_find_packed_sharded(cc_sharded *sh, const uint64_t *words, const uint8_t *flags, uint64_t nq, int64_t *out_index) {
    if (!sh || (nq && (!words || !out_index))) return fail(CC_ERR_ARG, "null argument");
    if (nq == 0) return CC_OK;
    return host_lookup(sh, HostIn::Packed, words, flags, nq, out_index);
}

int cc_find_ascii_sharded(cc_sharded *sh, const uint8_t *kmers, uint64_t nq, int64_t *out_index) {
    if (!sh || (nq && (!kmers || !out_index))) return fail(CC_ERR_ARG, "null argument");
    if (nq == 0) return CC_OK;
    return host_lookup(sh, HostIn::Ascii, kmers, nullptr, nq, out_index);
}

int cc_find_windows_sharded(cc_sharded *sh, const uint8_t *seq, uint64_t len, int64_t *out_index) {
    if (!sh) return fail(CC_ERR_ARG, "null graph");
    if (len < sh->h.k) return CC_OK;
    if (!seq || !out_index) return fail(CC_ERR_ARG, "null argument");
    return host_lookup(sh, HostIn::Windows, seq, nullptr, len - sh->h.k + 1, out_index);
}

// FindROIs over the shards: every device scans its slice; the concatenation in shard order is the single-GPU output.
int cc_find_novel_sharded(cc_sharded *sh, int32_t child, const int32_t *parents, int nparents, void *out_records, uint64_t *out_index,
                          uint64_t cap, uint64_t *out_count) {
    if (!sh || !out_count) return fail(CC_ERR_ARG, "null argument");
    cc::DevGuard guard;
    const int world = sh->world;
    const uint64_t O = 8ull * sh->h.s + 5;
    std::vector<void *> stage(world, nullptr), stage_idx(world, nullptr);
    std::vector<uint64_t> dcap(world), count(world, 0);
    struct Free {
        cc_sharded *sh; std::vector<void *> &a, &b;
        ~Free() {
            for (int r = 0; r < sh->world; ++r) {
                cudaSetDevice(sh->dev[r]);
                if (a[r]) cudaFreeAsync(a[r], sh->shard[r]->stream);
                if (b[r]) cudaFreeAsync(b[r], sh->shard[r]->stream);
            }
        }
    } cleanup{sh, stage, stage_idx};
    const uint64_t launches0 = g_launches.load();
    for (int r = 0; r < world; ++r) {
        cc_sharded::Rank &rk = sh->rank[r];
        CC_CUDA(cudaSetDevice(sh->dev[r]));
        if (!rk.scan_count) CC_CUDA(cudaMalloc(&rk.scan_count, 64));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool again = false;
        for (int r = 0; r < world; ++r) {
            cc_graph *g = sh->replicated ? sh->scan_view[r] : sh->shard[r];
            const uint64_t n = g->h.num_records;
            if (attempt == 0) dcap[r] = std::min<uint64_t>(n, std::max<uint64_t>(65536, n / 32));
            else if (count[r] <= dcap[r]) continue;
            else dcap[r] = count[r];
            CC_CUDA(cudaSetDevice(sh->dev[r]));
            cudaStream_t st = g->stream;
            if (stage[r]) { cudaFreeAsync(stage[r], st); stage[r] = nullptr; }
            if (stage_idx[r]) { cudaFreeAsync(stage_idx[r], st); stage_idx[r] = nullptr; }
            CC_CUDA(cudaMallocAsync(&stage[r], dcap[r] * O + 64, st));
            if (out_index) CC_CUDA(cudaMallocAsync(&stage_idx[r], dcap[r] * 8 + 64, st));
            if (int rc = cc_find_novel_dev(g, child, parents, nparents, stage[r], static_cast<uint64_t *>(stage_idx[r]), dcap[r], sh->rank[r].scan_count, st)) return rc;
            CC_CUDA(cudaMemcpyAsync(&count[r], sh->rank[r].scan_count, 8, cudaMemcpyDeviceToHost, st));
        }
        if (int rc = sync_all(sh)) return rc;
        // only the part of a shard's output that fits the caller's cap is needed
        uint64_t before = 0;
        for (int r = 0; r < world; ++r) {
            const uint64_t room = out_records && cap > before ? cap - before : 0;
            if (std::min(count[r], room) > dcap[r]) again = true;
            before += count[r];
        }
        if (!again) break;
    }
    uint64_t total = 0;
    for (int r = 0; r < world; ++r) {
        const uint64_t room = out_records && cap > total ? cap - total : 0;
        const uint64_t take = std::min(count[r], room);
        if (take) {
            CC_CUDA(cudaSetDevice(sh->dev[r]));
            cudaStream_t st = sh->shard[r]->stream;
            CC_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(out_records) + total * O, stage[r], take * O, cudaMemcpyDeviceToHost, st));
            if (out_index) CC_CUDA(cudaMemcpyAsync(out_index + total, stage_idx[r], take * 8, cudaMemcpyDeviceToHost, st));
        }
        total += count[r];
    }
    if (int rc = sync_all(sh)) return rc;
    *out_count = total;
    sh->stats = cc_sharded_stats{};
    sh->stats.launches = (uint32_t)(g_launches.load() - launches0);
    return CC_OK;
}

int cc_write_roi_file_sharded(cc_sharded *sh, int32_t child, const int32_t *parents, int nparents, const char *out_path, uint64_t *out_count) {
    if (!sh || !out_path) return fail(CC_ERR_ARG, "null argument");
    if (child < 0 || (uint32_t)child >= sh->h.c) return fail(CC_ERR_ARG, "child colour %d out of range (graph has %u colours)", child, sh->h.c);
    const uint64_t O = 8ull * sh->h.s + 5;
    uint64_t cap = std::min<uint64_t>(sh->h.num_records, std::max<uint64_t>(65536, sh->h.num_records / 32)), total = 0;
    std::vector<uint8_t> recs(std::max<uint64_t>(cap, 1) * O);
    if (int rc = cc_find_novel_sharded(sh, child, parents, nparents, recs.data(), nullptr, cap, &total)) return rc;
    if (total > cap) {
        cap = total;
        recs.resize(cap * O);
        if (int rc = cc_find_novel_sharded(sh, child, parents, nparents, recs.data(), nullptr, cap, &total)) return rc;
    }
    const std::string name = sh->h.colors.size() > (size_t)child ? sh->h.colors[child].sample_name : std::to_string(child);
    const std::vector<uint8_t> hdr = make_roi_header(sh->h.k, sh->h.s, name);
    FILE *f = fopen(out_path, "wb");
    if (!f) return fail(CC_ERR_IO, "Cortex graph file '%s' not found: %s", out_path, strerror(errno));
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    if (total) ok = ok && fwrite(recs.data(), 1, total * O, f) == total * O;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(CC_ERR_IO, "Error while writing Cortex graph file '%s': %s", out_path, strerror(errno));
    if (out_count) *out_count = total;
    return CC_OK;
}

}  // extern "C"
