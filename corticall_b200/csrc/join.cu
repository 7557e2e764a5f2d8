// join.cu -- merged view of several sorted graphs: the record stream CortexCollection.next() produces and Join writes.
//
// Reference semantics reproduced (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
//   S/utils/io/graph/cortex/CortexCollection.java:34-62    colours are concatenated in graph order
//   S/utils/io/graph/cortex/CortexCollection.java:245-293  next(): lowest k-mer among the graphs' heads; every graph
//                                                          holding that k-mer contributes its coverage / edges to its
//                                                          own colours, the others stay 0; output ascending by k-mer
//   S/commands/utils/Join.java:23-57                        Join = write that stream with CortexGraphWriter
//
// B200 design.  A k-way merge is folded into two-way unions.  One union of A and B (sorted, duplicate-free key columns),
// tiled (round 2): the merged order is cut into tiles of T elements by one merge-path search per TILE
// (join_split_kernel; a B key is kept in the tile of its equal A key); a tile's inputs are then two CONTIGUOUS slices of
// the key columns, which a CTA copies into shared memory with coalesced loads.  Count
// pass (join_count_kernel: merge the key slices in shared memory, count the records they start) -> exclusive scan of the
// tile counts -> emit pass (join_emit_kernel: same merge; per output record the source index in A and in B, and the key
// column of the result when another union follows, leave in output order) -> compose pass (compose_kernel: gathers both
// source records into the wider output record through shared memory so the output leaves as aligned 16-byte stores;
// neighbouring output records come from neighbouring source records, so the gathers share their cache lines).  Linear in
// |A| + |B|; no sort; scratch comes from the stream-ordered pool.  The round-1 form (one global merge-path search per thread) is kept for inputs beyond
// 2^32 records and as a cross-check (option "join_tiled" = 0).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "cc_internal.hpp"
#include "device_utils.cuh"

namespace cc {

namespace {

constexpr int kJBlock = 256;
constexpr int kVT = 8;                         // merged elements per thread
constexpr int kTile = kJBlock * kVT;           // merged elements per block
constexpr int kComposeRecords = 128;           // output records per compose block (128*S_out is a multiple of 16)

template <int S>
__device__ __forceinline__ void ldkey(const uint64_t *__restrict__ keys, uint64_t i, uint64_t (&out)[S]) {
#pragma unroll
    for (int w = 0; w < S; ++w) out[w] = __ldg(keys + i * S + w);
}
template <int S>
__device__ __forceinline__ bool key_le(const uint64_t (&a)[S], const uint64_t (&b)[S]) {      // a <= b
#pragma unroll
    for (int w = 0; w < S; ++w) {
        if (a[w] != b[w]) return a[w] < b[w];
    }
    return true;
}

// Number of A elements among the first d elements of the merged order (ties: A first).
template <int S>
__device__ __forceinline__ uint64_t merge_path(const uint64_t *__restrict__ A, uint64_t na, const uint64_t *__restrict__ B, uint64_t nb,
                                               uint64_t d) {
    uint64_t lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        uint64_t a[S], b[S];
        ldkey<S>(A, mid, a);
        ldkey<S>(B, d - 1 - mid, b);
        if (key_le<S>(a, b)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Walks this thread's kVT merged elements.  EMIT=false: returns how many output records they start.
// EMIT=true: `pos` is the output position of the first record this thread starts; writes src_a / src_b.
template <int S, bool EMIT>
__device__ __forceinline__ uint32_t walk(const uint64_t *__restrict__ A, uint64_t na, const uint64_t *__restrict__ B, uint64_t nb,
                                         uint64_t d0, uint64_t d1, uint64_t pos, int64_t *__restrict__ src_a, int64_t *__restrict__ src_b) {
    if (d0 >= d1) return 0;
    uint64_t a = merge_path<S>(A, na, B, nb, d0), b = d0 - a;
    uint64_t ka[S], kb[S];
    bool have_a = a < na, have_b = b < nb;
    if (have_a) ldkey<S>(A, a, ka);
    if (have_b) ldkey<S>(B, b, kb);
    uint32_t started = 0;
    for (uint64_t d = d0; d < d1; ++d) {
        const bool take_a = have_a && (!have_b || key_le<S>(ka, kb));
        if (take_a) {
            if (EMIT) src_a[pos + started] = (int64_t)a;
            ++started;
            ++a;
            have_a = a < na;
            if (have_a) ldkey<S>(A, a, ka);
        } else {
            // duplicate iff the A key right before it in the merged order is equal (A keys are unique, ties put A first)
            bool dup = false;
            if (a > 0) {
                uint64_t prev[S];
                ldkey<S>(A, a - 1, prev);
                dup = true;
#pragma unroll
                for (int w = 0; w < S; ++w) dup &= (prev[w] == kb[w]);
            }
            if (dup) {
                if (EMIT) src_b[pos + started - 1] = (int64_t)b;      // joins the record its A twin started (maybe another thread's)
            } else {
                if (EMIT) { src_a[pos + started] = -1; src_b[pos + started] = (int64_t)b; }
                ++started;
            }
            ++b;
            have_b = b < nb;
            if (have_b) ldkey<S>(B, b, kb);
        }
    }
    return started;
}

template <int S, bool EMIT>
__global__ void __launch_bounds__(kJBlock) union_kernel(const uint64_t *__restrict__ A, uint64_t na, const uint64_t *__restrict__ B, uint64_t nb,
                                                        uint64_t *__restrict__ tile_count /* count pass: out; emit pass: exclusive offsets */,
                                                        int64_t *__restrict__ src_a, int64_t *__restrict__ src_b) {
    __shared__ uint32_t warp_sum[kJBlock / 32];
    const uint64_t total = na + nb;
    const uint64_t t0 = (uint64_t)blockIdx.x * kTile;
    const uint64_t d0 = min(t0 + (uint64_t)threadIdx.x * kVT, total), d1 = min(d0 + kVT, total);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // pass 1 for both modes: how many records does each thread start
    uint32_t mine = walk<S, false>(A, na, B, nb, d0, d1, 0, nullptr, nullptr);
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += v;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kJBlock / 32; ++w) {
        if (w < (int)warp) before += warp_sum[w];
        block_total += warp_sum[w];
    }
    if (!EMIT) {
        if (threadIdx.x == 0) tile_count[blockIdx.x] = block_total;
    } else {
        const uint64_t pos = tile_count[blockIdx.x] + before + (inc - mine);
        // a duplicate B key at the very start of this thread's range attaches to position pos-1, which exists because its
        // A twin precedes it in the merged order
        walk<S, true>(A, na, B, nb, d0, d1, pos, src_a, src_b);
    }
}

template <typename IdxT>
struct ComposeParams {
    const uint8_t *body_a, *body_b;
    const uint64_t *keys_a, *keys_b;
    const IdxT *src_a, *src_b;                 // source record in A / B, or IdxT(-1)
    uint8_t *out;
    uint64_t n_out;
    uint32_t s, ca, cb, Sa, Sb, So;
};

// Output record: s words, (ca + cb) coverages, (ca + cb) edge bytes.
template <typename IdxT>
__global__ void __launch_bounds__(kComposeRecords) compose_kernel(const ComposeParams<IdxT> p) {
    extern __shared__ __align__(16) uint8_t cmp_smem[];
    const uint64_t r = (uint64_t)blockIdx.x * kComposeRecords + threadIdx.x;
    if (r < p.n_out) {
        const IdxT ia = p.src_a[r], ib = p.src_b[r];
        uint8_t *d = cmp_smem + (size_t)threadIdx.x * p.So;
        const uint8_t *ra = ia != (IdxT)-1 ? p.body_a + (uint64_t)ia * p.Sa : nullptr;
        const uint8_t *rb = ib != (IdxT)-1 ? p.body_b + (uint64_t)ib * p.Sb : nullptr;
        const uint8_t *rk = ra ? ra : rb;
        const uint32_t kb = 8u * p.s, c = p.ca + p.cb;
        for (uint32_t i = 0; i < kb; ++i) d[i] = rk[i];
        for (uint32_t i = 0; i < 4u * p.ca; ++i) d[kb + i] = ra ? ra[kb + i] : 0;
        for (uint32_t i = 0; i < 4u * p.cb; ++i) d[kb + 4u * p.ca + i] = rb ? rb[kb + i] : 0;
        for (uint32_t i = 0; i < p.ca; ++i) d[kb + 4u * c + i] = ra ? ra[kb + 4u * p.ca + i] : 0;
        for (uint32_t i = 0; i < p.cb; ++i) d[kb + 4u * c + p.ca + i] = rb ? rb[kb + 4u * p.cb + i] : 0;
    }
    __syncthreads();
    const uint64_t r0 = (uint64_t)blockIdx.x * kComposeRecords;
    const uint64_t nrec = min((uint64_t)kComposeRecords, p.n_out - r0);
    const uint64_t nbytes = nrec * p.So;
    uint8_t *dst = p.out + r0 * p.So;                      // r0*So is a multiple of 128: 16-byte aligned
    const uint64_t n16 = nbytes >> 4;
    for (uint64_t i = threadIdx.x; i < n16; i += kComposeRecords) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(cmp_smem)[i];
    for (uint64_t i = (n16 << 4) + threadIdx.x; i < nbytes; i += kComposeRecords) dst[i] = cmp_smem[i];
}

// ------------------------------------------------------------------ tiled union (round 2)
constexpr int kJT = 256;                       // threads per CTA of the tiled kernels
constexpr uint32_t kJoinTileMax = 2048, kJoinTileMin = 32;
constexpr uint16_t kNoSrc = 0xffffu;

template <int S>
__device__ __forceinline__ void ldkey_s(const uint64_t *keys, uint32_t i, uint64_t (&out)[S]) {
#pragma unroll
    for (int w = 0; w < S; ++w) out[w] = keys[(size_t)i * S + w];
}

// Tile boundaries: (a_split[t], b_split[t]) = how many A and B elements precede merged position t*T (ties: A first),
// with the B element moved into the earlier tile when it equals the A element right before it.
template <int S>
__global__ void join_split_kernel(const uint64_t *__restrict__ A, uint64_t na, const uint64_t *__restrict__ B, uint64_t nb, uint32_t T,
                                  uint64_t ntiles, uint64_t *__restrict__ a_split, uint64_t *__restrict__ b_split) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const uint64_t d = min(t * T, na + nb);
    const uint64_t a = merge_path<S>(A, na, B, nb, d);
    uint64_t b = d - a;
    if (a > 0 && b < nb) {
        uint64_t ka[S], kb[S];
        ldkey<S>(A, a - 1, ka);
        ldkey<S>(B, b, kb);
        bool eq = true;
#pragma unroll
        for (int w = 0; w < S; ++w) eq &= ka[w] == kb[w];
        if (eq) ++b;
    }
    a_split[t] = a;
    b_split[t] = b;
}

// The merge of a tile's key slices (shared memory), by this thread, over merged positions [d0, d1).  EMIT = false: how
// many output records start there.  EMIT = true: `pos` = tile-local output position of the first one; fills src_a /
// src_b (src_b pre-filled with kNoSrc: a duplicate B key joins the record its A twin started, maybe another thread's).
template <int S, bool EMIT>
__device__ __forceinline__ uint32_t tile_walk(const uint64_t *KA, uint32_t na, const uint64_t *KB, uint32_t nb, uint32_t d0, uint32_t d1,
                                              uint32_t pos, uint16_t *src_a, uint16_t *src_b) {
    if (d0 >= d1) return 0;
    uint32_t lo = d0 > nb ? d0 - nb : 0, hi = min(d0, na);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint64_t a[S], b[S];
        ldkey_s<S>(KA, mid, a);
        ldkey_s<S>(KB, d0 - 1 - mid, b);
        if (key_le<S>(a, b)) lo = mid + 1; else hi = mid;
    }
    uint32_t a = lo, b = d0 - lo;
    uint64_t ka[S], kb[S];
    bool have_a = a < na, have_b = b < nb;
    if (have_a) ldkey_s<S>(KA, a, ka);
    if (have_b) ldkey_s<S>(KB, b, kb);
    uint32_t started = 0;
    for (uint32_t d = d0; d < d1; ++d) {
        if (have_a && (!have_b || key_le<S>(ka, kb))) {
            if (EMIT) src_a[pos + started] = (uint16_t)a;
            ++started;
            ++a;
            have_a = a < na;
            if (have_a) ldkey_s<S>(KA, a, ka);
        } else {
            bool dup = false;
            if (a > 0) {                       // the A twin, if any, is the A key right before (and in this tile: join_split_kernel)
                uint64_t prev[S];
                ldkey_s<S>(KA, a - 1, prev);
                dup = true;
#pragma unroll
                for (int w = 0; w < S; ++w) dup &= prev[w] == kb[w];
            }
            if (dup) {
                if (EMIT) src_b[pos + started - 1] = (uint16_t)b;
            } else {
                if (EMIT) { src_a[pos + started] = kNoSrc; src_b[pos + started] = (uint16_t)b; }
                ++started;
            }
            ++b;
            have_b = b < nb;
            if (have_b) ldkey_s<S>(KB, b, kb);
        }
    }
    return started;
}

// Exclusive prefix of `mine` over the CTA; total in *block_total.  One barrier inside.
__device__ __forceinline__ uint32_t block_exclusive(uint32_t mine, uint32_t *warp_sum, uint32_t *block_total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += v;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kJT / 32; ++w) {
        if (w < (int)warp) before += warp_sum[w];
        total += warp_sum[w];
    }
    *block_total = total;
    return before + inc - mine;
}

template <int S>
__global__ void __launch_bounds__(kJT) join_count_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                         const uint64_t *__restrict__ a_split, const uint64_t *__restrict__ b_split,
                                                         uint64_t *__restrict__ tile_count) {
    extern __shared__ __align__(16) uint8_t join_smem[];
    __shared__ uint32_t warp_sum[kJT / 32];
    const uint64_t t = blockIdx.x;
    const uint64_t a0 = a_split[t], b0 = b_split[t];
    const uint32_t na = (uint32_t)(a_split[t + 1] - a0), nb = (uint32_t)(b_split[t + 1] - b0), tot = na + nb;
    uint64_t *KA = reinterpret_cast<uint64_t *>(join_smem), *KB = KA + (size_t)na * S;
    for (uint32_t j = threadIdx.x; j < na * S; j += kJT) KA[j] = A[a0 * S + j];
    for (uint32_t j = threadIdx.x; j < nb * S; j += kJT) KB[j] = B[b0 * S + j];
    __syncthreads();
    const uint32_t vt = (tot + kJT - 1) / kJT, d0 = min(threadIdx.x * vt, tot), d1 = min(d0 + vt, tot);
    const uint32_t mine = tile_walk<S, false>(KA, na, KB, nb, d0, d1, 0, nullptr, nullptr);
    uint32_t total;
    block_exclusive(mine, warp_sum, &total);
    if (threadIdx.x == 0) tile_count[t] = total;
}

struct JoinEmit {
    const uint64_t *keys_a, *keys_b, *a_split, *b_split, *tile_off;
    uint32_t *src_a, *src_b;                   // per output record: source record in A / B, or 0xffffffff
    uint64_t *out_keys;                        // key column of the result; may be null
    uint32_t T;
};

template <int S>
__global__ void __launch_bounds__(kJT) join_emit_kernel(const JoinEmit p) {
    extern __shared__ __align__(16) uint8_t join_smem[];
    __shared__ uint32_t warp_sum[kJT / 32];
    const uint64_t t = blockIdx.x;
    const uint64_t a0 = p.a_split[t], b0 = p.b_split[t], pos0 = p.tile_off[t];
    const uint32_t na = (uint32_t)(p.a_split[t + 1] - a0), nb = (uint32_t)(p.b_split[t + 1] - b0), tot = na + nb;
    const uint32_t cap = p.T + 1;
    uint64_t *KA = reinterpret_cast<uint64_t *>(join_smem), *KB = KA + (size_t)na * S;
    uint16_t *src_a = reinterpret_cast<uint16_t *>(join_smem + (size_t)cap * S * 8), *src_b = src_a + cap;
    for (uint32_t j = threadIdx.x; j < na * S; j += kJT) KA[j] = p.keys_a[a0 * S + j];
    for (uint32_t j = threadIdx.x; j < nb * S; j += kJT) KB[j] = p.keys_b[b0 * S + j];
    for (uint32_t j = threadIdx.x; j < cap; j += kJT) src_b[j] = kNoSrc;
    __syncthreads();
    const uint32_t vt = (tot + kJT - 1) / kJT, d0 = min(threadIdx.x * vt, tot), d1 = min(d0 + vt, tot);
    const uint32_t mine = tile_walk<S, false>(KA, na, KB, nb, d0, d1, 0, nullptr, nullptr);
    uint32_t n_out;
    const uint32_t before = block_exclusive(mine, warp_sum, &n_out);
    tile_walk<S, true>(KA, na, KB, nb, d0, d1, before, src_a, src_b);
    __syncthreads();
    // the tile's source lists (and keys) leave in output order: coalesced
    for (uint32_t r = threadIdx.x; r < n_out; r += kJT) {
        const uint32_t ia = src_a[r], ib = src_b[r];
        p.src_a[pos0 + r] = ia != kNoSrc ? (uint32_t)(a0 + ia) : 0xffffffffu;
        p.src_b[pos0 + r] = ib != kNoSrc ? (uint32_t)(b0 + ib) : 0xffffffffu;
        if (p.out_keys) {
            const uint64_t *kk = ia != kNoSrc ? KA + (size_t)ia * S : KB + (size_t)ib * S;
#pragma unroll
            for (int w = 0; w < S; ++w) p.out_keys[(pos0 + r) * S + w] = kk[w];
        }
    }
}

inline size_t join_tile_smem(uint32_t T, uint32_t s) { return (size_t)(T + 1) * (s * 8 + 4) + 16; }

#define CC_JOIN_DISPATCH_S(s, ...)                                                        \
    switch (s) {                                                                          \
        case 1: { constexpr int S_ = 1; __VA_ARGS__; break; }                             \
        case 2: { constexpr int S_ = 2; __VA_ARGS__; break; }                             \
        case 3: { constexpr int S_ = 3; __VA_ARGS__; break; }                             \
        case 4: { constexpr int S_ = 4; __VA_ARGS__; break; }                             \
        default: return fail(CC_ERR_UNSUPPORTED, "k-mers wider than 4 words (k > 128) are not supported by join"); \
    }

}  // namespace

// out[i] = body[perm[i]] for whole records, staged through shared memory so the output leaves as aligned 16-byte stores
__global__ void __launch_bounds__(kComposeRecords) gather_records_kernel(const uint8_t *__restrict__ body, const uint32_t *__restrict__ perm,
                                                                         uint64_t n, uint32_t S, uint8_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t gat_smem[];
    const uint64_t r = (uint64_t)blockIdx.x * kComposeRecords + threadIdx.x;
    if (r < n) {
        const uint8_t *src = body + (uint64_t)perm[r] * S;
        uint8_t *d = gat_smem + (size_t)threadIdx.x * S;
        for (uint32_t i = 0; i < S; ++i) d[i] = src[i];
    }
    __syncthreads();
    const uint64_t r0 = (uint64_t)blockIdx.x * kComposeRecords;
    const uint64_t nbytes = min((uint64_t)kComposeRecords, n - r0) * S;
    uint8_t *dst = out + r0 * S;
    const uint64_t n16 = nbytes >> 4;
    for (uint64_t i = threadIdx.x; i < n16; i += kComposeRecords) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(gat_smem)[i];
    for (uint64_t i = (n16 << 4) + threadIdx.x; i < nbytes; i += kComposeRecords) dst[i] = gat_smem[i];
}

int launch_gather_records(const uint8_t *body, const uint32_t *perm, uint64_t n, uint32_t S, uint8_t *out, cudaStream_t st) {
    if (n == 0) return CC_OK;
    const size_t smem = (size_t)kComposeRecords * S;
    CC_CUDA(cudaFuncSetAttribute(gather_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gather_records_kernel<<<(unsigned)((n + kComposeRecords - 1) / kComposeRecords), kComposeRecords, smem, st>>>(body, perm, n, S, out);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// Two-way union on the device, round-1 form (see the header comment): one global merge-path search per thread.
static int join_pair_untiled(const uint8_t *body_a, const uint64_t *keys_a, uint64_t na, uint32_t ca, const uint8_t *body_b, const uint64_t *keys_b,
                             uint64_t nb, uint32_t cb, uint32_t s, cudaStream_t st, void **out_body, uint64_t *out_n) {
    const uint64_t total = na + nb;
    const uint32_t Sa = 8 * s + 5 * ca, Sb = 8 * s + 5 * cb, So = 8 * s + 5 * (ca + cb);
    const uint64_t ntiles = (total + kTile - 1) / kTile;
    uint64_t *tile_cnt = nullptr, *tile_off = nullptr;
    int64_t *src_a = nullptr, *src_b = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    struct Free {
        cudaStream_t st; uint64_t *&a, *&b; int64_t *&c, *&d; void *&e;
        ~Free() { cudaFreeAsync(a, st); cudaFreeAsync(b, st); cudaFreeAsync(c, st); cudaFreeAsync(d, st); cudaFreeAsync(e, st); }
    } fr{st, tile_cnt, tile_off, src_a, src_b, tmp};
    CC_CUDA(cudaMallocAsync(&tile_cnt, (ntiles + 1) * 8, st));
    CC_CUDA(cudaMallocAsync(&tile_off, (ntiles + 1) * 8, st));
    CC_CUDA(cudaMemsetAsync(tile_cnt, 0, (ntiles + 1) * 8, st));
    CC_JOIN_DISPATCH_S(s, union_kernel<S_, false><<<(unsigned)ntiles, kJBlock, 0, st>>>(keys_a, na, keys_b, nb, tile_cnt, nullptr, nullptr));
    count_launch();
    CC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_cnt, tile_off, ntiles + 1, st));
    CC_CUDA(cudaMallocAsync(&tmp, tmp_bytes, st));
    CC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_cnt, tile_off, ntiles + 1, st));
    count_launch();
    uint64_t n_out = 0;
    CC_CUDA(cudaMemcpyAsync(&n_out, tile_off + ntiles, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    CC_CUDA(cudaMallocAsync(&src_a, std::max<uint64_t>(n_out, 1) * 8, st));
    CC_CUDA(cudaMallocAsync(&src_b, std::max<uint64_t>(n_out, 1) * 8, st));
    CC_CUDA(cudaMemsetAsync(src_b, 0xff, std::max<uint64_t>(n_out, 1) * 8, st));
    CC_JOIN_DISPATCH_S(s, union_kernel<S_, true><<<(unsigned)ntiles, kJBlock, 0, st>>>(keys_a, na, keys_b, nb, tile_off, src_a, src_b));
    count_launch();
    CC_CUDA(cudaMalloc(out_body, n_out * So + 256));
    ComposeParams<int64_t> p{body_a, body_b, keys_a, keys_b, src_a, src_b, static_cast<uint8_t *>(*out_body), n_out, s, ca, cb, Sa, Sb, So};
    if (n_out) {
        const size_t smem = (size_t)kComposeRecords * So;
        CC_CUDA(cudaFuncSetAttribute(compose_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        compose_kernel<int64_t><<<(unsigned)((n_out + kComposeRecords - 1) / kComposeRecords), kComposeRecords, smem, st>>>(p);
        count_launch();
    }
    CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(*out_body) + n_out * So, 0, 256, st));
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaStreamSynchronize(st));
    *out_n = n_out;
    return CC_OK;
}

// Two-way union on the device.  keys_* are the sorted key columns; body_* the record arrays.  Allocates *out_body
// (cudaMalloc: caller frees with cudaFree; scratch comes from the stream-ordered pool -- growing the pool by gigabytes on a
// first call costs ten times what cudaMalloc does, so the results stay outside it) and returns the number of output records; with out_keys != nullptr also allocates and
// fills the key column of the result (cudaFree), or leaves *out_keys null when the untiled form ran.
int join_pair(const uint8_t *body_a, const uint64_t *keys_a, uint64_t na, uint32_t ca, const uint8_t *body_b, const uint64_t *keys_b,
              uint64_t nb, uint32_t cb, uint32_t s, cudaStream_t st, void **out_body, uint64_t *out_n, uint64_t **out_keys) {
    const uint64_t total = na + nb;
    const uint32_t Sa = 8 * s + 5 * ca, Sb = 8 * s + 5 * cb, So = 8 * s + 5 * (ca + cb);
    *out_body = nullptr;
    *out_n = 0;
    if (out_keys) *out_keys = nullptr;
    if (total == 0) {
        CC_CUDA(cudaMalloc(out_body, 256));
        return CC_OK;
    }
    // tile = as many merged elements as the shared-memory budget holds keys for (8 per thread at the default)
    uint32_t T = kJoinTileMax;
    while (T > kJoinTileMin && join_tile_smem(T, s) > (size_t)options().join_tile_kb * 1024) T >>= 1;
    if (!options().join_tiled || na >= 0xffffffffull || nb >= 0xffffffffull || (size_t)kComposeRecords * So > 200 * 1024)
        return join_pair_untiled(body_a, keys_a, na, ca, body_b, keys_b, nb, cb, s, st, out_body, out_n);
    const bool trace = getenv("CC_JOIN_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_0 = now();
    const uint64_t ntiles = (total + T - 1) / T;
    uint64_t *a_split = nullptr, *b_split = nullptr, *tile_cnt = nullptr, *tile_off = nullptr;
    uint32_t *src_a = nullptr, *src_b = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    struct Free {
        cudaStream_t st; uint64_t *&a, *&b, *&c, *&d; void *&e; uint32_t *&f, *&g;
        ~Free() { cudaFreeAsync(a, st); cudaFreeAsync(b, st); cudaFreeAsync(c, st); cudaFreeAsync(d, st); cudaFreeAsync(e, st);
                  cudaFreeAsync(f, st); cudaFreeAsync(g, st); }
    } fr{st, a_split, b_split, tile_cnt, tile_off, tmp, src_a, src_b};
    CC_CUDA(cudaMallocAsync(&a_split, (ntiles + 1) * 8, st));
    CC_CUDA(cudaMallocAsync(&b_split, (ntiles + 1) * 8, st));
    CC_CUDA(cudaMallocAsync(&tile_cnt, (ntiles + 1) * 8, st));
    CC_CUDA(cudaMallocAsync(&tile_off, (ntiles + 1) * 8, st));
    CC_CUDA(cudaMemsetAsync(tile_cnt, 0, (ntiles + 1) * 8, st));
    const size_t smem_tile = join_tile_smem(T, s);
    CC_JOIN_DISPATCH_S(s, {
        join_split_kernel<S_><<<(unsigned)((ntiles + 1 + 127) / 128), 128, 0, st>>>(keys_a, na, keys_b, nb, T, ntiles, a_split, b_split);
        CC_CUDA(cudaFuncSetAttribute(join_count_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile));
        join_count_kernel<S_><<<(unsigned)ntiles, kJT, smem_tile, st>>>(keys_a, keys_b, a_split, b_split, tile_cnt);
    });
    count_launch(2);
    CC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_cnt, tile_off, ntiles + 1, st));
    CC_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 16), st));
    CC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_cnt, tile_off, ntiles + 1, st));
    count_launch();
    uint64_t n_out = 0;
    CC_CUDA(cudaMemcpyAsync(&n_out, tile_off + ntiles, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    const double t_1 = now();
    CC_CUDA(cudaMalloc(out_body, n_out * So + 256));
    if (out_keys) CC_CUDA(cudaMalloc(out_keys, std::max<uint64_t>(n_out * s, 2) * 8 + 64));
    CC_CUDA(cudaMallocAsync(&src_a, std::max<uint64_t>(n_out, 1) * 4, st));
    CC_CUDA(cudaMallocAsync(&src_b, std::max<uint64_t>(n_out, 1) * 4, st));
    const double t_2 = now();
    JoinEmit e{keys_a, keys_b, a_split, b_split, tile_off, src_a, src_b, out_keys ? *out_keys : nullptr, T};
    CC_JOIN_DISPATCH_S(s, {
        CC_CUDA(cudaFuncSetAttribute(join_emit_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile));
        join_emit_kernel<S_><<<(unsigned)ntiles, kJT, smem_tile, st>>>(e);
    });
    count_launch();
    if (n_out) {
        ComposeParams<uint32_t> p{body_a, body_b, keys_a, keys_b, src_a, src_b, static_cast<uint8_t *>(*out_body), n_out, s, ca, cb, Sa, Sb, So};
        const size_t smem = (size_t)kComposeRecords * So;
        CC_CUDA(cudaFuncSetAttribute(compose_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        compose_kernel<uint32_t><<<(unsigned)((n_out + kComposeRecords - 1) / kComposeRecords), kComposeRecords, smem, st>>>(p);
        count_launch();
    }
    CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(*out_body) + n_out * So, 0, 256, st));
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaStreamSynchronize(st));
    if (trace) fprintf(stderr, "join_pair: split+count+scan %.2f ms, allocate outputs %.2f ms, emit+compose %.2f ms\n", t_1 - t_0, t_2 - t_1, now() - t_2);
    *out_n = n_out;
    return CC_OK;
}

}  // namespace cc
