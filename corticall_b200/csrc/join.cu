// join.cu -- merged view of several sorted graphs: the record stream CortexCollection.next() produces and Join writes.
//
// Reference semantics reproduced (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
//   S/utils/io/graph/cortex/CortexCollection.java:34-62    colours are concatenated in graph order
//   S/utils/io/graph/cortex/CortexCollection.java:245-293  next(): lowest k-mer among the graphs' heads; every graph
//                                                          holding that k-mer contributes its coverage / edges to its
//                                                          own colours, the others stay 0; output ascending by k-mer
//   S/commands/utils/Join.java:23-57                        Join = write that stream with CortexGraphWriter
//
// B200 design.  A k-way merge is folded into two-way unions.  One union of A and B (sorted, duplicate-free key columns):
// every thread finds its slice of the merged order (ties: A first) by a merge-path search, a counting pass (a B key equal to the A key right
// before it is a duplicate and yields no output of its own), an exclusive scan of the tile counts, an emit pass that
// writes, for every output record, the source index in A and in B (or -1), and a compose pass that gathers both source
// records into the wider output record through shared memory so the output leaves as aligned 16-byte stores.
// Everything is linear in |A| + |B|; no sort.
#include <cub/device/device_scan.cuh>

#include <algorithm>

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

struct ComposeParams {
    const uint8_t *body_a, *body_b;
    const uint64_t *keys_a, *keys_b;
    const int64_t *src_a, *src_b;
    uint8_t *out;
    uint64_t n_out;
    uint32_t s, ca, cb, Sa, Sb, So;
};

// Output record: s words, (ca + cb) coverages, (ca + cb) edge bytes.
__global__ void __launch_bounds__(kComposeRecords) compose_kernel(const ComposeParams p) {
    extern __shared__ __align__(16) uint8_t cmp_smem[];
    const uint64_t r = (uint64_t)blockIdx.x * kComposeRecords + threadIdx.x;
    if (r < p.n_out) {
        const int64_t ia = p.src_a[r], ib = p.src_b[r];
        uint8_t *d = cmp_smem + (size_t)threadIdx.x * p.So;
        const uint8_t *ra = ia >= 0 ? p.body_a + (uint64_t)ia * p.Sa : nullptr;
        const uint8_t *rb = ib >= 0 ? p.body_b + (uint64_t)ib * p.Sb : nullptr;
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

// Two-way union on the device.  keys_* are the sorted key columns; body_* the record arrays.  Allocates *out_body
// (caller frees with cudaFree) and returns the number of output records.
int join_pair(const uint8_t *body_a, const uint64_t *keys_a, uint64_t na, uint32_t ca, const uint8_t *body_b, const uint64_t *keys_b,
              uint64_t nb, uint32_t cb, uint32_t s, cudaStream_t st, void **out_body, uint64_t *out_n) {
    const uint64_t total = na + nb;
    const uint32_t Sa = 8 * s + 5 * ca, Sb = 8 * s + 5 * cb, So = 8 * s + 5 * (ca + cb);
    *out_body = nullptr;
    *out_n = 0;
    if (total == 0) {
        CC_CUDA(cudaMalloc(out_body, 256));
        return CC_OK;
    }
    const uint64_t ntiles = (total + kTile - 1) / kTile;
    uint64_t *tile_cnt = nullptr, *tile_off = nullptr;
    int64_t *src_a = nullptr, *src_b = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    struct Free {
        uint64_t *&a, *&b; int64_t *&c, *&d; void *&e;
        ~Free() { cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d); cudaFree(e); }
    } fr{tile_cnt, tile_off, src_a, src_b, tmp};
    CC_CUDA(cudaMalloc(&tile_cnt, (ntiles + 1) * 8));
    CC_CUDA(cudaMalloc(&tile_off, (ntiles + 1) * 8));
    CC_CUDA(cudaMemsetAsync(tile_cnt, 0, (ntiles + 1) * 8, st));
    CC_JOIN_DISPATCH_S(s, union_kernel<S_, false><<<(unsigned)ntiles, kJBlock, 0, st>>>(keys_a, na, keys_b, nb, tile_cnt, nullptr, nullptr));
    count_launch();
    CC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_cnt, tile_off, ntiles + 1, st));
    CC_CUDA(cudaMalloc(&tmp, tmp_bytes));
    CC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, tile_cnt, tile_off, ntiles + 1, st));
    count_launch();
    uint64_t n_out = 0;
    CC_CUDA(cudaMemcpyAsync(&n_out, tile_off + ntiles, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    CC_CUDA(cudaMalloc(&src_a, std::max<uint64_t>(n_out, 1) * 8));
    CC_CUDA(cudaMalloc(&src_b, std::max<uint64_t>(n_out, 1) * 8));
    CC_CUDA(cudaMemsetAsync(src_b, 0xff, std::max<uint64_t>(n_out, 1) * 8, st));
    CC_JOIN_DISPATCH_S(s, union_kernel<S_, true><<<(unsigned)ntiles, kJBlock, 0, st>>>(keys_a, na, keys_b, nb, tile_off, src_a, src_b));
    count_launch();
    CC_CUDA(cudaMalloc(out_body, n_out * So + 256));
    ComposeParams p{body_a, body_b, keys_a, keys_b, src_a, src_b, static_cast<uint8_t *>(*out_body), n_out, s, ca, cb, Sa, Sb, So};
    if (n_out) {
        const size_t smem = (size_t)kComposeRecords * So;
        CC_CUDA(cudaFuncSetAttribute(compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        compose_kernel<<<(unsigned)((n_out + kComposeRecords - 1) / kComposeRecords), kComposeRecords, smem, st>>>(p);
        count_launch();
    }
    CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(*out_body) + n_out * So, 0, 256, st));
    CC_CUDA(cudaGetLastError());
    CC_CUDA(cudaStreamSynchronize(st));
    *out_n = n_out;
    return CC_OK;
}

}  // namespace cc
