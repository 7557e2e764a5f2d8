// lookup.cu -- K3 (canonicalise + 2-bit pack) and K4 (batched sorted-array lookups), plus the
// multi-GPU bucket / scatter helpers around the all-to-all.
//
// Reference semantics reproduced (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
//   canonical orientation   S/utils/sequence/SequenceUtils.java:206-225  (ASCII, signed bytes, tie -> forward)
//   complement              S/utils/sequence/SequenceUtils.java:61-86
//   2-bit packing           S/utils/io/graph/cortex/CortexRecord.java:313-360 (A0 C1 G2 T3, lower case accepted)
//   findRecord              S/utils/io/graph/cortex/CortexGraph.java:272-317: canonicalise, then EQUALITY of
//                           the ASCII query with a decoded (upper-case ACGT) record k-mer; for a sorted,
//                           duplicate-free graph with N >= 3 that is "index of the exact match, else null",
//                           and any query holding a byte outside ACGT (N, lower case) is a miss.
//
// B200 design.  K3: a CTA stages a tile of the sequence in shared memory, converts it once into a 2-bit
// big-endian bit stream plus "not ACGTacgt" and "lower case" bit masks, and every thread cuts its window
// out of the streams with funnel shifts (5 LDS for k=47), reverse-complements in registers (brev + pair
// swap + multiword shift) and keeps the smaller.  K4: the key column (records stripped of coverage and
// edges, 8s bytes per key) is searched through a prefix table over the top `bits` bits of the k-mer
// (lower bounds per prefix, sized to stay L2-resident), so a lookup costs one table read plus one short
// bucket scan issued as independent loads instead of ~27 dependent probes.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>

#include "cc_internal.hpp"
#include "device_utils.cuh"

namespace cc {

namespace {

constexpr int kBlock = 256;
constexpr uint32_t kTileBytes = 8192;          // sequence bytes staged per CTA iteration
constexpr uint32_t kPadBases = 32;             // zero bases in front of the streams (negative offsets of word 0)
constexpr uint32_t kStreamBases = kTileBytes + kPadBases + 64;

struct SeqTile {
    __align__(16) uint8_t ascii[kTileBytes + 64];
    uint32_t codes[kStreamBases / 16 + 4];     // 2 bits per base, base b in word b/16, bits 31-2*(b%16)..
    uint32_t inval[kStreamBases / 32 + 4];     // 1 bit per base: byte outside ACGTacgt
    uint32_t lower[kStreamBases / 32 + 4];     // 1 bit per base: acgt
};

// ------------------------------------------------------------------ tile staging
// Copies bytes [b0, b0+nb) of seq into t.ascii (at offset `off` = source misalignment, so 16-byte chunks
// are aligned on both sides) and builds the three bit streams (positions offset by kPadBases).
// Returns `off`; tile byte i lives at t.ascii[off + i].
__device__ __forceinline__ uint32_t stage_tile(SeqTile &t, const uint8_t *__restrict__ seq, uint64_t b0, uint32_t nb) {
    const uint8_t *src = seq + b0;
    const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    uint8_t *dst = t.ascii + off;
    const uint32_t head = (16u - off) & 15u;
    // head bytes (unaligned prefix), 16-byte body, tail bytes -- never reads outside [src, src+nb)
    for (uint32_t i = threadIdx.x; i < min(head, nb); i += blockDim.x) dst[i] = src[i];
    if (nb > head) {
        const uint32_t body = (nb - head) & ~15u;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
        for (uint32_t i = threadIdx.x; i < body / 16; i += blockDim.x) d4[i] = __ldg(s4 + i);
        for (uint32_t i = head + body + threadIdx.x; i < nb; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // streams: group g covers stream positions [32g, 32g+32); position p holds tile byte p - kPadBases
    const uint32_t ngroups = (nb + kPadBases + 31) / 32 + 2;
    for (uint32_t g = threadIdx.x; g < ngroups; g += blockDim.x) {
        uint32_t c0 = 0, c1 = 0, inv = 0, low = 0;
#pragma unroll 8
        for (uint32_t j = 0; j < 32; ++j) {
            const int32_t pos = (int32_t)(g * 32 + j) - (int32_t)kPadBases;
            uint32_t code = 0, bad = 0, lc = 0;
            if (pos >= 0 && pos < (int32_t)nb) {
                const uint32_t ch = dst[pos];
                const uint32_t f = ch | 0x20u;
                const bool ok = (f == 'a') | (f == 'c') | (f == 'g') | (f == 't');
                code = ok ? base_code(ch) : 0u;
                bad = ok ? 0u : 1u;
                lc = (ok && (ch & 0x20u)) ? 1u : 0u;
            }
            if (j < 16) c0 |= code << (30 - 2 * j); else c1 |= code << (30 - 2 * (j - 16));
            inv |= bad << (31 - j);
            low |= lc << (31 - j);
        }
        t.codes[2 * g] = c0;
        t.codes[2 * g + 1] = c1;
        t.inval[g] = inv;
        t.lower[g] = low;
    }
    __syncthreads();
    return off;
}

// 64 bits starting at bit offset P of a big-endian u32 bit stream.
__device__ __forceinline__ uint64_t extract64(const uint32_t *w, uint32_t P) {
    const uint32_t idx = P >> 5, sh = P & 31u;
    const uint32_t w0 = w[idx], w1 = w[idx + 1], w2 = w[idx + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, sh);
    const uint32_t lo = __funnelshift_l(w2, w1, sh);
    return ((uint64_t)hi << 32) | lo;
}
// OR of `nbits` bits starting at bit offset P of a 1-bit-per-base stream.
__device__ __forceinline__ bool any_bits(const uint32_t *w, uint32_t P, uint32_t nbits) {
    uint32_t acc = 0;
    for (uint32_t done = 0; done < nbits; done += 32) {
        const uint32_t q = P + done, idx = q >> 5, sh = q & 31u;
        uint32_t v = __funnelshift_l(w[idx + 1], w[idx], sh);
        const uint32_t rem = nbits - done;
        if (rem < 32) v &= ~(0xffffffffu >> rem);
        acc |= v;
    }
    return acc != 0;
}

// Canonical packed k-mer of the window starting at tile byte `start`.  Returns flags
// (bit0 flipped, bit1 not ACGTacgt, bit2 has lower case); words zeroed when bit1.
template <int S>
__device__ __forceinline__ uint32_t window_canonical(const SeqTile &t, uint32_t ascii_off, uint32_t start, uint32_t k, uint64_t (&out)[S]) {
    const uint32_t p0 = start + kPadBases;                   // stream position of the window's first base
    if (any_bits(t.inval, p0, k)) {
#pragma unroll
        for (int i = 0; i < S; ++i) out[i] = 0;
        return 2u;
    }
    uint64_t fw[S], rc[S];
#pragma unroll
    for (int j = 0; j < S; ++j) {
        // word j holds bases [k - 32(S-j), k - 32(S-j-1)); the first word may start before the window
        const int32_t b0 = (int32_t)k - 32 * (S - j);
        fw[j] = extract64(t.codes, 2u * (uint32_t)((int32_t)p0 + b0));
    }
    const uint32_t top_bits = 2u * k - 64u * (S - 1);
    if (top_bits < 64) fw[0] &= (1ull << top_bits) - 1ull;
    revcomp_words<S>(fw, rc, k);
    bool flip = words_less<S>(rc, fw);
    uint32_t flags = 0;
    if (any_bits(t.lower, p0, k)) {
        // Mixed / lower case: the reference compares ASCII bytes (signed), not 2-bit codes
        // (SequenceUtils.java:211-219).  Rare, so take the byte loop.
        flags |= 4u;
        flip = false;
        const uint8_t *a = t.ascii + ascii_off + start;
        for (uint32_t i = 0; i < k; ++i) {
            const int8_t f = (int8_t)a[i];
            const int8_t r = (int8_t)complement_ascii(a[k - 1 - i]);
            if (f < r) break;
            if (f > r) { flip = true; break; }
        }
    }
#pragma unroll
    for (int i = 0; i < S; ++i) out[i] = flip ? rc[i] : fw[i];
    return flags | (flip ? 1u : 0u);
}

// ------------------------------------------------------------------ key column access + prefix table
template <int S>
__device__ __forceinline__ void load_key(const uint64_t *__restrict__ keys, uint64_t i, uint64_t (&out)[S]) {
    if (S == 2) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(keys) + i);
        out[0] = v.x; out[1 % S] = v.y;
    } else {
#pragma unroll
        for (int w = 0; w < S; ++w) out[w] = __ldg(keys + i * S + w);
    }
}

struct IndexView {
    const uint64_t *keys;
    const uint32_t *table;
    uint64_t n;
    uint64_t first_index;
    uint32_t bits;      // table covers the top `bits` bits of the 2k-bit key
    uint32_t shift;     // 2k - bits
};

template <int S>
__device__ __forceinline__ uint32_t key_prefix(const uint64_t (&q)[S], uint32_t shift) {
    const uint32_t wsh = shift >> 6, bsh = shift & 63u;
    const int hi_w = S - 1 - (int)wsh;
    uint64_t low = 0;
#pragma unroll
    for (int w = 0; w < S; ++w) {
        if (w == hi_w) low |= q[w] >> bsh;
        if (w == hi_w - 1 && bsh) low |= q[w] << (64 - bsh);
    }
    return (uint32_t)low;
}

constexpr int kMaxShards = 64;
constexpr uint32_t kLinear = 8;     // bucket remainder scanned with independent loads

// Index of the exact match of q in keys[lo, hi) (lowest on duplicates), or -1.
template <int S>
__device__ __forceinline__ int64_t search_range(const uint64_t *__restrict__ keys, uint64_t lo, uint64_t hi, const uint64_t (&q)[S]) {
    while (hi - lo > kLinear) {                  // lower_bound steps
        const uint64_t mid = lo + ((hi - lo) >> 1);
        uint64_t km[S];
        load_key<S>(keys, mid, km);
        if (words_less<S>(km, q)) lo = mid + 1; else hi = mid + 1;   // keep the first key >= q inside [lo, hi)
    }
    uint64_t kk[kLinear][S];
#pragma unroll
    for (uint32_t j = 0; j < kLinear; ++j) {
        if (lo + j < hi) load_key<S>(keys, lo + j, kk[j]);
        else {
#pragma unroll
            for (int w = 0; w < S; ++w) kk[j][w] = ~0ull;
        }
    }
    int64_t res = -1;
#pragma unroll
    for (int j = (int)kLinear - 1; j >= 0; --j) {
        if (lo + j < hi && words_equal<S>(kk[j], q)) res = (int64_t)(lo + j);
    }
    return res;
}

template <int S>
__device__ __forceinline__ int64_t lookup_bucketed(const IndexView &ix, const uint64_t (&q)[S]) {
    const uint32_t p = key_prefix<S>(q, ix.shift);
    const uint64_t lo = __ldg(ix.table + p), hi = __ldg(ix.table + p + 1);
    const int64_t r = search_range<S>(ix.keys, lo, hi, q);
    return r < 0 ? r : r + (int64_t)ix.first_index;
}
template <int S>
__device__ __forceinline__ int64_t lookup_bsearch(const IndexView &ix, const uint64_t (&q)[S]) {
    const int64_t r = search_range<S>(ix.keys, 0, ix.n, q);
    return r < 0 ? r : r + (int64_t)ix.first_index;
}

// ------------------------------------------------------------------ kernels: pack, find (fused with pack), find packed
struct SeqJob {
    const uint8_t *seq;
    uint64_t nq;          // windows / rows
    uint64_t stride;      // 1 = sliding windows, k = independent rows
    uint32_t k;
    uint32_t per_tile;    // windows per tile
};

template <int S, bool FIND, bool BUCKETED>
__global__ void __launch_bounds__(kBlock) seq_kernel(SeqJob job, uint64_t *__restrict__ out_words, uint8_t *__restrict__ out_flags,
                                                     IndexView ix, int64_t *__restrict__ out_index) {
    __shared__ SeqTile tile;
    const uint64_t ntiles = (job.nq + job.per_tile - 1) / job.per_tile;
    for (uint64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
        const uint64_t w0 = tix * job.per_tile;
        const uint32_t nw = (uint32_t)min((uint64_t)job.per_tile, job.nq - w0);
        const uint32_t nb = (uint32_t)((nw - 1) * job.stride + job.k);
        __syncthreads();                                    // previous tile fully consumed
        const uint32_t ascii_off = stage_tile(tile, job.seq, w0 * job.stride, nb);
        for (uint32_t j = threadIdx.x; j < nw; j += kBlock) {
            uint64_t q[S];
            const uint32_t flags = window_canonical<S>(tile, ascii_off, (uint32_t)(j * job.stride), job.k, q);
            if (FIND) {
                int64_t r = -1;
                if ((flags & 6u) == 0) r = BUCKETED ? lookup_bucketed<S>(ix, q) : lookup_bsearch<S>(ix, q);
                out_index[w0 + j] = r;
            } else {
                if (S == 2) {
                    reinterpret_cast<ulonglong2 *>(out_words)[w0 + j] = make_ulonglong2(q[0], q[1 % S]);
                } else {
#pragma unroll
                    for (int w = 0; w < S; ++w) out_words[(w0 + j) * S + w] = q[w];
                }
                out_flags[w0 + j] = (uint8_t)flags;
            }
        }
    }
}

// ------------------------------------------------------------------ K3 for independent k-byte rows (query lists)
// A warp takes 32 consecutive rows (32*k contiguous bytes): 16-byte coalesced loads of the aligned superset into a
// per-warp shared buffer, then lane i packs row i base by base.  Same flags and canonical rule as window_canonical.
template <int S>
__global__ void __launch_bounds__(kBlock) pack_rows_kernel(const uint8_t *__restrict__ kmers, uint64_t nq, uint32_t k,
                                                           uint64_t *__restrict__ out_words, uint8_t *__restrict__ out_flags) {
    extern __shared__ __align__(16) uint8_t rows_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t wbytes = (32u * k + 32u + 15u) & ~15u;
    uint8_t *buf = rows_smem + (size_t)warp * wbytes;
    const uint64_t ngroups = (nq + 31) / 32;
    const uint64_t gstride = (uint64_t)gridDim.x * (kBlock / 32);
    for (uint64_t grp = (uint64_t)blockIdx.x * (kBlock / 32) + warp; grp < ngroups; grp += gstride) {
        const uint64_t row0 = grp * 32;
        const uint32_t rows = (uint32_t)min((uint64_t)32, nq - row0);
        const uint8_t *src = kmers + row0 * k;
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint4 *src4 = reinterpret_cast<const uint4 *>(src - shift);
        const uint32_t n16 = (shift + rows * k + 15u) >> 4;
        __syncwarp();
        for (uint32_t i = lane; i < n16; i += 32) reinterpret_cast<uint4 *>(buf)[i] = __ldg(src4 + i);
        __syncwarp();
        if (lane < rows) {
            const uint8_t *a = buf + shift + lane * k;
            uint64_t fw[S];
#pragma unroll
            for (int w = 0; w < S; ++w) fw[w] = 0;
            uint32_t bad = 0, low = 0;
            for (uint32_t i = 0; i < k; ++i) {
                const uint32_t ch = a[i];
                const uint32_t f = ch | 0x20u;
                const bool ok = (f == 'a') | (f == 'c') | (f == 'g') | (f == 't');
                bad |= ok ? 0u : 1u;
                low |= (ok && (ch & 0x20u)) ? 1u : 0u;
                const uint64_t code = ok ? base_code(ch) : 0u;
#pragma unroll
                for (int w = 0; w < S - 1; ++w) fw[w] = (fw[w] << 2) | (fw[w + 1] >> 62);
                fw[S - 1] = (fw[S - 1] << 2) | code;
            }
            uint32_t flags = 0;
            uint64_t outw[S];
            if (bad) {
#pragma unroll
                for (int w = 0; w < S; ++w) outw[w] = 0;
                flags = 2u;
            } else {
                uint64_t rc[S];
                revcomp_words<S>(fw, rc, k);
                bool flip = words_less<S>(rc, fw);
                if (low) {          // mixed / lower case: the reference compares ASCII bytes (SequenceUtils.java:211-219)
                    flags |= 4u;
                    flip = false;
                    for (uint32_t i = 0; i < k; ++i) {
                        const int8_t f = (int8_t)a[i];
                        const int8_t r = (int8_t)complement_ascii(a[k - 1 - i]);
                        if (f < r) break;
                        if (f > r) { flip = true; break; }
                    }
                }
#pragma unroll
                for (int w = 0; w < S; ++w) outw[w] = flip ? rc[w] : fw[w];
                flags |= flip ? 1u : 0u;
            }
            const uint64_t row = row0 + lane;
            if (S == 2) reinterpret_cast<ulonglong2 *>(out_words)[row] = make_ulonglong2(outw[0], outw[1 % S]);
            else {
#pragma unroll
                for (int w = 0; w < S; ++w) out_words[row * S + w] = outw[w];
            }
            out_flags[row] = (uint8_t)flags;
        }
    }
}

template <int S, bool BUCKETED>
__global__ void __launch_bounds__(kBlock) find_packed_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags,
                                                             uint64_t nq, IndexView ix, int64_t *__restrict__ out_index) {
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * kBlock) {
        uint64_t q[S];
        load_key<S>(words, i, q);
        int64_t r = -1;
        const bool skip = flags && (flags[i] & 6u);
        if (!skip) r = BUCKETED ? lookup_bucketed<S>(ix, q) : lookup_bsearch<S>(ix, q);
        out_index[i] = r;
    }
}

// The bucketed search with Q queries in flight per thread: all table reads are issued, then the first kProbe keys of
// every bucket (independent loads), and only buckets longer than kProbe (rare at ~1 key per bucket) take the general
// search on their remainder.  Lookups are latency-bound (two dependent random sectors each), so throughput follows the
// number of independent loads in flight.
constexpr uint32_t kProbe = 4;

// Q queries per thread: table reads for all, then the first kProbe keys of every bucket, then compare.
template <int S, int Q>
__device__ __forceinline__ void lookup_mlp(const IndexView &ix, const uint64_t (&q)[Q][S], const bool (&live)[Q], int64_t (&r)[Q]) {
    uint32_t lo[Q], hi[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        lo[j] = hi[j] = 0;
        if (live[j]) {
            const uint32_t p = key_prefix<S>(q[j], ix.shift);
            lo[j] = __ldg(ix.table + p);
            hi[j] = __ldg(ix.table + p + 1);
        }
    }
    uint64_t kk[Q][kProbe][S];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
#pragma unroll
        for (uint32_t t = 0; t < kProbe; ++t) {
            if (live[j] && lo[j] + t < hi[j]) load_key<S>(ix.keys, (uint64_t)lo[j] + t, kk[j][t]);
            else {
#pragma unroll
                for (int w = 0; w < S; ++w) kk[j][t][w] = ~0ull;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        r[j] = -1;
        if (!live[j]) continue;
#pragma unroll
        for (int t = (int)kProbe - 1; t >= 0; --t)
            if (lo[j] + (uint32_t)t < hi[j] && words_equal<S>(kk[j][t], q[j])) r[j] = (int64_t)lo[j] + t;
        if (r[j] < 0 && hi[j] - lo[j] > kProbe) r[j] = search_range<S>(ix.keys, (uint64_t)lo[j] + kProbe, hi[j], q[j]);
        if (r[j] >= 0) r[j] += (int64_t)ix.first_index;
    }
}

template <int S, int Q>
__global__ void __launch_bounds__(kBlock) find_packed_mlp_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags,
                                                                 uint64_t nq, IndexView ix, int64_t *__restrict__ out_index) {
    const uint64_t span = (uint64_t)gridDim.x * kBlock;
    for (uint64_t base = (uint64_t)blockIdx.x * kBlock + threadIdx.x; base < nq; base += span * Q) {
        uint64_t q[Q][S];
        bool live[Q];
        int64_t r[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const uint64_t i = base + (uint64_t)j * span;
            live[j] = i < nq;
            if (live[j]) {
                load_key<S>(words, i, q[j]);
                if (flags && (flags[i] & 6u)) live[j] = false, out_index[i] = -1;
            }
        }
        lookup_mlp<S, Q>(ix, q, live, r);
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j]) out_index[base + (uint64_t)j * span] = r[j];
    }
}

template <int S>
__device__ __forceinline__ uint32_t owner_of(const uint64_t (&q)[S], const uint64_t *__restrict__ splitters, int nshards) {
    // number of splitters <= q  (splitter j = first key of shard j+1)
    uint32_t lo = 0, hi = (uint32_t)(nshards - 1);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        uint64_t sp[S];
#pragma unroll
        for (int w = 0; w < S; ++w) sp[w] = splitters[mid * S + w];
        if (words_less<S>(q, sp)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ------------------------------------------------------------------ multi-GPU: routed lookups over peer memory
// One kernel per leg, each fused with its transfer (no NCCL on the data path; buffers are peer-mapped over NVLink):
//   route   owner of every query (splitter search) + block-aggregated reservation + P2P STORE of the key into the
//           owner's inbox segment reserved for this rank; the original slot is kept locally
//   search  the owner walks all inbox segments, searches its shard and P2P-STORES each result into the origin's
//           return buffer at the same segment position
//   gather  the origin scatters the returned indices to the original slots (local)
constexpr int kRouteQ = 8;                   // queries per thread per tile of the route kernel

struct PeerPtrs {
    void *p[kMaxShards];
};

// Tile of BLOCK*kRouteQ queries per block iteration.  The tile is regrouped by owner in shared memory so that each
// owner's run leaves as fully coalesced stores (whole 128-byte lines over NVLink instead of 16-byte fragments).
template <int S, int BLOCK>
__global__ void __launch_bounds__(BLOCK) route_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags, uint64_t nq,
                                                       const uint64_t *__restrict__ splitters, int nshards, int my_rank, uint64_t cap,
                                                       PeerPtrs inbox, uint32_t *__restrict__ slots, unsigned long long *cursors,
                                                       int64_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t route_smem[];
    constexpr uint32_t tile_q = BLOCK * kRouteQ;
    uint64_t *stage_keys = reinterpret_cast<uint64_t *>(route_smem);                       // [tile_q][S], grouped by owner
    uint32_t *stage_slot = reinterpret_cast<uint32_t *>(route_smem + (size_t)tile_q * S * 8);   // [tile_q]
    __shared__ uint32_t hist[kMaxShards], loc[kMaxShards + 1];
    __shared__ unsigned long long base[kMaxShards];
    __shared__ uint64_t spl[(kMaxShards - 1) * S];
    for (int i = threadIdx.x; i < (nshards - 1) * S; i += BLOCK) spl[i] = splitters[i];
    const uint64_t ntiles = (nq + tile_q - 1) / tile_q;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < nshards; i += BLOCK) hist[i] = 0;
        __syncthreads();
        uint64_t q[kRouteQ][S];
        uint32_t own[kRouteQ], rank_in[kRouteQ];
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint64_t i = tile * tile_q + (uint64_t)j * BLOCK + threadIdx.x;
            own[j] = 0xffffffffu;
            if (i < nq) {
                if (flags && (flags[i] & 6u)) out[i] = -1;           // never routed: cannot match
                else {
                    load_key<S>(words, i, q[j]);
                    own[j] = owner_of<S>(q[j], spl, nshards);
                }
            }
        }
        // warp-aggregated ranking: one shared-memory atomic per (warp, owner) instead of one per query -- with few
        // owners the per-query atomics all hit the same two or eight addresses and serialise
        const uint32_t lane = threadIdx.x & 31u, lane_lt = (1u << lane) - 1u;
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint32_t peers = __match_any_sync(0xffffffffu, own[j]);
            const uint32_t leader = __ffs(peers) - 1u;
            uint32_t b = 0;
            if (lane == leader && own[j] != 0xffffffffu) b = atomicAdd(&hist[own[j]], (uint32_t)__popc(peers));
            b = __shfl_sync(0xffffffffu, b, leader);
            rank_in[j] = b + __popc(peers & lane_lt);
        }
        __syncthreads();
        if (threadIdx.x < (uint32_t)nshards)
            base[threadIdx.x] = hist[threadIdx.x] ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)hist[threadIdx.x]) : 0ull;
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (int i = 0; i < nshards; ++i) { loc[i] = acc; acc += hist[i]; }
            loc[nshards] = acc;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            if (own[j] == 0xffffffffu) continue;
            const uint32_t at = loc[own[j]] + rank_in[j];
#pragma unroll
            for (int w = 0; w < S; ++w) stage_keys[(size_t)at * S + w] = q[j][w];
            stage_slot[at] = (uint32_t)(tile * tile_q + (uint64_t)j * BLOCK + threadIdx.x);
        }
        __syncthreads();
        for (int o = 0; o < nshards; ++o) {
            const uint32_t cnt = hist[o];
            if (cnt == 0) continue;
            const uint64_t b0 = base[o];
            const uint32_t keep = b0 >= cap ? 0u : (uint32_t)min((unsigned long long)cnt, (unsigned long long)(cap - b0));
            uint64_t *dst = static_cast<uint64_t *>(inbox.p[o]) + ((uint64_t)my_rank * cap + b0) * S;
            const uint64_t *src = stage_keys + (size_t)loc[o] * S;
            for (uint32_t t = threadIdx.x; t < keep * S; t += BLOCK) dst[t] = src[t];
            uint32_t *sdst = slots + (uint64_t)o * cap + b0;
            for (uint32_t t = threadIdx.x; t < keep; t += BLOCK) sdst[t] = stage_slot[loc[o] + t];
        }
        __syncthreads();
    }
    __threadfence_system();
}

// counts_in[my_rank] on every owner := number of keys this rank routed to it (P2P stores of 8 bytes)
__global__ void publish_counts_kernel(const unsigned long long *cursors, int nshards, int my_rank, uint64_t cap, PeerPtrs counts_in) {
    const int o = threadIdx.x;
    if (o < nshards) {
        const unsigned long long c = cursors[o] < cap ? cursors[o] : cap;
        static_cast<unsigned long long *>(counts_in.p[o])[my_rank] = c;
    }
    __threadfence_system();
}

template <int S, int Q>
__global__ void __launch_bounds__(kBlock) find_routed_kernel(const uint64_t *__restrict__ inbox, const unsigned long long *__restrict__ counts_in,
                                                             int nshards, int my_rank, uint64_t cap, IndexView ix, PeerPtrs ret) {
    __shared__ unsigned long long pre[kMaxShards + 1];
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < nshards; ++i) { pre[i] = acc; acc += counts_in[i]; }
        pre[nshards] = acc;
    }
    __syncthreads();
    const uint64_t total = pre[nshards];
    const uint64_t span = (uint64_t)gridDim.x * kBlock;
    for (uint64_t basei = (uint64_t)blockIdx.x * kBlock + threadIdx.x; basei < total; basei += span * Q) {
        uint64_t q[Q][S];
        bool live[Q];
        int64_t r[Q];
        uint32_t src[Q];
        uint64_t pos[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const uint64_t f = basei + (uint64_t)j * span;
            live[j] = f < total;
            src[j] = 0; pos[j] = 0;
            if (live[j]) {
                uint32_t sgm = 0;
                while (sgm + 1 < (uint32_t)nshards && f >= pre[sgm + 1]) ++sgm;
                src[j] = sgm;
                pos[j] = f - pre[sgm];
                load_key<S>(inbox, (uint64_t)sgm * cap + pos[j], q[j]);
            }
        }
        lookup_mlp<S, Q>(ix, q, live, r);
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (live[j]) static_cast<int64_t *>(ret.p[src[j]])[(uint64_t)my_rank * cap + pos[j]] = r[j];
    }
    __threadfence_system();
}

__global__ void gather_routed_kernel(const int64_t *__restrict__ ret, const uint32_t *__restrict__ slots,
                                     const unsigned long long *__restrict__ sent, int nshards, uint64_t cap,
                                     int64_t *__restrict__ out) {
    for (int o = 0; o < nshards; ++o) {
        const uint64_t n = sent[o] < cap ? sent[o] : cap;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
            out[slots[(uint64_t)o * cap + i]] = ret[(uint64_t)o * cap + i];
    }
}

// ------------------------------------------------------------------ index construction
template <int S>
__global__ void check_sorted_kernel(const uint64_t *__restrict__ keys, uint64_t n, unsigned long long *unsorted_at) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t a[S], b[S];
        load_key<S>(keys, i - 1, a);
        load_key<S>(keys, i, b);
        if (words_less<S>(b, a)) atomicMin(unsorted_at, (unsigned long long)i);
    }
}

// table[b] = number of keys whose prefix is < b  (b in [0, 2^bits]).
template <int S>
__global__ void build_table_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t bits, uint32_t shift, uint32_t *table) {
    const uint64_t nb = 1ull << bits;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t prev = -1, cur = (int64_t)nb;
        if (i > 0) { uint64_t a[S]; load_key<S>(keys, i - 1, a); prev = (int64_t)key_prefix<S>(a, shift); }
        if (i < n) { uint64_t b[S]; load_key<S>(keys, i, b); cur = (int64_t)key_prefix<S>(b, shift); }
        for (int64_t b = prev + 1; b <= cur; ++b) table[b] = (uint32_t)i;
    }
}

// ------------------------------------------------------------------ sorted-merge support
template <int S>
__global__ void extract_word_kernel(const uint64_t *__restrict__ words, const uint32_t *__restrict__ perm, uint64_t nq, int w,
                                    uint64_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t src = perm ? perm[i] : i;
        out[i] = words[src * S + w];
    }
}
__global__ void iota_kernel(uint32_t *p, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
// Sorted pass: thread i resolves query perm[i]; neighbouring threads probe neighbouring keys.
template <int S>
__global__ void __launch_bounds__(kBlock) find_sorted_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags,
                                                             const uint32_t *__restrict__ perm, uint64_t nq, IndexView ix,
                                                             int64_t *__restrict__ out_index) {
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * kBlock) {
        const uint64_t src = perm[i];
        uint64_t q[S];
        load_key<S>(words, src, q);
        int64_t r = -1;
        if (!(flags && (flags[src] & 6u))) r = lookup_bucketed<S>(ix, q);
        out_index[src] = r;
    }
}

// ------------------------------------------------------------------ multi-GPU helpers
template <int S>
__global__ void __launch_bounds__(kBlock) owner_count_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags, uint64_t nq,
                                                             const uint64_t *__restrict__ splitters, int nshards,
                                                             unsigned long long *counts) {
    __shared__ uint32_t hist[kMaxShards];
    for (int i = threadIdx.x; i < nshards; i += kBlock) hist[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * kBlock) {
        if (flags && (flags[i] & 6u)) continue;
        uint64_t q[S];
        load_key<S>(words, i, q);
        atomicAdd(&hist[owner_of<S>(q, splitters, nshards)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nshards; i += kBlock) if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

// cursors[j] starts at the exclusive prefix of counts; each block reserves its share per owner.
template <int S>
__global__ void __launch_bounds__(kBlock) owner_scatter_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags, uint64_t nq,
                                                               const uint64_t *__restrict__ splitters, int nshards,
                                                               unsigned long long *cursors, uint64_t *__restrict__ sorted_words,
                                                               uint32_t *__restrict__ slots) {
    __shared__ uint32_t hist[kMaxShards];
    __shared__ unsigned long long base[kMaxShards];
    const uint64_t per_block = (nq + gridDim.x - 1) / gridDim.x;
    const uint64_t b0 = (uint64_t)blockIdx.x * per_block, b1 = min(nq, b0 + per_block);
    for (int i = threadIdx.x; i < nshards; i += kBlock) hist[i] = 0;
    __syncthreads();
    for (uint64_t i = b0 + threadIdx.x; i < b1; i += kBlock) {
        if (flags && (flags[i] & 6u)) continue;
        uint64_t q[S];
        load_key<S>(words, i, q);
        atomicAdd(&hist[owner_of<S>(q, splitters, nshards)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nshards; i += kBlock) {
        base[i] = hist[i] ? atomicAdd(&cursors[i], (unsigned long long)hist[i]) : 0ull;
        hist[i] = 0;
    }
    __syncthreads();
    for (uint64_t i = b0 + threadIdx.x; i < b1; i += kBlock) {
        if (flags && (flags[i] & 6u)) continue;
        uint64_t q[S];
        load_key<S>(words, i, q);
        const uint32_t o = owner_of<S>(q, splitters, nshards);
        const uint64_t pos = base[o] + atomicAdd(&hist[o], 1u);
#pragma unroll
        for (int w = 0; w < S; ++w) sorted_words[pos * S + w] = q[w];
        slots[pos] = (uint32_t)i;
    }
}

__global__ void exclusive_scan_small_kernel(const unsigned long long *counts, int n, unsigned long long *cursors) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < n; ++i) { cursors[i] = acc; acc += counts[i]; }
    }
}

__global__ void scatter_results_kernel(const int64_t *__restrict__ values, const uint32_t *__restrict__ slots, uint64_t n, int64_t *__restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[slots[i]] = values[i];
}

// ------------------------------------------------------------------ host helpers
int grid_for(uint64_t work_items, int per_block, int sm_count, int blocks_per_sm) {
    uint64_t blocks = (work_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)sm_count * blocks_per_sm;
    return (int)std::max<uint64_t>(1, std::min(blocks, cap));
}

int sm_count_now() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

IndexView view_of(const cc_graph *g) {
    IndexView v{};
    v.keys = g->index.keys;
    v.table = g->index.table;
    v.n = g->h.num_records;
    v.first_index = g->first_index;
    v.bits = (uint32_t)g->index.bits;
    v.shift = 2u * g->h.k - (uint32_t)g->index.bits;
    return v;
}

#define CC_DISPATCH_S(s, ...)                                                             \
    switch (s) {                                                                          \
        case 1: { constexpr int S_ = 1; __VA_ARGS__; break; }                             \
        case 2: { constexpr int S_ = 2; __VA_ARGS__; break; }                             \
        case 3: { constexpr int S_ = 3; __VA_ARGS__; break; }                             \
        case 4: { constexpr int S_ = 4; __VA_ARGS__; break; }                             \
        default: return fail(CC_ERR_UNSUPPORTED, "k-mers wider than 4 words (k > 128) are not supported by pack/lookup"); \
    }

int check_k(uint32_t k) {
    if (k == 0) return fail(CC_ERR_ARG, "k must be positive");
    if (k > 128) return fail(CC_ERR_UNSUPPORTED, "k-mers wider than 4 words (k > 128) are not supported by pack/lookup");
    if (k > kTileBytes / 4) return fail(CC_ERR_UNSUPPORTED, "k too large for the sequence tile");
    return CC_OK;
}

SeqJob make_job(const uint8_t *seq, uint64_t nq, uint64_t stride, uint32_t k) {
    SeqJob j{};
    j.seq = seq; j.nq = nq; j.stride = stride; j.k = k;
    j.per_tile = (uint32_t)((kTileBytes - k) / stride + 1);
    return j;
}

}  // namespace

// ------------------------------------------------------------------ launchers
int launch_pack_windows(const uint8_t *dev_seq, uint64_t /*len*/, uint32_t k, uint64_t *dev_words, uint8_t *dev_flags,
                        uint64_t row_stride, uint64_t nq, cudaStream_t st) {
    if (int rc = check_k(k)) return rc;
    if (nq == 0) return CC_OK;
    const uint32_t s = (k + 31) / 32;
    if (row_stride == k && row_stride > 1) {       // independent rows
        const size_t smem = (size_t)(kBlock / 32) * ((32u * k + 32u + 15u) & ~15u);
        const int per_sm = std::max(1, std::min(8, (int)((200u << 10) / (smem + 1024))));
        const int grid = grid_for((nq + 31) / 32, kBlock / 32, sm_count_now(), per_sm);
        CC_DISPATCH_S(s, {
            CC_CUDA(cudaFuncSetAttribute(pack_rows_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pack_rows_kernel<S_><<<grid, kBlock, smem, st>>>(dev_seq, nq, k, dev_words, dev_flags);
        });
        count_launch();
        CC_CUDA(cudaGetLastError());
        return CC_OK;
    }
    SeqJob job = make_job(dev_seq, nq, row_stride, k);
    const int grid = grid_for((nq + job.per_tile - 1) / job.per_tile, 1, sm_count_now(), 6);
    IndexView none{};
    CC_DISPATCH_S(s, seq_kernel<S_, false, false><<<grid, kBlock, 0, st>>>(job, dev_words, dev_flags, none, nullptr));
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_find_seq(cc_graph *g, const uint8_t *dev_seq, uint64_t /*len*/, uint64_t row_stride, uint64_t nq, int64_t *dev_index,
                    int algo, cudaStream_t st) {
    if (int rc = check_k(g->h.k)) return rc;
    if (nq == 0) return CC_OK;
    if (algo == CC_ALGO_MERGE) {
        // sort needs materialised words: pack, then the packed path
        uint64_t *words = nullptr; uint8_t *flags = nullptr;
        CC_CUDA(cudaMallocAsync(&words, nq * g->h.s * sizeof(uint64_t), st));
        CC_CUDA(cudaMallocAsync(&flags, nq, st));
        int rc = launch_pack_windows(dev_seq, 0, g->h.k, words, flags, row_stride, nq, st);
        if (!rc) rc = launch_find_packed(g, words, flags, nq, dev_index, algo, st);
        cudaFreeAsync(words, st); cudaFreeAsync(flags, st);
        return rc;
    }
    if (row_stride > 1 && nq >= (1u << 16)) {
        // Independent rows cost k bytes of staging per query, which leaves the fused kernel at half occupancy for the
        // latency-bound search; for large batches pack first (streaming), then search at full occupancy.
        uint64_t *words = nullptr; uint8_t *flags = nullptr;
        CC_CUDA(cudaMallocAsync(&words, nq * g->h.s * sizeof(uint64_t), st));
        CC_CUDA(cudaMallocAsync(&flags, nq, st));
        int rc = launch_pack_windows(dev_seq, 0, g->h.k, words, flags, row_stride, nq, st);
        if (!rc) rc = launch_find_packed(g, words, flags, nq, dev_index, algo, st);
        cudaFreeAsync(words, st); cudaFreeAsync(flags, st);
        return rc;
    }
    SeqJob job = make_job(dev_seq, nq, row_stride, g->h.k);
    const int grid = grid_for((nq + job.per_tile - 1) / job.per_tile, 1, g->sm_count, 6);
    IndexView ix = view_of(g);
    if (algo == CC_ALGO_BSEARCH) {
        CC_DISPATCH_S(g->h.s, seq_kernel<S_, true, false><<<grid, kBlock, 0, st>>>(job, nullptr, nullptr, ix, dev_index));
    } else {
        CC_DISPATCH_S(g->h.s, seq_kernel<S_, true, true><<<grid, kBlock, 0, st>>>(job, nullptr, nullptr, ix, dev_index));
    }
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// Stable ascending order of n multi-word keys: LSD radix sort by (word s-1 ... word 0), carrying the permutation.
// *perm_out is stream-ordered memory (cudaFreeAsync).
int sort_permutation(const uint64_t *dev_words, uint64_t n, uint32_t s, uint32_t k, cudaStream_t st, uint32_t **perm_out) {
    if (n >= (1ull << 32)) return fail(CC_ERR_UNSUPPORTED, "sorting is limited to 2^32-1 keys");
    uint32_t *perm_a = nullptr, *perm_b = nullptr; uint64_t *key_a = nullptr, *key_b = nullptr; void *tmp = nullptr;
    size_t tmp_bytes = 0;
    CC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_a, key_b, perm_a, perm_b, (int64_t)n, 0, 64, st));
    CC_CUDA(cudaMallocAsync(&perm_a, std::max<uint64_t>(n, 1) * 4, st)); CC_CUDA(cudaMallocAsync(&perm_b, std::max<uint64_t>(n, 1) * 4, st));
    CC_CUDA(cudaMallocAsync(&key_a, std::max<uint64_t>(n, 1) * 8, st)); CC_CUDA(cudaMallocAsync(&key_b, std::max<uint64_t>(n, 1) * 8, st));
    CC_CUDA(cudaMallocAsync(&tmp, tmp_bytes, st));
    const int grid = grid_for(std::max<uint64_t>(n, 1), kBlock, sm_count_now(), 8);
    iota_kernel<<<grid, kBlock, 0, st>>>(perm_a, n); count_launch();
    const uint32_t top_bits = 2u * k - 64u * (s - 1);
    int rc = CC_OK;
    for (int w = (int)s - 1; w >= 0 && rc == CC_OK && n > 0; --w) {
        CC_DISPATCH_S(s, extract_word_kernel<S_><<<grid, kBlock, 0, st>>>(dev_words, perm_a, n, w, key_a));
        count_launch();
        const int end_bit = (w == 0) ? (int)top_bits : 64;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_a, key_b, perm_a, perm_b, (int64_t)n, 0, end_bit, st);
        count_launch(2 * ((end_bit + 7) / 8));
        if (e != cudaSuccess) rc = cuda_fail(e, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
        std::swap(perm_a, perm_b);
    }
    cudaFreeAsync(perm_b, st); cudaFreeAsync(key_a, st); cudaFreeAsync(key_b, st); cudaFreeAsync(tmp, st);
    if (rc) { cudaFreeAsync(perm_a, st); return rc; }
    *perm_out = perm_a;
    return CC_OK;
}

int launch_find_packed(cc_graph *g, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, int64_t *dev_index,
                       int algo, cudaStream_t st) {
    if (int rc = check_k(g->h.k)) return rc;
    if (nq == 0) return CC_OK;
    IndexView ix = view_of(g);
    const uint32_t s = g->h.s;
    const int grid = grid_for(nq, kBlock, g->sm_count, 8);
    if (algo == CC_ALGO_MERGE) {
        uint32_t *perm = nullptr;
        if (int rc = sort_permutation(dev_words, nq, s, g->h.k, st, &perm)) return rc;
        CC_DISPATCH_S(s, find_sorted_kernel<S_><<<grid, kBlock, 0, st>>>(dev_words, dev_flags, perm, nq, ix, dev_index));
        count_launch();
        cudaFreeAsync(perm, st);
    } else if (algo == CC_ALGO_BSEARCH) {
        CC_DISPATCH_S(s, find_packed_kernel<S_, false><<<grid, kBlock, 0, st>>>(dev_words, dev_flags, nq, ix, dev_index));
        count_launch();
    } else {
        const int qpt = options().lookup_queries_per_thread;
        const int g2 = grid_for((nq + std::max(qpt, 1) - 1) / std::max(qpt, 1), kBlock, g->sm_count, 8);
        if (qpt >= 4) { CC_DISPATCH_S(s, find_packed_mlp_kernel<S_, 4><<<g2, kBlock, 0, st>>>(dev_words, dev_flags, nq, ix, dev_index)); }
        else if (qpt >= 2) { CC_DISPATCH_S(s, find_packed_mlp_kernel<S_, 2><<<g2, kBlock, 0, st>>>(dev_words, dev_flags, nq, ix, dev_index)); }
        else if (qpt == 1) { CC_DISPATCH_S(s, find_packed_mlp_kernel<S_, 1><<<g2, kBlock, 0, st>>>(dev_words, dev_flags, nq, ix, dev_index)); }
        else { CC_DISPATCH_S(s, find_packed_kernel<S_, true><<<grid, kBlock, 0, st>>>(dev_words, dev_flags, nq, ix, dev_index)); }
        count_launch();
    }
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int build_index(cc_graph *g, int bits_req) {
    if (int rc = check_k(g->h.k)) return rc;
    const uint64_t n = g->h.num_records;
    const uint32_t s = g->h.s, k = g->h.k;
    if (n >= 0xffffffffull) return fail(CC_ERR_UNSUPPORTED, "more than 2^32-2 records per device shard");
    LookupIndex &ix = g->index;
    if (ix.keys) { cudaFree(ix.keys); ix.keys = nullptr; }
    if (ix.table) { cudaFree(ix.table); ix.table = nullptr; }
    ix.built = false;
    cudaStream_t st = g->stream;

    int bits = bits_req > 0 ? bits_req : options().index_bits;
    if (bits <= 0) {
        // about one key per bucket: a lookup is then one table sector plus one key sector (measured on B200 at
        // n = 1e8: 2^24 buckets 1.2e10 lookups/s, 2^26 1.9e10; the table costs 4 bytes per bucket next to 8s per key)
        int lg = 0;
        while ((1ull << lg) < std::max<uint64_t>(n, 1)) ++lg;
        bits = std::max(1, lg);
    }
    bits = std::min<int>(bits, (int)std::min<uint32_t>(2u * k, 30u));
    ix.bits = bits;

    if (int rc = g->scan_ws.ensure(0, 0)) return rc;
    CC_CUDA(cudaMalloc(&ix.keys, std::max<uint64_t>(n * s, 2) * sizeof(uint64_t) + 64));
    CC_CUDA(cudaMalloc(&ix.table, ((1ull << bits) + 1) * sizeof(uint32_t)));
    if (int rc = launch_decode_columns(g->dev_body, n, s, g->h.c, ix.keys, nullptr, nullptr, g->scan_ws, g->sm_count, st)) return rc;

    unsigned long long *d_unsorted = reinterpret_cast<unsigned long long *>(g->scan_ws.totals + 8);
    const unsigned long long none = ~0ull;
    CC_CUDA(cudaMemcpyAsync(d_unsorted, &none, 8, cudaMemcpyHostToDevice, st));
    const int grid = grid_for(n + 1, 256, g->sm_count, 8);
    const uint32_t shift = 2u * k - (uint32_t)bits;
    CC_DISPATCH_S(s, check_sorted_kernel<S_><<<grid, 256, 0, st>>>(ix.keys, n, d_unsorted));
    count_launch();
    CC_DISPATCH_S(s, build_table_kernel<S_><<<grid, 256, 0, st>>>(ix.keys, n, (uint32_t)bits, shift, ix.table));
    count_launch();
    CC_CUDA(cudaGetLastError());
    unsigned long long at = none;
    CC_CUDA(cudaMemcpyAsync(&at, d_unsorted, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    ix.sorted = (at == none);
    ix.unsorted_at = at;
    ix.built = true;
    return CC_OK;
}

int launch_bucket_by_owner(const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t s,
                           const uint64_t *dev_splitters, int nshards, uint64_t *dev_counts, uint64_t *dev_sorted_words,
                           uint32_t *dev_slots, cudaStream_t st) {
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    if (nq >= (1ull << 32)) return fail(CC_ERR_UNSUPPORTED, "bucket batches are limited to 2^32-1 queries");
    CC_CUDA(cudaMemsetAsync(dev_counts, 0, sizeof(uint64_t) * nshards, st));
    if (nq == 0) return CC_OK;
    unsigned long long *cursors = nullptr;
    CC_CUDA(cudaMallocAsync(&cursors, sizeof(unsigned long long) * kMaxShards, st));
    const int sms = sm_count_now();
    const int grid = grid_for(nq, kBlock * 8, sms, 8);
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(dev_counts);
    CC_DISPATCH_S(s, owner_count_kernel<S_><<<grid, kBlock, 0, st>>>(dev_words, dev_flags, nq, dev_splitters, nshards, counts));
    exclusive_scan_small_kernel<<<1, 32, 0, st>>>(counts, nshards, cursors);
    CC_DISPATCH_S(s, owner_scatter_kernel<S_><<<grid, kBlock, 0, st>>>(dev_words, dev_flags, nq, dev_splitters, nshards, cursors,
                                                                         dev_sorted_words, dev_slots));
    count_launch(3);
    cudaFreeAsync(cursors, st);
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_scatter_results(const int64_t *dev_values, const uint32_t *dev_slots, uint64_t n, int64_t *dev_out, cudaStream_t st) {
    if (n == 0) return CC_OK;
    const int grid = grid_for(n, 256, sm_count_now(), 8);
    scatter_results_kernel<<<grid, 256, 0, st>>>(dev_values, dev_slots, n, dev_out);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// ------------------------------------------------------------------ routed lookups (peer memory) launchers
int launch_route(const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t s, const uint64_t *dev_splitters, int nshards,
                 int my_rank, uint64_t cap, void *const *peer_inbox, void *const *peer_counts, uint32_t *dev_slots, uint64_t *dev_sent,
                 int64_t *dev_out, cudaStream_t st) {
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    if (my_rank < 0 || my_rank >= nshards) return fail(CC_ERR_ARG, "rank %d out of range", my_rank);
    if (nq >= (1ull << 32)) return fail(CC_ERR_UNSUPPORTED, "routed batches are limited to 2^32-1 queries per rank");
    PeerPtrs inbox{}, counts{};
    for (int i = 0; i < nshards; ++i) { inbox.p[i] = peer_inbox[i]; counts.p[i] = peer_counts[i]; }
    unsigned long long *cursors = reinterpret_cast<unsigned long long *>(dev_sent);
    CC_CUDA(cudaMemsetAsync(cursors, 0, sizeof(uint64_t) * nshards, st));
    if (nq) {
        // small blocks (128 threads) when the leg is to be overlapped with the search on another stream: one of them
        // fits next to three resident search blocks
        const bool small = options().route_blocks_per_sm > 0;
        const int block = small ? 128 : kBlock;
        const size_t smem = (size_t)block * kRouteQ * (8 * s + 4);
        int per_sm = std::max(1, std::min(4, (int)((200u << 10) / (smem + 4096))));
        if (small) per_sm = std::min(per_sm, options().route_blocks_per_sm);
        const int grid = grid_for((nq + (uint64_t)block * kRouteQ - 1) / ((uint64_t)block * kRouteQ), 1, sm_count_now(), per_sm);
        if (small) {
            CC_DISPATCH_S(s, {
                CC_CUDA(cudaFuncSetAttribute(route_kernel<S_, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                route_kernel<S_, 128><<<grid, 128, smem, st>>>(dev_words, dev_flags, nq, dev_splitters, nshards, my_rank, cap, inbox,
                                                                dev_slots, cursors, dev_out);
            });
        } else {
            CC_DISPATCH_S(s, {
                CC_CUDA(cudaFuncSetAttribute(route_kernel<S_, kBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                route_kernel<S_, kBlock><<<grid, kBlock, smem, st>>>(dev_words, dev_flags, nq, dev_splitters, nshards, my_rank, cap, inbox,
                                                                      dev_slots, cursors, dev_out);
            });
        }
        count_launch();
    }
    publish_counts_kernel<<<1, kMaxShards, 0, st>>>(cursors, nshards, my_rank, cap, counts);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_find_routed(cc_graph *g, const uint64_t *dev_inbox, const uint64_t *dev_counts_in, int nshards, int my_rank, uint64_t cap,
                       void *const *peer_ret, cudaStream_t st) {
    if (int rc = check_k(g->h.k)) return rc;
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    PeerPtrs ret{};
    for (int i = 0; i < nshards; ++i) ret.p[i] = peer_ret[i];
    IndexView ix = view_of(g);
    const int grid = g->sm_count * std::max(1, options().routed_search_blocks_per_sm);
    CC_DISPATCH_S(g->h.s, find_routed_kernel<S_, 2><<<grid, kBlock, 0, st>>>(dev_inbox, reinterpret_cast<const unsigned long long *>(dev_counts_in),
                                                                             nshards, my_rank, cap, ix, ret));
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_gather_routed(const int64_t *dev_ret, const uint32_t *dev_slots, const uint64_t *dev_sent, int nshards, uint64_t cap,
                         int64_t *dev_out, cudaStream_t st) {
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    const bool small = options().gather_blocks_per_sm > 0 && options().gather_blocks_per_sm < 8;
    gather_routed_kernel<<<sm_count_now() * std::max(1, options().gather_blocks_per_sm), small ? 128 : kBlock, 0, st>>>(dev_ret, dev_slots, reinterpret_cast<const unsigned long long *>(dev_sent),
                                                                nshards, cap, dev_out);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
