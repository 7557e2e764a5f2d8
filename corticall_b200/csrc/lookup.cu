// lookup.cu -- K3 (canonicalise + 2-bit pack), K4 (batched sorted-array lookups) and the multi-GPU legs around K4.
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
// Sections (in file order):
//   tile staging + window_canonical      sliding windows of a sequence (seq_kernel): 2-bit stream + masks in shared memory; mixed
//                                        case decided on packed codes + case bits
//   the lookup index                     64-byte bucket lines (LineLayout, IndexView, key_bin / key_line), search by lane pairs with
//                                        256-bit loads (lookup_lines_warp), per-thread form, overflow into the key column
//   seq_kernel / find kernels            K3 on windows (optionally fused with K4), K4 on packed queries (find_packed_lines_kernel)
//   rows_kernel                          K3 (+K4) on independent k-byte rows: TMA-staged tiles, byte-parallel conversion,
//                                        warp-cooperative exact path for rows with N / lower case; find_small_kernel (findRecord)
//   routed lookups over peer memory      route_kernel (owner + bulk copies into the owners' inboxes), find_routed_kernel,
//                                        gather_routed_kernel (P2P pull)
//   index construction, sort support, NCCL-formulation helpers (bucket / scatter), launchers
//
// B200 design in one paragraph.  K3: the sequence is converted once per tile into a 2-bit big-endian bit stream in shared
// memory and every thread cuts its k-mer out with funnel shifts, reverse-complements in registers (brev + pair swap +
// multiword shift) and keeps the smaller.  K4: HBM serves about 3.9e10 random lines per second whatever their size up to 128
// bytes (tools/micro/randline.cu), so a lookup must cost ONE random access: the sorted keys are laid out a second time as an
// order-preserving table of 64-byte bucket lines addressed through a small bin table in shared memory, and two lanes share a
// line (one 256-bit load each); the kernel runs at 86 % of that access rate.
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
    uint32_t dirty[kStreamBases / 32 + 4];     // inval | lower: one test clears the common window (upper-case ACGT only)
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
    // streams: group g covers stream positions [32g, 32g+32); position p holds tile byte p - kPadBases.  Four bases per 32-bit
    // operation: code = ((u>>1)^(u>>2))&3 on the upper-cased byte u, the letter rebuilt from its code with PRMT (a byte that
    // differs is not in ACGTacgt), the four 2-bit codes and the four flag bits gathered by one multiply each.
    const uint32_t ngroups = (nb + kPadBases + 31) / 32 + 2;
    for (uint32_t g = threadIdx.x; g < ngroups; g += blockDim.x) {
        uint32_t cw[2] = {0u, 0u}, inv = 0, low = 0;
#pragma unroll
        for (uint32_t kq = 0; kq < 8; ++kq) {
            const int32_t pw = (int32_t)(g * 32 + 4 * kq) - (int32_t)kPadBases;       // tile byte of the word's first base (multiple of 4)
            if (pw < 0 || pw >= (int32_t)nb) continue;                                 // outside the tile: clean zeros
            uint32_t keep = 0x01010101u;                                               // bytes inside [0, nb)
            if (pw + 4 > (int32_t)nb) keep = 0x01010101u >> (8 * (uint32_t)(pw + 4 - (int32_t)nb));
            const uint32_t wv = lds_u32<false>(dst + pw);
            const uint32_t u = wv & 0xdfdfdfdfu;
            const uint32_t x = ((u >> 1) ^ (u >> 2)) & 0x03030303u;
            const uint32_t y = x | (x >> 4);
            const uint32_t z = y & 0x00ff00ffu;
            const uint32_t sel = (z | (z >> 8)) & 0xffffu;
            const uint32_t d = u ^ __byte_perm(0x54474341u, 0u, sel);
            const uint32_t nzb = ((d | ((d & 0x7f7f7f7fu) + 0x7f7f7f7fu)) >> 7) & keep;   // 1 per byte that is not a letter of ACGTacgt
            const uint32_t okb = keep & ~nzb;
            const uint32_t lcb = (wv >> 5) & okb;                                         // 1 per lower-case letter
            const uint32_t code8 = ((x & (okb * 3u)) * 0x40100401u) >> 24;                // c0<<6 | c1<<4 | c2<<2 | c3
            cw[kq >> 2] |= code8 << (24u - 8u * (kq & 3u));
            inv |= ((nzb * 0x80402010u) >> 28) << (28u - 4u * kq);                        // first base = highest bit
            low |= ((lcb * 0x80402010u) >> 28) << (28u - 4u * kq);
        }
        t.codes[2 * g] = cw[0];
        t.codes[2 * g + 1] = cw[1];
        t.inval[g] = inv;
        t.lower[g] = low;
        t.dirty[g] = inv | low;
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

__device__ __forceinline__ uint64_t spread_bits(uint32_t x);

// rc < fw for the canonical choice, as the borrow of the 64*S-bit subtraction rc - fw: one subtract-with-carry chain
// instead of a compare / branch ladder per word (words_less costs ~27 instructions at S = 2, this 5).
template <int S>
__device__ __forceinline__ bool less_by_borrow(const uint64_t (&a)[S], const uint64_t (&b)[S]) {
    uint64_t r;
    if constexpr (S == 1) {
        return a[0] < b[0];
    } else if constexpr (S == 2) {
        asm("{\n .reg .u64 t;\n sub.cc.u64 t, %1, %3;\n subc.cc.u64 t, %2, %4;\n subc.u64 %0, 0, 0;\n}"
            : "=l"(r) : "l"(a[1]), "l"(a[0]), "l"(b[1]), "l"(b[0]));
        return r != 0;
    } else if constexpr (S == 3) {
        asm("{\n .reg .u64 t;\n sub.cc.u64 t, %1, %4;\n subc.cc.u64 t, %2, %5;\n subc.cc.u64 t, %3, %6;\n subc.u64 %0, 0, 0;\n}"
            : "=l"(r) : "l"(a[2]), "l"(a[1]), "l"(a[0]), "l"(b[2]), "l"(b[1]), "l"(b[0]));
        return r != 0;
    } else if constexpr (S == 4) {
        asm("{\n .reg .u64 t;\n sub.cc.u64 t, %1, %5;\n subc.cc.u64 t, %2, %6;\n subc.cc.u64 t, %3, %7;\n subc.cc.u64 t, %4, %8;\n subc.u64 %0, 0, 0;\n}"
            : "=l"(r) : "l"(a[3]), "l"(a[2]), "l"(a[1]), "l"(a[0]), "l"(b[3]), "l"(b[2]), "l"(b[1]), "l"(b[0]));
        return r != 0;
    } else {
        return words_less<S>(a, b);
    }
}

// Canonical packed k-mer of the window starting at tile byte `start`.  Returns flags
// (bit0 flipped, bit1 not ACGTacgt, bit2 has lower case); words zeroed when bit1.
template <int S>
__device__ __forceinline__ uint32_t window_canonical(const SeqTile &t, uint32_t ascii_off, uint32_t start, uint32_t k, uint64_t (&out)[S]) {
    const uint32_t p0 = start + kPadBases;                   // stream position of the window's first base
    const bool dirty = any_bits(t.dirty, p0, k);
    if (dirty && any_bits(t.inval, p0, k)) {
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
    bool flip = less_by_borrow<S>(rc, fw);
    uint32_t flags = 0;
    if (dirty) {
        // Mixed / lower case (soft-masked genomes are half lower case): the reference compares ASCII bytes, seq[i] against
        // complement(seq[k-1-i]), first difference decides (SequenceUtils.java:211-219).  For bytes in ACGTacgt that is the
        // lexicographic order of (case, code) per base -- every upper-case letter sorts before every lower-case one and the
        // 2-bit codes follow the ASCII order within a case; complement keeps the case -- so the comparison is done on the
        // packed words plus the case bits spread to the same 2-bit grid, without touching the bytes again.
        flags |= 4u;
        uint64_t fc[S], nfc[S], rcc[S];
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const uint32_t P = (uint32_t)((int32_t)p0 + (int32_t)k - 32 * (S - j));     // first base of word j in the case stream
            const uint32_t x = __funnelshift_l(t.lower[(P >> 5) + 1], t.lower[P >> 5], P & 31u);
            const uint64_t e = spread_bits(x);
            fc[j] = e | (e << 1);
        }
        if (top_bits < 64) fc[0] &= (1ull << top_bits) - 1ull;
#pragma unroll
        for (int j = 0; j < S; ++j) nfc[j] = ~fc[j];
        revcomp_words<S>(nfc, rcc, k);                      // reverse (the complement inside cancels the ~): case of the reverse strand
        flip = false;
        bool decided = false;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const uint64_t dc = fc[j] ^ rcc[j], d = (fw[j] ^ rc[j]) | dc;
            if (!decided && d != 0) {
                decided = true;
                const uint64_t m = 3ull << ((63u - (uint32_t)__clzll((long long)d)) & ~1u);      // the first base that differs
                flip = (dc & m) ? (fc[j] & m) != 0 : (fw[j] & m) > (rc[j] & m);                    // case first, then the letter
            }
        }
    }
#pragma unroll
    for (int i = 0; i < S; ++i) out[i] = flip ? rc[i] : fw[i];
    return flags | (flip ? 1u : 0u);
}

// ------------------------------------------------------------------ key column access
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

// ------------------------------------------------------------------ the lookup index: 64-byte bucket lines
// A lookup must cost ONE random DRAM access.  The sorted keys are laid out a second time as an order-preserving hash
// table of 64-byte lines (one DRAM atom / two L2 sectors): line = f(key) with f monotone, so a line holds a contiguous run
// of the sorted array and "base index + slot" is the record index.
//   f: the array's OWN key range [first key, last key] is cut into `nbins` equal slices (a k-mer-range shard covers only
//      1/world of the key space); bin b owns bins[b].y consecutive lines starting at bins[b].x, sized from the number of
//      keys that fall into it (lines = ceil(keys / fill)), and a key's line inside its bin is the equal slice of the bin
//      it falls into.  The bin table is what adapts the table to the key density (canonical k-mers are twice as dense at
//      the low end of the key space as on average and vanish at the high end); it is a few KB to 64 KB and lives in shared
//      memory in the search kernels.
//   line: 16 32-bit words = the base index (position of the line's first key in the sorted array) + CAP keys in the wire
//      format (KW = ceil(2k/32) words, most significant first: 12 bytes at k = 47), ascending; unused slots hold the
//      LARGEST key of the array (it sorts after every real key of the line, can only equal a query that maps to the last
//      line, where it is real, and makes "last slot < query" the exact test for "the bucket may continue").
//   overflow: a bucket with more than CAP keys continues in the key column at base + CAP (rare by construction: fill = 50 %
//      of CAP on average); the search reads on from there.
// The word layout inside a line is chosen so that two lanes, each holding one 32-byte half of the line, can compare without
// moving key words between them (LineLayout).
constexpr uint32_t kLineWords = 16;
constexpr uint32_t kMaxBinsLog2 = 13;          // 8192 bins x 8 bytes = 64 KB of shared memory
constexpr uint32_t kMiss32 = 0xffffffffu;

template <int KW>
struct LineLayout {
    static constexpr int CAP = (KW == 3) ? 5 : (KW == 2) ? 7 : (KW == 4) ? 3 : 15 / KW;
    static constexpr int kBaseWord = 0;
    // word index of word p (0 = most significant) of slot t.  For KW = 2, 3, 4 a key never straddles the two 32-byte halves
    // of the line except key 4 of KW = 3 (word 0 closes the first half, words 1-2 the second): two lanes, each holding one
    // half, compare without moving key words between them.
    __host__ __device__ static constexpr int word(int t, int p) {
        return KW == 3 ? (t < 2 ? 1 + 3 * t + p : t < 4 ? 8 + 3 * (t - 2) + p : (p == 0 ? 7 : 13 + p))
             : KW == 2 ? (t < 3 ? 2 + 2 * t + p : 8 + 2 * (t - 3) + p)     // first half: base, -, keys 0-2; second half: keys 3-6
             : KW == 4 ? (t < 1 ? 4 + p : 8 + 4 * (t - 1) + p)             // first half: base, -, -, -, key 0; second half: keys 1-2
                       : 1 + t * KW + p;
    }
};

struct IndexView {
    const uint64_t *keys;       // key column [n * s] (binary search, bucket overflow)
    const uint4 *lines;         // [nlines] 64-byte lines
    const uint2 *bins;          // [nbins] {first line, lines} (global copy; kernels may stage it in shared memory)
    uint64_t n;
    uint64_t first_index;
    uint64_t base;              // top 64 bits (left-aligned) of the first key of THIS array
    uint64_t span;              // top64(last key) - base
    uint64_t scale;             // bin.frac = ((top64(q) - base) << norm) * scale as a 64.64 fixed-point number
    uint32_t nbins;
    uint32_t norm;
    uint32_t k;
    uint32_t hints;             // bit0: line loads evict-first in L2 (touched at random, never again)
    uint32_t pad[8];            // wire form of the largest key
};

// hints: 0 = plain loads, 1 = evict-first in L2 (+ no L1 allocation), 2 = evict-normal policy object, 3 = evict-last
__device__ __forceinline__ uint64_t make_line_policy(uint32_t hints) {
    uint64_t p = 0;
    if (hints == 1u) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (hints == 2u) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    else if (hints == 3u) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Line loads carry the L2::64B prefetch-size qualifier: without it every random access pulls the whole 128-byte L2 line out of
// DRAM (147 bytes of DRAM reads per lookup measured); with it the 64-byte line is all that moves (89 bytes per lookup).  The
// access RATE does not change -- HBM serves about 3.9e10 random lines per second at 32, 64 or 128 bytes alike.
struct Half { uint32_t w[8]; };       // one 32-byte half of a line
__device__ __forceinline__ Half ld_line32(const uint4 *p, uint64_t policy) {
    Half h;
    if (policy == 0) {
        asm volatile("ld.global.nc.L2::64B.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(h.w[0]), "=r"(h.w[1]), "=r"(h.w[2]), "=r"(h.w[3]), "=r"(h.w[4]), "=r"(h.w[5]), "=r"(h.w[6]), "=r"(h.w[7]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                     : "=r"(h.w[0]), "=r"(h.w[1]), "=r"(h.w[2]), "=r"(h.w[3]), "=r"(h.w[4]), "=r"(h.w[5]), "=r"(h.w[6]), "=r"(h.w[7])
                     : "l"(p), "l"(policy));
    }
    return h;
}
__device__ __forceinline__ uint4 ld_line16(const uint4 *p, uint64_t policy) {
    uint4 v;
    if (policy == 0) {              // no hint at all: a table small enough for L2 to help is left to the default replacement
        asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    }
    return v;
}

// The 64 most significant bits of the 2k-bit key, left-aligned.
template <int S>
__host__ __device__ __forceinline__ uint64_t key_top64(const uint64_t *q, uint32_t k) {
    const uint32_t top_bits = 2u * k - 64u * (S - 1);          // bits used in word 0: 2..64
    if (S == 1 || top_bits == 64) return q[0] << (64u - top_bits);
    return (q[0] << (64u - top_bits)) | (q[S > 1 ? 1 : 0] >> top_bits);
}

// Bin of a key and its position inside the bin as a 64-bit fraction.  False: the key lies outside [first key, last key]
// of this array (as far as its top 64 bits tell) and cannot match.
template <int S>
__device__ __forceinline__ bool key_bin(const IndexView &ix, const uint64_t (&q)[S], uint32_t &bin, uint64_t &frac) {
    const uint64_t t = key_top64<S>(q, ix.k);
    if (t < ix.base) return false;
    const uint64_t d = t - ix.base;
    if (d > ix.span) return false;
    const uint64_t x = d << ix.norm;                           // span << norm does not overflow, so neither does this
    bin = (uint32_t)__umul64hi(x, ix.scale);
    frac = x * ix.scale;
    return true;
}
// `bins` may point to shared or global memory.
template <int S>
__device__ __forceinline__ bool key_line(const IndexView &ix, const uint2 *bins, const uint64_t (&q)[S], uint32_t &line) {
    uint32_t bin;
    uint64_t frac;
    if (!key_bin<S>(ix, q, bin, frac)) return false;
    const uint2 b = bins[bin];
    line = b.x + (uint32_t)__umul64hi(frac, (uint64_t)b.y);
    return true;
}

constexpr int kMaxShards = 64;
constexpr uint32_t kLinear = 8;     // range remainder scanned with independent loads

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

// Continuation of a bucket in the key column: every key before `lo` is < q; the first key >= q decides.
template <int S>
__device__ __noinline__ int64_t search_overflow(const uint64_t *__restrict__ keys, uint64_t lo, uint64_t n, const uint64_t (&q)[S]) {
#pragma unroll 1
    for (int step = 0; step < 8 && lo < n; ++step, ++lo) {
        uint64_t kk[S];
        load_key<S>(keys, lo, kk);
        if (!words_less<S>(kk, q)) return words_equal<S>(kk, q) ? (int64_t)lo : -1;
    }
    if (lo >= n) return -1;
    return search_range<S>(keys, lo, n, q);      // a pathologically long bucket: bounded by log2(n) probes
}

template <int S>
__device__ __forceinline__ int64_t lookup_bsearch(const IndexView &ix, const uint64_t (&q)[S]) {
    const int64_t r = search_range<S>(ix.keys, 0, ix.n, q);
    return r < 0 ? r : r + (int64_t)ix.first_index;
}

template <int S, int KW>
__device__ __forceinline__ void key_to_wire(const uint64_t (&q)[S], uint32_t *dst) {
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const int idx = KW - 1 - j;                       // 32-bit word index counted from the least significant end
        const uint64_t w = q[S - 1 - idx / 2];
        dst[j] = (idx & 1) ? (uint32_t)(w >> 32) : (uint32_t)w;
    }
}

// One query per thread, the whole line in one thread (any KW): the plain form of the line search, used where the
// threads of a warp do not walk in step (sorted pass) and for key widths without a cooperative layout.
template <int S, int KW>
__device__ __forceinline__ int64_t lookup_lines_thread(const IndexView &ix, const uint2 *bins, uint64_t pol, const uint64_t (&q)[S]) {
    using LL = LineLayout<KW>;
    uint32_t line;
    if (ix.n == 0 || !key_line<S>(ix, bins, q, line)) return -1;
    uint32_t w[kLineWords];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 v = ld_line16(ix.lines + (size_t)line * 4 + c, pol);
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
    uint32_t qw[KW];
    key_to_wire<S, KW>(q, qw);
    int slot = -1;
#pragma unroll
    for (int t = LL::CAP - 1; t >= 0; --t) {
        bool eq = true;
#pragma unroll
        for (int p = 0; p < KW; ++p) eq &= (w[LL::word(t, p)] == qw[p]);
        if (eq) slot = t;
    }
    const uint64_t base = w[LL::kBaseWord];
    if (slot >= 0) return (int64_t)(base + (uint64_t)slot + ix.first_index);
    bool less = false, decided = false;                    // last slot < q ?
#pragma unroll
    for (int p = 0; p < KW; ++p) {
        const uint32_t a = w[LL::word(LL::CAP - 1, p)];
        if (!decided && a != qw[p]) { decided = true; less = a < qw[p]; }
    }
    if (!less) return -1;
    const int64_t r = search_overflow<S>(ix.keys, base + LL::CAP, ix.n, q);
    return r < 0 ? r : r + (int64_t)ix.first_index;
}

// The line search by a whole warp, one query per lane (all 32 lanes must call it; `live` = this lane has a query).
// Two lanes share a line: lane h of a pair loads bytes [32h, 32h + 32) with one 256-bit load -- a warp-wide load fetches 16
// whole lines (16 wavefronts instead of the 64 a thread-per-line search costs) -- and the pair works through its own two
// queries in two rounds; both loads are issued before the first compare.  What crosses the lanes per round is the query
// (KW words), the line index, one word of match bits and the base index.  Layouts exist for KW = 2, 3, 4 (k = 17..64);
// other widths take the per-thread form.
template <int S, int KW>
__device__ __forceinline__ int64_t lookup_lines_warp(const IndexView &ix, const uint2 *bins, uint64_t pol, const uint64_t (&q)[S], bool live) {
    if constexpr (KW < 2 || KW > 4) {
        return live ? lookup_lines_thread<S, KW>(ix, bins, pol, q) : -1;
    } else {
    using LL = LineLayout<KW>;
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t h = threadIdx.x & 1u;
    uint32_t line = 0;                       // a lane without a query reads line 0 and ignores it: no branch around the loads
    bool in = live && ix.n != 0;
    if (in) in = key_line<S>(ix, bins, q, line);
    if (!in) line = 0;
    uint32_t qw[KW];
    key_to_wire<S, KW>(q, qw);
    Half v[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t l = __shfl_sync(FULL, line, r, 2);
        v[r] = ld_line32(ix.lines + (size_t)l * 4 + 2 * h, pol);
    }
    int slot_mine = -1;
    bool less_mine = false;
    uint32_t base_mine = 0;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        uint32_t qs[KW];
#pragma unroll
        for (int p = 0; p < KW; ++p) qs[p] = __shfl_sync(FULL, qw[p], r, 2);
        const uint32_t (&a)[8] = v[r].w;
        // bits 0..6: slot t matches; bit 8: last slot < query (decided in this half); bits 9, 10: KW = 3 only, see below
        uint32_t m = 0;
        if constexpr (KW == 3) {
            // half 0: base, key 0, key 1, word 0 of key 4;  half 1: key 2, key 3, words 1-2 of key 4
            uint32_t b[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) b[i] = h ? a[i] : a[i + 1];
            const bool ma = (b[0] == qs[0]) & (b[1] == qs[1]) & (b[2] == qs[2]);
            const bool mb = (b[3] == qs[0]) & (b[4] == qs[1]) & (b[5] == qs[2]);
            m = ((uint32_t)ma | ((uint32_t)mb << 1)) << (2u * h);
            if (h == 0) m |= ((uint32_t)(a[7] == qs[0]) << 9) | ((uint32_t)(a[7] < qs[0]) << 8);               // key 4, word 0: equal / below
            else m |= ((uint32_t)((a[6] == qs[1]) & (a[7] == qs[2])) << 10) |
                      ((uint32_t)((a[6] < qs[1]) | ((a[6] == qs[1]) & (a[7] < qs[2]))) << 11);                  // key 4, words 1-2: equal / below
        } else if constexpr (KW == 2) {
            // half 0: base, -, keys 0-2;  half 1: keys 3-6
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool e = (a[2 * t] == qs[0]) & (a[2 * t + 1] == qs[1]);
                if (t > 0) m |= (uint32_t)(e & (h == 0)) << (t - 1);
                m |= (uint32_t)(e & (h == 1)) << (3 + t);
            }
            if (h == 1) m |= (uint32_t)((a[6] < qs[0]) | ((a[6] == qs[0]) & (a[7] < qs[1]))) << 8;
        } else {
            // half 0: base, -, -, -, key 0;  half 1: keys 1-2
            const bool e1 = (a[4] == qs[0]) & (a[5] == qs[1]) & (a[6] == qs[2]) & (a[7] == qs[3]);
            const bool e0 = (a[0] == qs[0]) & (a[1] == qs[1]) & (a[2] == qs[2]) & (a[3] == qs[3]);
            if (h == 0) m = (uint32_t)e1;
            else {
                bool lt = false, dec = false;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (!dec && a[4 + p] != qs[p]) { dec = true; lt = a[4 + p] < qs[p]; }
                }
                m = ((uint32_t)e0 << 1) | ((uint32_t)e1 << 2) | ((uint32_t)lt << 8);
            }
        }
        m |= __shfl_xor_sync(FULL, m, 1);
        const uint32_t base = __shfl_sync(FULL, a[0], 0, 2);
        int slot = -1;
        bool less;
        if constexpr (KW == 3) {
            if (m & 0xfu) slot = __ffs(m & 0xfu) - 1;
            else if ((m & 0x600u) == 0x600u) slot = 4;
            less = ((m >> 8) & 1u) | (((m >> 9) & 1u) & ((m >> 11) & 1u));      // word 0 below, or equal and the rest below
        } else {
            if (m & 0x7fu) slot = __ffs(m & 0x7fu) - 1;
            less = (m >> 8) & 1u;
        }
        if (h == (uint32_t)r) { slot_mine = slot; less_mine = less; base_mine = base; }
    }
    if (!in) return -1;
    if (slot_mine >= 0) return (int64_t)((uint64_t)base_mine + (uint64_t)slot_mine + ix.first_index);
    if (!less_mine) return -1;
    const int64_t res = search_overflow<S>(ix.keys, (uint64_t)base_mine + LL::CAP, ix.n, q);
    return res < 0 ? res : res + (int64_t)ix.first_index;
    }
}

// Stages the bin table in shared memory when the kernel was given room for it (smem_bins != nullptr), else uses the global
// copy.  Must be called by every thread of the CTA (barrier inside).
__device__ __forceinline__ const uint2 *stage_bins(const IndexView &ix, uint2 *smem_bins) {
    if (!smem_bins) return ix.bins;
    for (uint32_t i = threadIdx.x; i < ix.nbins; i += blockDim.x) smem_bins[i] = ix.bins[i];
    __syncthreads();
    return smem_bins;
}

// ------------------------------------------------------------------ kernels: pack, find (fused with pack), find packed
struct SeqJob {
    const uint8_t *seq;
    uint64_t nq;          // windows / rows
    uint64_t stride;      // 1 = sliding windows, k = independent rows
    uint32_t k;
    uint32_t per_tile;    // windows per tile
};

// MODE: 0 = pack only, 1 = pack + line search, 2 = pack + binary search over the key column.
enum { SEQ_PACK = 0, SEQ_FIND = 1, SEQ_BSEARCH = 2 };

template <int S, int KW, int MODE>
__global__ void __launch_bounds__(kBlock) seq_kernel(SeqJob job, uint64_t *__restrict__ out_words, uint8_t *__restrict__ out_flags,
                                                     IndexView ix, int64_t *__restrict__ out_index) {
    __shared__ SeqTile tile;
    const uint64_t ntiles = (job.nq + job.per_tile - 1) / job.per_tile;
    const uint64_t pol = make_line_policy(ix.hints);
    for (uint64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
        const uint64_t w0 = tix * job.per_tile;
        const uint32_t nw = (uint32_t)min((uint64_t)job.per_tile, job.nq - w0);
        const uint32_t nb = (uint32_t)((nw - 1) * job.stride + job.k);
        __syncthreads();                                    // previous tile fully consumed
        const uint32_t ascii_off = stage_tile(tile, job.seq, w0 * job.stride, nb);
        // whole warps stay in the loop (the line search is warp-collective); lanes past the end carry no query
        for (uint32_t j0 = 0; j0 < nw; j0 += kBlock) {
            const uint32_t j = j0 + threadIdx.x;
            const bool have = j < nw;
            uint64_t q[S];
            uint32_t flags = 2u;
            if (have) flags = window_canonical<S>(tile, ascii_off, (uint32_t)(j * job.stride), job.k, q);
            else {
#pragma unroll
                for (int w = 0; w < S; ++w) q[w] = 0;
            }
            if (MODE == SEQ_FIND) {
                const int64_t r = lookup_lines_warp<S, KW>(ix, ix.bins, pol, q, have && (flags & 6u) == 0);
                if (have) out_index[w0 + j] = r;
            } else if (MODE == SEQ_BSEARCH) {
                if (have) out_index[w0 + j] = (flags & 6u) == 0 ? lookup_bsearch<S>(ix, q) : -1;
            } else if (have) {
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

template <int S>
__global__ void __launch_bounds__(kBlock) find_bsearch_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags,
                                                              uint64_t nq, IndexView ix, int64_t *__restrict__ out_index) {
    for (uint64_t i = (uint64_t)blockIdx.x * kBlock + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * kBlock) {
        uint64_t q[S];
        load_key<S>(words, i, q);
        const bool skip = flags && (flags[i] & 6u);
        out_index[i] = skip ? -1 : lookup_bsearch<S>(ix, q);
    }
}

// K4 on packed queries: a persistent grid, the bin table in shared memory, one query per lane and four line loads in
// flight per lane (lookup_lines_warp); the next iteration's query is fetched before this one's lines are awaited.
// Per lookup the kernel moves the query (8s + 1 bytes), one 64-byte line and the 8-byte result.
constexpr int kFindBlock = 512;

template <int S, int KW>
__global__ void __launch_bounds__(kFindBlock, 2) find_packed_lines_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags,
                                                                          uint64_t nq, IndexView ix, int64_t *__restrict__ out_index, int bins_in_smem) {
    extern __shared__ __align__(16) uint2 find_bins_smem[];
    const uint2 *bins = stage_bins(ix, bins_in_smem ? find_bins_smem : nullptr);
    const uint64_t pol = make_line_policy(ix.hints);
    const uint64_t span = (uint64_t)gridDim.x * kFindBlock;
    uint64_t i = (uint64_t)blockIdx.x * kFindBlock + threadIdx.x;
    uint64_t q[S];
    bool live = false;
    auto fetch = [&](uint64_t at) {
        live = at < nq;
#pragma unroll
        for (int w = 0; w < S; ++w) q[w] = 0;
        if (live) {
            load_key<S>(words, at, q);
            if (flags && (__ldg(flags + at) & 6u)) live = false;
        }
    };
    fetch(i);
    // warp-uniform trip count: the first lane of the warp decides
    for (uint64_t w0 = i - (threadIdx.x & 31u); w0 < nq; w0 += span, i += span) {
        uint64_t qc[S];
#pragma unroll
        for (int w = 0; w < S; ++w) qc[w] = q[w];
        const bool lc = live;
        fetch(i + span);
        const int64_t r = lookup_lines_warp<S, KW>(ix, bins, pol, qc, lc);
        if (i < nq) out_index[i] = lc ? r : -1;
    }
}

// ------------------------------------------------------------------ K3 (+K4) for independent k-byte rows (query lists)
// A CTA takes kRowsPerTile consecutive rows = one contiguous byte range.  Phase 1: every thread pulls 16-byte chunks of
// the 16-byte-aligned superset straight into registers (coalesced LDG.128) and converts each to one 32-bit word of 2-bit
// codes with byte-parallel arithmetic (4 bases per 32-bit op: code = ((u>>1)^(u>>2))&3 on u = byte & ~0x20, packed by one
// multiply; validity by rebuilding the letter from its code with PRMT and comparing).  Only the code stream goes to shared
// memory (0.25 byte per base).  A chunk holding anything but upper-case ACGT marks the rows it overlaps as suspect.
// Phase 2: thread r cuts row r out of the stream with funnel shifts, reverse-complements in registers, keeps the smaller
// and either writes words + flags (FIND = false) or searches the index with two rows in flight (FIND = true).
// Suspect rows (N, lower case, other bytes: rare) take the reference's byte loop on the global bytes instead.
constexpr uint32_t kRowPadWords = 2;           // 32 zero bases in front of the stream (word 0 of row 0 starts before it)

__device__ __forceinline__ uint32_t codes_of_4(uint32_t w, uint32_t &diff) {
    const uint32_t u = w & 0xdfdfdfdfu;                                   // upper-cased
    const uint32_t x = ((u >> 1) ^ (u >> 2)) & 0x03030303u;              // 2-bit code in every byte (A0 C1 G2 T3)
    const uint32_t y = x | (x >> 4);
    const uint32_t z = y & 0x00ff00ffu;
    const uint32_t sel = (z | (z >> 8)) & 0xffffu;                        // the four codes as PRMT selectors
    diff |= u ^ __byte_perm(0x54474341u, 0u, sel);                        // 'A','C','G','T' rebuilt from the codes
    return x * 0x40100401u;                                               // byte 3 = c0<<6 | c1<<4 | c2<<2 | c3
}

// Spread the 32 bits of x to the even bit positions of a 64-bit word.
__device__ __forceinline__ uint64_t spread_bits(uint32_t x) {
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// Exact path for one row, evaluated by a whole warp (same flags and canonical rule as window_canonical): lane i looks
// at bytes i, i+32, ...; the 2-bit planes are collected with ballots and interleaved into packed words; the reference's
// "first differing byte decides" loop (SequenceUtils.java:211-219) becomes two ballots per 32 bytes.  Every lane
// returns the same flags and words.  `a` may point to shared or global memory.
template <int S>
__device__ __forceinline__ uint32_t row_canonical_warp(const uint8_t *a, uint32_t k, uint32_t lane, uint64_t (&out)[S]) {
    uint64_t t[S];                                   // the sequence left-aligned in 64*S bits
    uint32_t bad = 0, low = 0;
#pragma unroll
    for (int c = 0; c < S; ++c) {
        const uint32_t i = 32u * c + lane;
        const bool in = i < k;
        const uint32_t ch = in ? a[i] : (uint32_t)'A';
        const uint32_t f = ch | 0x20u;
        const bool ok = (f == 'a') | (f == 'c') | (f == 'g') | (f == 't');
        bad |= __ballot_sync(0xffffffffu, in && !ok);
        low |= __ballot_sync(0xffffffffu, in && ok && (ch & 0x20u));
        const uint32_t code = ok ? base_code(ch) : 0u;
        const uint32_t p0 = __brev(__ballot_sync(0xffffffffu, code & 1u)), p1 = __brev(__ballot_sync(0xffffffffu, code & 2u));
        t[c] = (spread_bits(p1) << 1) | spread_bits(p0);     // lane 0 = first base = the two highest bits
    }
    if (bad) {
#pragma unroll
        for (int w = 0; w < S; ++w) out[w] = 0;
        return 2u;
    }
    uint64_t fw[S], rc[S];
    const uint32_t sh = 64u * S - 2u * k;                     // right-align (0..62, even)
#pragma unroll
    for (int i = 0; i < S; ++i) {
        uint64_t v = sh ? (t[i] >> sh) : t[i];
        if (i > 0 && sh) v |= t[i - 1] << (64u - sh);
        fw[i] = v;
    }
    revcomp_words<S>(fw, rc, k);
    bool flip = less_by_borrow<S>(rc, fw);
    uint32_t flags = 0;
    if (low) {                                                // mixed / lower case: ASCII bytes decide, not 2-bit codes
        flags |= 4u;
        flip = false;
#pragma unroll
        for (int c = 0; c < S; ++c) {
            const uint32_t i = 32u * c + lane;
            const bool in = i < k;
            const int8_t f = in ? (int8_t)a[i] : (int8_t)0;
            const int8_t r = in ? (int8_t)complement_ascii(a[k - 1u - i]) : (int8_t)0;
            const uint32_t lt = __ballot_sync(0xffffffffu, in && f < r), gt = __ballot_sync(0xffffffffu, in && f > r);
            if (lt | gt) {
                flip = (gt >> (__ffs(lt | gt) - 1)) & 1u;
                break;
            }
        }
    }
#pragma unroll
    for (int w = 0; w < S; ++w) out[w] = flip ? rc[w] : fw[w];
    return flags | (flip ? 1u : 0u);
}

// Shared memory of rows_kernel: two raw tiles (filled by 1-D bulk copies, the TMA engine) and two code streams.
__host__ __device__ inline uint32_t rows_tile_bytes(uint32_t rows_per_tile, uint32_t k) { return (rows_per_tile * k + 15u + 15u) & ~15u; }
__host__ __device__ inline uint32_t rows_stream_words(uint32_t rows_per_tile, uint32_t k) {
    return ((rows_tile_bytes(rows_per_tile, k) / 16u + kRowPadWords + 3u) + 3u) & ~3u;
}
inline size_t rows_smem_bytes(uint32_t rows_per_tile, uint32_t k, bool find) {
    return (find ? 2u : 3u) * (size_t)rows_tile_bytes(rows_per_tile, k) + 2u * (size_t)rows_stream_words(rows_per_tile, k) * 4u;
}

template <int S, int KW, bool FIND, int RPT>
__global__ void __launch_bounds__(kBlock, (FIND && S <= 2) ? 3 : 1) rows_kernel(const uint8_t *__restrict__ kmers, uint64_t nq, uint32_t k,
                                                                                uint64_t *__restrict__ out_words, uint8_t *__restrict__ out_flags,
                                                                                IndexView ix, int64_t *__restrict__ out_index) {
    constexpr uint32_t kRows = RPT * kBlock;                      // rows per tile
    // Raw tiles in flight.  The pack variant keeps a tile's bytes until its phase 2 is over (the exact path for suspect rows
    // reads them from shared memory), which takes a third buffer; the find variant reads those rare rows from global
    // memory instead and keeps the shared memory for occupancy (the search needs the warps).
    constexpr uint32_t NRAW = FIND ? 2u : 3u;
    extern __shared__ __align__(128) uint8_t rows_smem[];
    const uint32_t tile_cap = rows_tile_bytes(kRows, k), stream_cap = rows_stream_words(kRows, k);
    uint32_t *const codes0 = reinterpret_cast<uint32_t *>(rows_smem + NRAW * tile_cap);   // raw tiles at 0, tile_cap, ...
    __shared__ uint32_t suspect[3][kRows / 32];    // per tile in flight; see the clearing rule in phase 1
    __shared__ __align__(8) uint64_t full[NRAW];
    const uint64_t ntiles = (nq + kRows - 1) / kRows;
    const uint32_t top_bits = 2u * k - 64u * (S - 1);
    const uint64_t policy = make_evict_first_policy();
    const uint64_t pol = make_line_policy(FIND ? ix.hints : 0u);

    // tile t of this CTA: rows [row0, row0 + rows); the copy fetches the 16-byte-aligned superset of its bytes
    auto issue = [&](uint64_t tile, uint32_t buf) {
        const uint64_t row0 = tile * kRows;
        const uint32_t rows = (uint32_t)min((uint64_t)kRows, nq - row0);
        const uint8_t *src = kmers + row0 * k;
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint32_t bytes = (off + rows * k + 15u) & ~15u;
        mbar_arrive_expect_tx(&full[buf], bytes);
        bulk_g2s(rows_smem + buf * tile_cap, src - off, bytes, &full[buf], policy);
    };

    if (threadIdx.x == 0) {
        for (uint32_t b = 0; b < NRAW; ++b) mbar_init(&full[b], 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 3 * (kRows / 32)) (&suspect[0][0])[threadIdx.x] = 0;
    if (threadIdx.x < kRowPadWords) codes0[threadIdx.x] = codes0[stream_cap + threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x < ntiles) issue(blockIdx.x, 0);
        if (blockIdx.x + (uint64_t)gridDim.x < ntiles) issue(blockIdx.x + (uint64_t)gridDim.x, 1);
    }
    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1u, rb = it % NRAW;
        const uint64_t row0 = tile * kRows;
        const uint32_t rows = (uint32_t)min((uint64_t)kRows, nq - row0);
        const uint8_t *src = kmers + row0 * k;
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint32_t nbytes = rows * k;
        const uint32_t nchunks = (off + nbytes + 15u) >> 4;
        const uint32_t sus = it % 3u;
        // With one barrier per tile, suspect[(it+1)%3] was last read in phase 2 of tile it-2 (every thread is past the
        // barrier of tile it-1, so done with it) and is next written in phase 1 of tile it+1 (after this tile's barrier).
        if (threadIdx.x < kRows / 32) suspect[(it + 1u) % 3u][threadIdx.x] = 0;
        mbar_wait(&full[rb], (it / NRAW) & 1u, nullptr, DEV_TIMEOUT_FULL);
        // ---- phase 1: 16-byte chunks of the raw tile -> code words
        const uint4 *raw = reinterpret_cast<const uint4 *>(rows_smem + rb * tile_cap);
        uint32_t *stream = codes0 + par * stream_cap;
#pragma unroll 2
        for (uint32_t j = threadIdx.x; j < nchunks; j += kBlock) {
            const uint4 v = raw[j];
            uint32_t diff = 0;
            const uint32_t m0 = codes_of_4(v.x, diff), m1 = codes_of_4(v.y, diff);
            const uint32_t m2 = codes_of_4(v.z, diff), m3 = codes_of_4(v.w, diff);
            const uint32_t lo = __byte_perm(m3, m2, 0x7373u), hi = __byte_perm(m1, m0, 0x7373u);
            stream[kRowPadWords + j] = __byte_perm(lo, hi, 0x5410u);
            diff |= (v.x | v.y | v.z | v.w) & 0x20202020u;
            if (diff) {                                          // rare: mark every row this chunk overlaps
                const int64_t b_lo = max((int64_t)16 * j - off, (int64_t)0);
                const int64_t b_hi = min((int64_t)16 * j + 16 - off, (int64_t)nbytes);
                if (b_lo < b_hi) {
                    const uint32_t r0 = (uint32_t)b_lo / k, r1 = (uint32_t)(b_hi - 1) / k;
                    for (uint32_t r = r0; r <= r1; ++r) atomicOr(&suspect[sus][r >> 5], 1u << (r & 31u));
                }
            }
        }
        if (threadIdx.x < 3) stream[kRowPadWords + nchunks + threadIdx.x] = 0;
        __syncthreads();                                         // stream complete
        // the tile after next goes into the raw buffer whose last readers are behind this barrier: this tile's own buffer
        // (NRAW = 2: its phase 2 does not touch raw bytes) or the previous tile's (NRAW = 3)
        if (threadIdx.x == 0 && tile + 2ull * gridDim.x < ntiles) issue(tile + 2ull * gridDim.x, (it + 2u) % NRAW);
        // ---- phase 2: one thread per row
        uint64_t q[RPT][S];
        uint32_t flags[RPT];
        bool have[RPT];
#pragma unroll
        for (int h = 0; h < RPT; ++h) {
            const uint32_t r = threadIdx.x + h * kBlock;
            have[h] = r < rows;
            flags[h] = 2u;
#pragma unroll
            for (int w = 0; w < S; ++w) q[h][w] = 0;
            const bool sus_row = have[h] && ((suspect[sus][r >> 5] >> (r & 31u)) & 1u);
            // suspect rows (rare) are evaluated exactly, one at a time, by the whole warp
            uint32_t todo = __ballot_sync(0xffffffffu, sus_row);
            while (todo) {
                const uint32_t sl = __ffs(todo) - 1u;
                todo &= todo - 1u;
                const uint32_t rs = (threadIdx.x & ~31u) + sl + h * kBlock;
                const uint8_t *a = FIND ? src + (size_t)rs * k : rows_smem + rb * tile_cap + off + (size_t)rs * k;
                uint64_t exact[S];
                const uint32_t fl = row_canonical_warp<S>(a, k, threadIdx.x & 31u, exact);
                if ((threadIdx.x & 31u) == sl) {
                    flags[h] = fl;
#pragma unroll
                    for (int w = 0; w < S; ++w) q[h][w] = exact[w];
                }
            }
            if (have[h] && !sus_row) {
                const uint32_t p0 = 16u * kRowPadWords + off + r * k;           // stream position (in bases) of the row's first base
                uint64_t fw[S], rc[S];
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const int32_t b0 = (int32_t)k - 32 * (S - j);                 // word j holds bases [b0, b0 + 32)
                    fw[j] = extract64(stream, 2u * (uint32_t)((int32_t)p0 + b0));
                }
                if (top_bits < 64) fw[0] &= (1ull << top_bits) - 1ull;
                revcomp_words<S>(fw, rc, k);
                const bool flip = less_by_borrow<S>(rc, fw);
#pragma unroll
                for (int w = 0; w < S; ++w) q[h][w] = flip ? rc[w] : fw[w];
                flags[h] = flip ? 1u : 0u;
            }
        }
        if (FIND) {
            // the line search is warp-collective: every thread of the CTA gets here for every h
#pragma unroll
            for (int h = 0; h < RPT; ++h) {
                const int64_t res = lookup_lines_warp<S, KW>(ix, ix.bins, pol, q[h], have[h] && (flags[h] & 6u) == 0);
                if (have[h]) out_index[row0 + threadIdx.x + h * kBlock] = res;
            }
        } else {
#pragma unroll
            for (int h = 0; h < RPT; ++h) {
                if (!have[h]) continue;
                const uint64_t row = row0 + threadIdx.x + h * kBlock;
                if (S == 2) reinterpret_cast<ulonglong2 *>(out_words)[row] = make_ulonglong2(q[h][0], q[h][1 % S]);
                else {
#pragma unroll
                    for (int w = 0; w < S; ++w) out_words[row * S + w] = q[h][w];
                }
                out_flags[row] = (uint8_t)flags[h];
            }
        }
    }
}

// rows_warp_kernel: the pack-only form of rows_kernel with WARPS as the unit instead of CTAs.  A warp owns 32 consecutive
// rows: lane 0 fetches their bytes with one bulk copy into the warp's private ring (NRAW buffers, one mbarrier each), the
// warp converts its ~32k/16 chunks (three rounds at k = 47) into its private code stream, and lane r cuts row r out of it.
// Nothing is shared between warps, so there is no CTA barrier in the loop: rows_kernel lost 36 % of its stall samples at
// the one __syncthreads per 256-row tile (profiles/r1_rows_kernel_hotspots.txt).  Suspect rows are a 32-bit mask in a
// register (REDUX over the lanes' chunk findings) instead of shared-memory atomics.  Validity costs 3 instructions per
// 4 bases here (codes_good_4) instead of the 8 of rebuilding the letters with PRMT; the per-tile bookkeeping runs on
// pointers advanced by a constant and on 32-bit shared addresses computed once.
inline size_t rows_warp_smem_bytes(uint32_t k, uint32_t nraw) {
    return (size_t)(kBlock / 32) * (nraw * (size_t)rows_tile_bytes(32, k) + (size_t)rows_stream_words(32, k) * 4u);
}

// Four bases -> byte 3 = c0 | c1<<2 | c2<<4 | c3<<6 (first base lowest: the warp kernel's stream is little-endian), and
// `good` keeps bit 4 of every byte set only while the byte's low five bits are those of an upper-case A, C, G or T:
// (b2 b1 b0) in {001, 011, 111, 100} and b4 == !b0 (T is the one letter with b0 = 0, and the one with b4 = 1).  Bits 7, 6,
// 5, 3 are checked once per 16 bytes by the caller.  Written with LEFT shifts only, and those as multiplications by
// factors the compiler cannot see (kernel arguments): the integer ALU pipe (LOP3 / SHF / PRMT, one warp instruction per
// two cycles) is what limits this kernel, the FMA pipe (IMAD) is idle -- 5 IMAD + 3 LOP3 per four bases.
struct RowsMul { uint32_t m2, m4, m8, m16, pack; };
constexpr RowsMul kRowsMul{2u, 4u, 8u, 16u, 0x00410410u};
__device__ __forceinline__ uint32_t codes_good_4(uint32_t w, uint32_t &good, const RowsMul &mu) {
    const uint32_t l1 = w * mu.m2, l2 = w * mu.m4, l3 = w * mu.m8, l4 = w * mu.m16;
    const uint32_t h = (l4 & (l3 | ~l2)) | (~l4 & ~l3 & l2);            // bit 4: b0 (b1 | !b2) | !b0 !b1 b2
    good &= h & (w ^ l4);
    return ((l1 ^ w) & 0x0c0c0c0cu) * mu.pack;                           // codes (b2^b3, b1^b2) at bits 3:2 of every byte
}

__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

template <int S, int KW, bool FIND, int NRAW>
__global__ void __launch_bounds__(kBlock, FIND ? 4 : 0) rows_warp_kernel(const uint8_t *__restrict__ kmers, uint64_t nq, uint32_t k,
                                                           uint64_t *__restrict__ out_words, uint8_t *__restrict__ out_flags, const RowsMul mu,
                                                           IndexView ix, int64_t *__restrict__ out_index) {
    constexpr uint32_t kWarps = kBlock / 32;
    extern __shared__ __align__(128) uint8_t rows_smem[];
    __shared__ __align__(8) uint64_t full[kWarps][NRAW];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t raw_cap = rows_tile_bytes(32, k), stream_cap = rows_stream_words(32, k);
    uint8_t *const raw0 = rows_smem + (size_t)warp * (NRAW * raw_cap + stream_cap * 4u);
    uint32_t *const stream = reinterpret_cast<uint32_t *>(raw0 + NRAW * raw_cap);
    const uint32_t raw_s = smem_u32(raw0), bar_s = smem_u32(&full[warp][0]);
    const uint64_t ntiles = (nq + 31u) / 32u, last = ntiles - 1;
    const uint64_t gw = (uint64_t)blockIdx.x * kWarps + warp, nw = (uint64_t)gridDim.x * kWarps;
    const uint64_t step = nw * 32u * k;                             // bytes between consecutive tiles of this warp
    const uint32_t tail_rows = (uint32_t)(nq - last * 32u);
    const uint32_t top_bits = 2u * k - 64u * (S - 1);
    const uint64_t policy = make_evict_first_policy();
    const uint64_t pol = make_line_policy(FIND ? ix.hints : 0u);

    auto issue = [&](const uint8_t *src, uint32_t rows, uint32_t buf) {
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint32_t bytes = (off + rows * k + 15u) & ~15u;
        const uint32_t bar = bar_s + 8u * buf;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(raw_s + buf * raw_cap), "l"(src - off), "r"(bytes), "r"(bar), "l"(policy) : "memory");
    };

    if (lane == 0) {
        for (uint32_t b = 0; b < NRAW; ++b) mbar_init(&full[warp][b], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint8_t *src = kmers + gw * 32u * k;
    if (lane == 0) {
#pragma unroll
        for (uint32_t b = 0; b < NRAW; ++b) {
            const uint64_t t = gw + b * nw;
            if (t < ntiles) issue(src + b * step, t == last ? tail_rows : 32u, b);
        }
    }
    uint32_t rb = 0, par = 0;
    for (uint64_t tile = gw; tile < ntiles; tile += nw, src += step) {
        const uint32_t rows = tile == last ? tail_rows : 32u;
        const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint32_t nbytes = rows * k;
        const uint32_t nchunks = (off + nbytes + 15u) >> 4;
        if (!mbar_try_wait_s(bar_s + 8u * rb, par)) mbar_wait(&full[warp][rb], par, nullptr, DEV_TIMEOUT_FULL);
        // ---- phase 1: the warp's 16-byte chunks -> code words
        const uint8_t *const rawb = raw0 + rb * raw_cap;
        const uint4 *raw = reinterpret_cast<const uint4 *>(rawb);
        uint32_t sus = 0;
        for (uint32_t j = lane; j < nchunks; j += 32u) {
            const uint4 v = raw[j];
            uint32_t good = 0xffffffffu;
            const uint32_t m0 = codes_good_4(v.x, good, mu), m1 = codes_good_4(v.y, good, mu);
            const uint32_t m2 = codes_good_4(v.z, good, mu), m3 = codes_good_4(v.w, good, mu);
            const uint32_t lo = __byte_perm(m0, m1, 0x7373u), hi = __byte_perm(m2, m3, 0x7373u);
            stream[j] = __byte_perm(lo, hi, 0x5410u);            // base i of the chunk at bits 2i+1:2i
            // bits 7, 5, 3 clear (5 = lower case) and bit 6 set in all 16 bytes; the low bits as collected in `good`
            const uint32_t diff = ((v.x | v.y | v.z | v.w) & 0xa8a8a8a8u) | (~(v.x & v.y & v.z & v.w) & 0x40404040u) | (~good & 0x10101010u);
            if (diff) {                                          // rare: every row this chunk overlaps is suspect
                const int32_t b_lo = max((int32_t)(16u * j) - (int32_t)off, 0);
                const int32_t b_hi = min((int32_t)(16u * j + 16u) - (int32_t)off, (int32_t)nbytes);
                if (b_lo < b_hi) {
                    const uint32_t r0 = (uint32_t)b_lo / k, r1 = (uint32_t)(b_hi - 1) / k;
                    sus |= ((2u << r1) - 1u) & ~((1u << r0) - 1u);
                }
            }
        }
        if (lane < 3) stream[nchunks + lane] = 0;
        sus = __reduce_or_sync(0xffffffffu, sus);
        __syncwarp();                                            // stream complete
        // ---- phase 2: lane r = row r
        uint64_t q[S];
        uint32_t flags = 2u;
#pragma unroll
        for (int w = 0; w < S; ++w) q[w] = 0;
        for (uint32_t todo = sus; todo;) {                      // suspect rows (rare): exact, one at a time, by the whole warp
            const uint32_t sl = __ffs(todo) - 1u;
            todo &= todo - 1u;
            uint64_t exact[S];
            const uint32_t fl = row_canonical_warp<S>(rawb + off + sl * k, k, lane, exact);
            if (lane == sl) {
                flags = fl;
#pragma unroll
                for (int w = 0; w < S; ++w) q[w] = exact[w];
            }
        }
        if (lane < rows) {
            if (!((sus >> lane) & 1u)) {
                // the row's 2k bits, first base lowest, cut out of the little-endian stream; E[j] = bits [64j, 64j + 64)
                const uint32_t P = 2u * (off + lane * k), idx = P >> 5, sh = P & 31u;
                uint64_t E[S];
                uint32_t prev = stream[idx];
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const uint32_t w1 = stream[idx + 2 * j + 1], w2 = stream[idx + 2 * j + 2];
                    E[j] = ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(prev, w1, sh);
                    prev = w2;
                }
                if (top_bits < 64) E[S - 1] &= (1ull << top_bits) - 1ull;
                // reverse complement = the complement read in this order; forward = the 2-bit groups reversed, right-aligned
                uint64_t fw[S], rc[S], t[S];
#pragma unroll
                for (int i = 0; i < S; ++i) { rc[i] = ~E[S - 1 - i]; t[i] = rev2(E[i]); }
                if (top_bits < 64) rc[0] &= (1ull << top_bits) - 1ull;
                const uint32_t rs = 64u * S - 2u * k;                           // 0..62, even
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    uint64_t v = t[i] >> rs;
                    if (i > 0 && rs) v |= t[i - 1] << (64u - rs);
                    fw[i] = v;
                }
                const bool flip = less_by_borrow<S>(rc, fw);
#pragma unroll
                for (int w = 0; w < S; ++w) q[w] = flip ? rc[w] : fw[w];
                flags = flip ? 1u : 0u;
            }
            if (!FIND) {
                const uint64_t row = tile * 32u + lane;
                if (S == 2) reinterpret_cast<ulonglong2 *>(out_words)[row] = make_ulonglong2(q[0], q[1 % S]);
                else {
#pragma unroll
                    for (int w = 0; w < S; ++w) out_words[row * S + w] = q[w];
                }
                out_flags[row] = (uint8_t)flags;
            }
        }
        if (FIND) {                                              // warp-collective: every lane gets here
            const int64_t res = lookup_lines_warp<S, KW>(ix, ix.bins, pol, q, lane < rows && (flags & 6u) == 0);
            if (lane < rows) out_index[tile * 32u + lane] = res;
        }
        __syncwarp();                                            // every lane is done with this tile's bytes and stream
        if (lane == 0) {
            const uint64_t t = tile + NRAW * nw;
            if (t < ntiles) issue(src + NRAW * step, t == last ? tail_rows : 32u, rb);
        }
        if (++rb == NRAW) { rb = 0; par ^= 1u; }
    }
}

// findRecord for a handful of k-mers in one launch (the legacy per-record API: a vertex and its neighbours): one warp per
// query reads the ASCII k-mer straight from mapped host memory, canonicalises it exactly (row_canonical_warp), searches
// the line index and copies the record's bytes back into mapped host memory -- no staging copies on either side.
template <int S, int KW>
__global__ void __launch_bounds__(256) find_small_kernel(const uint8_t *__restrict__ kmers, uint32_t nq, uint32_t k, IndexView ix,
                                                         const uint8_t *__restrict__ body, uint32_t rec_bytes,
                                                         int64_t *__restrict__ out_index, uint8_t *__restrict__ out_raw) {
    const uint32_t w = (blockIdx.x * 256u + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const bool have = w < nq;                                   // whole warps: the search below is warp-collective
    uint64_t q[S];
    uint32_t fl = 2u;
    if (have) fl = row_canonical_warp<S>(kmers + (size_t)w * k, k, lane, q);
    else {
#pragma unroll
        for (int i = 0; i < S; ++i) q[i] = 0;
    }
    const uint64_t pol = make_line_policy(0u);
    int64_t r = lookup_lines_warp<S, KW>(ix, ix.bins, pol, q, have && lane == 0 && (fl & 6u) == 0);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (!have) return;
    if (lane == 0) out_index[w] = r;
    if (out_raw) {
        uint8_t *dst = out_raw + (size_t)w * rec_bytes;
        const uint8_t *src = r >= 0 ? body + (uint64_t)(r - (int64_t)ix.first_index) * rec_bytes : nullptr;
        for (uint32_t b = lane; b < rec_bytes; b += 32u) dst[b] = src ? src[b] : (uint8_t)0;
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
//   route   owner of every query (splitter search) + block-aggregated reservation + P2P STORE of the key, in the
//           compact wire format, into the owner's inbox segment reserved for this rank.  What stays local is only the
//           tile bookkeeping: the position of every query inside its tile's regrouped order (2 bytes) and, per (tile,
//           owner), where that run went.
//   search  the owner walks all inbox segments, searches its shard and P2P-STORES each result (4-byte index local to
//           the shard) into the origin's return buffer at the same segment position
//   gather  the origin re-reads, tile by tile, the runs its tile sent (coalesced), rebases them by the owner's first
//           record index in shared memory and writes out[] in query order (coalesced) -- no scattered global access.
// Wire format of a key: the KW = ceil(2k/32) low 32-bit words of the 64*S-bit number, most significant first
// (12 bytes instead of 16 at k = 47).
constexpr int kRouteQ = 8;                   // queries per thread per tile of the route kernel
// Tile = block * kRouteQ queries.  Few owners: 128-thread blocks (1024-query tiles, more independent CTAs per SM for the
// barrier-heavy tile loop); many owners (virtual shards): 256-thread blocks, so that an owner's run in a tile stays a
// few hundred bytes.  Route and gather of one batch must agree on it: both derive it from nshards.
__host__ __device__ inline int route_block_for(int nshards) { return nshards > 16 ? 256 : 128; }
constexpr uint32_t kNotRouted = 0xffffu;
constexpr uint32_t kWireMiss = 0xffffffffu;

struct PeerPtrs {
    void *p[kMaxShards];
};

template <int S, int KW>
__device__ __forceinline__ void wire_to_key(const uint32_t *__restrict__ src, uint64_t (&q)[S]) {
#pragma unroll
    for (int w = 0; w < S; ++w) q[w] = 0;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const int idx = KW - 1 - j;
        const uint64_t v = __ldg(src + j);
        q[S - 1 - idx / 2] |= (idx & 1) ? (v << 32) : v;
    }
}

// Route state local to the origin (cc_route_state_bytes): at16[cap_q] | tile_base[ntiles][nshards] | tile_cnt[ntiles][nshards]
struct RouteState {
    uint16_t *at16;
    uint32_t *tile_base;
    uint32_t *tile_cnt;
};
__host__ __device__ inline uint64_t route_tiles(uint64_t nq, uint32_t tile_q) { return (nq + tile_q - 1) / tile_q; }
inline uint64_t route_state_bytes(uint64_t max_q, int nshards) {
    const uint64_t at = (max_q * 2 + 15) & ~15ull;
    return at + 2 * route_tiles(max_q, route_block_for(nshards) * kRouteQ) * (uint64_t)nshards * 4 + 16;
}
inline RouteState route_state_of(void *buf, uint64_t max_q, int nshards) {
    RouteState r;
    uint8_t *b = static_cast<uint8_t *>(buf);
    r.at16 = reinterpret_cast<uint16_t *>(b);
    b += (max_q * 2 + 15) & ~15ull;
    r.tile_base = reinterpret_cast<uint32_t *>(b);
    r.tile_cnt = r.tile_base + route_tiles(max_q, route_block_for(nshards) * kRouteQ) * (uint64_t)nshards;
    return r;
}
// Exclusive prefix sums of up to 64 values by one warp: lane l holds elements 2l and 2l+1; returns their prefixes, and the
// total in every lane.
__device__ __forceinline__ void warp_excl_scan64(uint32_t v0, uint32_t v1, uint32_t lane, uint32_t &p0, uint32_t &p1, uint32_t &total) {
    const uint32_t s2 = v0 + v1;
    uint32_t inc = s2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    p0 = inc - s2;
    p1 = p0 + v0;
    total = __shfl_sync(0xffffffffu, inc, 31);
}

constexpr uint32_t kOwnerLutBits = 10;         // owner lookup: 2^10 key prefixes -> first candidate owner

// Dynamic shared memory of route_kernel: two staging areas of wire keys (each owner's run padded to its destination's
// 16-byte phase).
inline uint32_t route_stage_words(uint32_t tile_q, uint32_t kw, int nshards) { return (tile_q * kw + 3u * kw * (uint32_t)nshards + 3u) & ~3u; }   // + up to 3 pad keys per owner
inline size_t route_smem_bytes(uint32_t tile_q, uint32_t kw, int nshards, int depth) {
    return (size_t)depth * route_stage_words(tile_q, kw, nshards) * 4u;     // `depth` staging areas: tiles' runs leave while the next are built
}

// Tile of BLOCK * kRouteQ queries per block iteration.  The tile is regrouped by owner in shared memory, every owner's run
// at the 16-byte phase of its destination, and leaves as aligned 16-byte stores (full NVLink write packets instead of
// per-query fragments).
template <int S, int KW, int BLOCK>
__global__ void __launch_bounds__(BLOCK) route_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags, uint64_t nq,
                                                      const uint64_t *__restrict__ splitters, int nshards, int my_rank, uint64_t cap,
                                                      PeerPtrs inbox, RouteState rs, unsigned long long *cursors, uint32_t stage_words,
                                                      uint32_t k_bases, uint32_t depth) {
    constexpr uint32_t kTile = BLOCK * kRouteQ;
    extern __shared__ __align__(128) uint32_t stage_all[];   // 2 x [stage_words] wire keys (double-buffered: see the copy-out)
    __shared__ uint32_t hist[kMaxShards], loc[kMaxShards + 1], locw[kMaxShards], endw[kMaxShards];
    __shared__ unsigned long long base[kMaxShards];
    __shared__ uint32_t *dptr[kMaxShards];                // dptr[o] + w = destination of staged word w of owner o
    __shared__ uint64_t spl[(kMaxShards - 1) * S];
    __shared__ uint8_t owner_lut[1u << kOwnerLutBits];    // number of splitters below the smallest key of each prefix
    for (int i = threadIdx.x; i < (nshards - 1) * S; i += BLOCK) spl[i] = splitters[i];
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < (1u << kOwnerLutBits); p += BLOCK) {
        // smallest key with these top bits, as S right-aligned words of a 2k-bit key
        uint64_t lowkey[S];
        const uint32_t kbits = 2u * k_bases;
#pragma unroll
        for (int w = 0; w < S; ++w) lowkey[w] = 0;
        if (kbits >= kOwnerLutBits) {
            const uint32_t sh = kbits - kOwnerLutBits;          // bit position of the prefix inside the 64*S-bit number
            const uint32_t wi = S - 1 - sh / 64, bi = sh % 64;
            lowkey[wi] = (uint64_t)p << bi;
            if (bi + kOwnerLutBits > 64 && wi > 0) lowkey[wi - 1] = (uint64_t)p >> (64 - bi);
        }
        uint32_t o = 0;                                         // splitters strictly below lowkey
        while (o < (uint32_t)nshards - 1) {
            uint64_t sp[S];
#pragma unroll
            for (int w = 0; w < S; ++w) sp[w] = spl[o * S + w];
            if (!words_less<S>(sp, lowkey)) break;
            ++o;
        }
        owner_lut[p] = (uint8_t)o;
    }
    const uint64_t ntiles = route_tiles(nq, kTile);
    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        uint32_t *stage = stage_all + (it % depth) * stage_words;
        for (int i = threadIdx.x; i < nshards; i += BLOCK) hist[i] = 0;
        // this tile's staging area was last read by the bulk copies of tile it-depth: their issuers wait for those reads here
        // (the copies of the depth-1 tiles in between stay pending); the barrier below publishes that to the whole CTA
        if (threadIdx.x < (uint32_t)nshards) {
            if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else if (depth == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
        }
        __syncthreads();
        uint64_t q[kRouteQ][S];
        uint32_t own[kRouteQ], rank_in[kRouteQ];
        uint32_t fl[kRouteQ];
        // all loads of the tile are issued before anything depends on them (one exposed round trip per tile)
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint64_t i = tile * kTile + (uint64_t)j * BLOCK + threadIdx.x;
            fl[j] = (i < nq) ? ((flags != nullptr) ? (uint32_t)__ldg(flags + i) : 0u) : 6u;
        }
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint64_t i = tile * kTile + (uint64_t)j * BLOCK + threadIdx.x;
            if (i < nq) load_key<S>(words, i, q[j]);
            else {
#pragma unroll
                for (int w = 0; w < S; ++w) q[j][w] = 0;
            }
        }
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {                // flagged queries are never routed: they cannot match
            own[j] = 0xffffffffu;
            if (!(fl[j] & 6u)) {
                // owner = number of splitters <= q: the prefix table gives those below the prefix, the rest is a short scan
                const uint32_t kbits = 2u * k_bases;
                uint32_t o = 0;
                if (kbits >= kOwnerLutBits) o = owner_lut[(uint32_t)(key_top64<S>(q[j], k_bases) >> (64u - kOwnerLutBits))];
                while (o < (uint32_t)nshards - 1) {
                    uint64_t sp[S];
#pragma unroll
                    for (int w = 0; w < S; ++w) sp[w] = spl[o * S + w];
                    if (words_less<S>(q[j], sp)) break;
                    ++o;
                }
                own[j] = o;
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
        // warps 2-3 reserve the owners' segments space (one global atomic per owner with keys in this tile) while warp 0
        // turns the histogram into the tile's regrouped order
        if (threadIdx.x >= 64 && threadIdx.x < 64u + (uint32_t)nshards) {
            const uint32_t o = threadIdx.x - 64;
            const uint32_t cnt = hist[o];
            // space is reserved in multiples of 4 keys: every run then starts AND ends on a 16-byte boundary of the owner's inbox
            // (4 keys = KW 16-byte units) and leaves as one bulk copy; the up to 3 pad slots hold zero keys that the owner searches
            // and nobody reads back (4-byte stores for ragged run ends doubled the number of NVLink packets)
            const unsigned long long b0 = cnt ? atomicAdd(&cursors[o], (unsigned long long)((cnt + 3u) & ~3u)) : 0ull;
            base[o] = b0;
            rs.tile_base[tile * nshards + o] = (uint32_t)b0;
            rs.tile_cnt[tile * nshards + o] = cnt;
        }
        if (threadIdx.x < 32) {
            const uint32_t e0 = 2 * lane, e1 = e0 + 1;
            const uint32_t v0 = e0 < (uint32_t)nshards ? hist[e0] : 0u, v1 = e1 < (uint32_t)nshards ? hist[e1] : 0u;
            uint32_t p0, p1, tot;
            warp_excl_scan64(v0, v1, lane, p0, p1, tot);
            if (e0 < (uint32_t)nshards) loc[e0] = p0;
            if (e1 < (uint32_t)nshards) loc[e1] = p1;
            if (lane == 0) loc[nshards] = tot;
        }
        __syncthreads();
        // staging layout: the owners' runs back to back, each padded to a multiple of 4 keys (16-byte aligned on both sides, in
        // shared memory and in the destination: rank * cap + base is a multiple of 4 keys)
        if (threadIdx.x < 32) {
            uint32_t len[2], keepw[2];
            uint64_t dw[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t o = 2 * lane + e;
                len[e] = 0; keepw[e] = 0; dw[e] = 0;
                if (o < (uint32_t)nshards) {
                    const uint64_t b0 = base[o];
                    const uint32_t padded = (hist[o] + 3u) & ~3u;
                    dw[e] = ((uint64_t)my_rank * cap + b0) * KW;                    // destination word index in the owner's inbox
                    // a segment holds `cap` keys; keys beyond it are dropped here and reported through sent[o] > cap
                    const uint32_t keep = b0 >= cap ? 0u : (uint32_t)min((unsigned long long)padded, (unsigned long long)(cap - b0));
                    keepw[e] = keep * KW;
                    len[e] = padded * KW;
                }
            }
            uint32_t p[2], tot;
            warp_excl_scan64(len[0], len[1], lane, p[0], p[1], tot);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint32_t o = 2 * lane + e;
                if (o < (uint32_t)nshards) {
                    locw[o] = p[e];
                    endw[o] = p[e] + keepw[e];
                    dptr[o] = static_cast<uint32_t *>(inbox.p[o]) + dw[e] - p[e];
                    for (uint32_t w = p[e] + hist[o] * KW; w < p[e] + len[e]; ++w) stage[w] = 0u;      // the pad keys
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint64_t i = tile * kTile + (uint64_t)j * BLOCK + threadIdx.x;
            if (i >= nq) continue;
            uint32_t at = kNotRouted;
            if (own[j] != 0xffffffffu) {
                at = loc[own[j]] + rank_in[j];                       // position in the tile's regrouped order (for the gather)
                const uint32_t w0 = locw[own[j]] + rank_in[j] * KW;
                key_to_wire<S, KW>(q[j], stage + w0);
            }
            rs.at16[i] = (uint16_t)at;
        }
        // the staged keys are read by the async proxy (bulk copies): order the generic-proxy writes before it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // copy-out: one thread per owner hands its run to the bulk-copy (TMA) engine -- shared memory -> the owner's inbox, over
        // NVLink when the owner is a peer.  The copies are asynchronous: the CTA goes on to the next tile while the link drains this one,
        // and no warp sits in a backed-up store queue (plain stores from all warps left the SMs stalled on the link: the leg
        // took compute + transfer instead of max(compute, transfer)).
        if (threadIdx.x < (uint32_t)nshards) {
            const uint32_t o = threadIdx.x, lo = locw[o], hi = endw[o];             // both multiples of 4 words
            if (hi > lo)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(dptr[o] + lo), "r"(smem_u32(stage + lo)), "r"((hi - lo) * 4u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");       // one group per tile, empty or not: wait_group counts tiles
        }
    }
    if (threadIdx.x < (uint32_t)nshards) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __threadfence_system();
}

// counts_in[my_rank] on every owner := number of keys this rank routed to it (P2P stores of 8 bytes)
__global__ void publish_counts_kernel(const unsigned long long *cursors, int nshards, int my_rank, uint64_t cap, PeerPtrs counts_in) {
    const int o = threadIdx.x;
    if (o < nshards) static_cast<unsigned long long *>(counts_in.p[o])[my_rank] = cursors[o] < cap ? cursors[o] : cap;
    __threadfence_system();
}

// The owner's side.  With vsub > 1 every rank's shard is cut into vsub contiguous sub-ranges ("virtual shards": the route
// kernel simply sees world * vsub owners) and the sub-ranges are searched one after another, so the slice of the key
// column and of the prefix table in use at any time fits L2 -- the partitioned (sort-merge-like) form of the lookup for
// large batches.  inbox: [vsub][world][cap][KW], counts_in: [vsub][world]; ret on the origin: [world * vsub][cap].
template <int S, int KW>
__global__ void __launch_bounds__(kFindBlock, 2) find_routed_kernel(const uint32_t *__restrict__ inbox, const unsigned long long *__restrict__ counts_in,
                                                                    int world, int vsub, uint64_t cap, IndexView ix, uint32_t *__restrict__ res,
                                                                    int bins_in_smem) {
    // all segments in (sub-range, source) order form one flat sequence of keys; the grid sweeps it front to back, so the
    // CTAs work on the same sub-range at the same time
    extern __shared__ __align__(16) uint2 find_bins_smem[];
    __shared__ unsigned long long pre[kMaxShards + 1];
    const uint32_t nseg = (uint32_t)(world * vsub);
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < nseg; ++i) { pre[i] = acc; acc += counts_in[i] < cap ? counts_in[i] : cap; }
        pre[nseg] = acc;
    }
    const uint2 *bins = stage_bins(ix, bins_in_smem ? find_bins_smem : nullptr);
    __syncthreads();
    const uint64_t total = pre[nseg];
    const uint64_t span = (uint64_t)gridDim.x * kFindBlock;
    const uint64_t pol = make_line_policy(ix.hints);
    uint64_t f = (uint64_t)blockIdx.x * kFindBlock + threadIdx.x;
    uint64_t q[S];
    uint64_t slot = 0;                                       // segment * cap + position: same index in the inbox and in res
    bool live = false;
    auto fetch = [&](uint64_t at) {
        live = at < total;
#pragma unroll
        for (int w = 0; w < S; ++w) q[w] = 0;
        if (live) {
            uint32_t lo = 0, hi = nseg - 1;                  // segment e with pre[e] <= at < pre[e + 1]
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (pre[mid] <= at) lo = mid; else hi = mid - 1;
            }
            slot = (uint64_t)lo * cap + (at - pre[lo]);
            wire_to_key<S, KW>(inbox + slot * KW, q);
        }
    };
    fetch(f);
    for (uint64_t w0 = f - (threadIdx.x & 31u); w0 < total; w0 += span, f += span) {
        uint64_t qc[S];
#pragma unroll
        for (int w = 0; w < S; ++w) qc[w] = q[w];
        const bool lc = live;
        const uint64_t sc = slot;
        fetch(f + span);
        const int64_t r = lookup_lines_warp<S, KW>(ix, bins, pol, qc, lc);     // ix.first_index is 0 here: results are local to the shard
        // results stay on the owner, in the inbox's own layout; the origin pulls its runs in the gather leg.  (Storing them
        // straight into the origins' buffers cost up to 1.4 ms per batch on most ranks at 8 GPUs: see DESIGN.md section 5.)
        if (lc) res[sc] = r < 0 ? kWireMiss : (uint32_t)r;
    }
    __threadfence_system();
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) gather_routed_kernel(PeerPtrs res, RouteState rs, uint64_t nq,
                                                              const uint64_t *__restrict__ shard_first, int nshards, uint64_t cap,
                                                              int64_t *__restrict__ out) {
    constexpr uint32_t kTile = BLOCK * kRouteQ;
    __shared__ int64_t staged[kTile];
    __shared__ uint32_t loc[kMaxShards + 1], base[kMaxShards];
    __shared__ uint64_t first[kMaxShards];
    for (int i = threadIdx.x; i < nshards; i += BLOCK) first[i] = shard_first[i];
    const uint64_t ntiles = route_tiles(nq, kTile);
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();                                  // previous tile's staged[] fully consumed
        if (threadIdx.x < (uint32_t)nshards) base[threadIdx.x] = rs.tile_base[tile * nshards + threadIdx.x];
        if (threadIdx.x == 32) {                          // another warp than the base[] loaders
            uint32_t acc = 0;
            for (int i = 0; i < nshards; ++i) { loc[i] = acc; acc += rs.tile_cnt[tile * nshards + i]; }
            loc[nshards] = acc;
        }
        __syncthreads();
        const uint32_t total = loc[nshards];
        // all remote loads of the tile are issued before the first result is used: the NVLink round trip (microseconds) is
        // paid once per tile, not once per loop iteration
        uint32_t own[kRouteQ], val[kRouteQ];
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint32_t p = threadIdx.x + (uint32_t)j * BLOCK;
            own[j] = 0;
            val[j] = kWireMiss;
            if (p < total) {
                uint32_t lo = 0, hi = (uint32_t)nshards - 1;  // owner o with loc[o] <= p < loc[o+1]
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (loc[mid] <= p) lo = mid; else hi = mid - 1;
                }
                own[j] = lo;
                const uint64_t at = (uint64_t)base[lo] + (p - loc[lo]);
                // res.p[o] = this origin's result segment on owner o (peer memory: read over NVLink, never cached in L1)
                if (at < cap) val[j] = __ldcv(static_cast<const uint32_t *>(res.p[lo]) + at);   // >= cap: dropped by an overflowing segment
            }
        }
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint32_t p = threadIdx.x + (uint32_t)j * BLOCK;
            if (p < total) staged[p] = val[j] == kWireMiss ? -1 : (int64_t)(first[own[j]] + val[j]);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kRouteQ; ++j) {
            const uint64_t i = tile * kTile + (uint64_t)j * BLOCK + threadIdx.x;
            if (i < nq) {
                const uint32_t at = rs.at16[i];
                out[i] = at == kNotRouted ? -1 : staged[at];
            }
        }
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

// key_start[b] = number of keys whose bin is < b  (b in [0, nbins]).
template <int S>
__global__ void bin_bounds_kernel(IndexView ix, uint32_t *key_start) {
    const uint64_t n = ix.n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t prev = -1, cur = (int64_t)ix.nbins;
        uint32_t bin; uint64_t frac;
        if (i > 0) { uint64_t a[S]; load_key<S>(ix.keys, i - 1, a); prev = key_bin<S>(ix, a, bin, frac) ? (int64_t)bin : (int64_t)ix.nbins - 1; }
        if (i < n) { uint64_t b[S]; load_key<S>(ix.keys, i, b); cur = key_bin<S>(ix, b, bin, frac) ? (int64_t)bin : (int64_t)ix.nbins - 1; }
        for (int64_t b = prev + 1; b <= cur; ++b) key_start[b] = (uint32_t)i;
    }
}

// bins[b] = {first line, lines} with lines = max(1, ceil(keys in bin * 16 / fill_x16)); *total = all lines.  One CTA.
__global__ void __launch_bounds__(1024) bins_scan_kernel(const uint32_t *__restrict__ key_start, uint32_t nbins, uint32_t fill_x16,
                                                         uint2 *__restrict__ bins, unsigned long long *total) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (nbins + 1023u) / 1024u;
    const uint32_t b0 = min(nbins, threadIdx.x * per), b1 = min(nbins, b0 + per);
    auto lines_of = [&](uint32_t b) -> unsigned long long {
        const unsigned long long cnt = key_start[b + 1] - key_start[b];
        const unsigned long long nl = (cnt * 16ull + fill_x16 - 1ull) / fill_x16;
        return nl ? nl : 1ull;
    };
    unsigned long long acc = 0;
    for (uint32_t b = b0; b < b1; ++b) acc += lines_of(b);
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; ++t) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    unsigned long long at = part[threadIdx.x];
    for (uint32_t b = b0; b < b1; ++b) {
        const unsigned long long nl = lines_of(b);
        bins[b] = make_uint2((uint32_t)at, (uint32_t)nl);
        at += nl;
    }
}

// Writes every line: thread i owns the lines in (line of key i-1, line of key i] -- the empty ones in between and, when
// key i is the first of its line, that line with up to CAP keys.
template <int S, int KW>
__global__ void fill_lines_kernel(IndexView ix, uint32_t *__restrict__ lines, uint64_t nlines) {
    using LL = LineLayout<KW>;
    const uint64_t n = ix.n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t prev = -1, cur = (int64_t)nlines;
        uint32_t l;
        uint64_t kq[S];
        if (i > 0) { load_key<S>(ix.keys, i - 1, kq); prev = key_line<S>(ix, ix.bins, kq, l) ? (int64_t)l : (int64_t)nlines - 1; }
        if (i < n) { load_key<S>(ix.keys, i, kq); cur = key_line<S>(ix, ix.bins, kq, l) ? (int64_t)l : (int64_t)nlines - 1; }
        uint32_t w[kLineWords];
        for (int64_t e = prev + 1; e < cur + (i < n && cur != prev ? 1 : 0); ++e) {
#pragma unroll
            for (uint32_t x = 0; x < kLineWords; ++x) w[x] = 0;
            w[LL::kBaseWord] = (uint32_t)i;
#pragma unroll
            for (int t = 0; t < LL::CAP; ++t) {
#pragma unroll
                for (int p = 0; p < KW; ++p) w[LL::word(t, p)] = ix.pad[p];
            }
            if (e == cur) {                         // key i opens this line: it and its followers that map to the same line
                bool more = true;
#pragma unroll
                for (int t = 0; t < LL::CAP; ++t) {
                    if (more && i + t < n) {
                        uint64_t kk[S];
                        uint32_t lt;
                        load_key<S>(ix.keys, i + t, kk);
                        more = t == 0 || ((key_line<S>(ix, ix.bins, kk, lt) ? (int64_t)lt : (int64_t)nlines - 1) == cur);
                        if (more) {
                            uint32_t kw[KW];
                            key_to_wire<S, KW>(kk, kw);
#pragma unroll
                            for (int p = 0; p < KW; ++p) w[LL::word(t, p)] = kw[p];
                        }
                    } else more = false;
                }
            }
            uint4 *dst = reinterpret_cast<uint4 *>(lines + (uint64_t)e * kLineWords);
#pragma unroll
            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
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
// ------------------------------------------------------------------ sorted-merge lookup (CC_ALGO_MERGE)
// For a batch in ascending key order (the records of another graph: FindShared, RecoverExcludedKmers, Join-like callers; or any
// batch after a radix sort) the lookup is a merge of two sorted arrays.  The batch is cut into tiles of kMergeTile queries; a
// first kernel finds, by binary search, where each tile's first query falls in the key column (merge-path partition: one search
// per tile instead of one per query); a tile's queries can then only match keys inside its window [b[t], b[t+1]], which is
// staged in shared memory once and searched there (a batch denser than the table has windows of a few hundred keys), or searched
// in place when it is too long to stage (a sparse batch: the window only narrows the binary search).  Queries and results
// stream through once, coalesced when the batch arrived sorted: 8s + 8 bytes per query + the key column once.
constexpr uint32_t kMergeTile = 1024;         // queries per tile (4 per thread)
constexpr uint32_t kMergeWin = 2048;          // keys staged per window

template <int S>
__global__ void sorted_check_kernel(const uint64_t *__restrict__ words, uint64_t nq, unsigned int *unsorted) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < nq; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t a[S], b[S];
        load_key<S>(words, i - 1, a);
        load_key<S>(words, i, b);
        if (words_less<S>(b, a)) *unsorted = 1u;
    }
}

// bounds[t] = lower_bound(keys, first query of tile t) for t < ntiles; bounds[ntiles] = lower_bound(keys, last query).
template <int S>
__global__ void merge_bounds_kernel(const uint64_t *__restrict__ words, const uint32_t *__restrict__ perm, uint64_t nq, uint64_t ntiles,
                                    const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ bounds) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = t < ntiles ? t * kMergeTile : nq - 1;
        uint64_t q[S];
        load_key<S>(words, perm ? perm[j] : j, q);
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            uint64_t km[S];
            load_key<S>(keys, mid, km);
            if (words_less<S>(km, q)) lo = mid + 1; else hi = mid;
        }
        bounds[t] = lo;
    }
}

template <int S>
__global__ void __launch_bounds__(kBlock) find_merge_kernel(const uint64_t *__restrict__ words, const uint8_t *__restrict__ flags,
                                                            const uint32_t *__restrict__ perm, uint64_t nq, uint64_t ntiles,
                                                            const uint64_t *__restrict__ bounds, IndexView ix, int64_t *__restrict__ out_index) {
    extern __shared__ __align__(16) uint64_t merge_win[];           // [kMergeWin][S]
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t lo = bounds[t], hi = min(bounds[t + 1] + 1, ix.n);
        const uint32_t w = (uint32_t)min(hi - lo, (uint64_t)kMergeWin + 1);
        const bool staged = w <= kMergeWin;
        __syncthreads();                                            // the previous tile's window is no longer read
        if (staged) {
            for (uint32_t i = threadIdx.x; i < w * S; i += kBlock) merge_win[i] = __ldg(ix.keys + lo * S + i);
        }
        __syncthreads();
        // thread i takes kMergeTile / kBlock CONSECUTIVE queries of the sorted order: one binary search for the first, the
        // others continue from where their predecessor stopped (ascending queries only ever move forward in the window)
        constexpr int PER = (int)(kMergeTile / kBlock);
        uint32_t at = 0;
        bool have_at = false;
#pragma unroll
        for (int h = 0; h < PER; ++h) {
            const uint64_t j = t * kMergeTile + (uint64_t)threadIdx.x * PER + h;
            if (j >= nq) continue;
            const uint64_t src = perm ? perm[j] : j;
            uint64_t q[S];
            load_key<S>(words, src, q);
            int64_t r = -1;
            if (staged) {
                if (!have_at) {
                    uint32_t a = 0, b = w;                          // first window key >= q
                    while (a < b) {
                        const uint32_t mid = (a + b) >> 1;
                        uint64_t km[S];
#pragma unroll
                        for (int x = 0; x < S; ++x) km[x] = merge_win[mid * S + x];
                        if (words_less<S>(km, q)) a = mid + 1; else b = mid;
                    }
                    at = a;
                    have_at = true;
                }
                uint64_t km[S];
                while (at < w) {                                    // the merge step: advance to the first key >= q
#pragma unroll
                    for (int x = 0; x < S; ++x) km[x] = merge_win[at * S + x];
                    if (!words_less<S>(km, q)) break;
                    ++at;
                }
                if (at < w && words_equal<S>(km, q)) r = (int64_t)(lo + at + ix.first_index);
            } else {
                const int64_t f = search_range<S>(ix.keys, lo, hi, q);
                r = f < 0 ? f : f + (int64_t)ix.first_index;
            }
            if (flags && (__ldg(flags + src) & 6u)) r = -1;
            out_index[src] = r;
        }
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

// Grid of a persistent (grid-stride) kernel: exactly the CTAs that are resident at once, so that no second, partial wave
// trails behind (the occupancy comes from the runtime, not from a guess about registers).
template <typename Kernel>
int resident_grid(Kernel kernel, int block, size_t smem, uint64_t work_blocks, int sm_count, int max_per_sm = 32) {
    int per = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kernel, block, smem) != cudaSuccess || per < 1) per = 1;
    per = std::min(per, max_per_sm);
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(work_blocks, (uint64_t)sm_count * per));
}

int sm_count_now() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

IndexView view_of(const cc_graph *g) {
    IndexView v{};
    v.keys = g->index.keys;
    v.lines = static_cast<const uint4 *>(g->index.lines);
    v.bins = static_cast<const uint2 *>(g->index.bins);
    v.n = g->h.num_records;
    v.first_index = g->first_index;
    v.base = g->index.base;
    v.span = g->index.span;
    v.scale = g->index.scale;
    v.nbins = g->index.nbins;
    v.norm = g->index.norm;
    // a table of a few hundred MB gets a little help from L2 (about 18 % hits at 320 MB: each die caches its own copy) and loses it
    // when its lines are marked evict-first; a multi-GB table is touched at random and never again
    v.hints = options().lookup_l2_hints >= 0 ? (uint32_t)options().lookup_l2_hints : (g->index.nlines * 64ull > (1ull << 30) ? 1u : 0u);
    v.k = g->h.k;
    for (int i = 0; i < 8; ++i) v.pad[i] = g->index.pad[i];
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

#define CC_DISPATCH_SKW(s, kw, ...)                                                        \
    switch ((s) * 2 - (kw)) {                                                              \
        case 0: CC_DISPATCH_S(s, { constexpr int KW_ = 2 * S_; __VA_ARGS__; }) break;      \
        case 1: CC_DISPATCH_S(s, { constexpr int KW_ = 2 * S_ - 1; __VA_ARGS__; }) break;  \
        default: return fail(CC_ERR_ARG, "inconsistent k-mer size for the wire format");   \
    }

inline uint32_t wire_words(uint32_t k) { return (2 * k + 31) / 32; }

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
        IndexView none{};
        // tile = RPT * 256 rows, sized to about 16 KB of sequence (two tiles in flight per CTA, several CTAs per SM)
#define CC_ROWS_PACK(RPT_) CC_DISPATCH_S(s, {                                                                                  \
            const size_t smem = rows_smem_bytes(RPT_ * kBlock, k, false);                                                          \
            CC_CUDA(cudaFuncSetAttribute(rows_kernel<S_, 2 * S_, false, RPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
            const int grid = resident_grid(rows_kernel<S_, 2 * S_, false, RPT_>, kBlock, smem, (nq + RPT_ * kBlock - 1) / (RPT_ * kBlock), sm_count_now()); \
            rows_kernel<S_, 2 * S_, false, RPT_><<<grid, kBlock, smem, st>>>(dev_seq, nq, k, dev_words, dev_flags, none, nullptr); })
        if (options().rows_warp) {
#define CC_ROWS_WARP(NRAW_) CC_DISPATCH_S(s, {                                                                                 \
                const size_t smem = rows_warp_smem_bytes(k, NRAW_);                                                                 \
                CC_CUDA(cudaFuncSetAttribute(rows_warp_kernel<S_, 2 * S_, false, NRAW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                const int grid = resident_grid(rows_warp_kernel<S_, 2 * S_, false, NRAW_>, kBlock, smem, (nq + kBlock - 1) / kBlock, sm_count_now()); \
                rows_warp_kernel<S_, 2 * S_, false, NRAW_><<<grid, kBlock, smem, st>>>(dev_seq, nq, k, dev_words, dev_flags, kRowsMul, none, nullptr); })
            if (options().rows_warp == 3) { CC_ROWS_WARP(3); } else { CC_ROWS_WARP(2); }
#undef CC_ROWS_WARP
        } else if (k <= (uint32_t)options().rows_rpt2_max_k) { CC_ROWS_PACK(2); } else { CC_ROWS_PACK(1); }
#undef CC_ROWS_PACK
        count_launch();
        CC_CUDA(cudaGetLastError());
        return CC_OK;
    }
    SeqJob job = make_job(dev_seq, nq, row_stride, k);
    const uint64_t ntiles = (nq + job.per_tile - 1) / job.per_tile;
    IndexView none{};
    CC_DISPATCH_S(s, {
        const int grid = resident_grid(seq_kernel<S_, 2 * S_, SEQ_PACK>, kBlock, 0, ntiles, sm_count_now());
        seq_kernel<S_, 2 * S_, SEQ_PACK><<<grid, kBlock, 0, st>>>(job, dev_words, dev_flags, none, nullptr);
    });
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
    if (row_stride == g->h.k && row_stride > 1) {
        if (algo == CC_ALGO_AUTO && options().rows_fused) {
            // independent rows: pack and search in one kernel (the stream of 2-bit codes is all that is staged)
            IndexView ix = view_of(g);
            const uint32_t k = g->h.k;
#define CC_ROWS_FIND(RPT_) CC_DISPATCH_SKW(g->h.s, wire_words(k), {                                                            \
                const size_t smem = rows_smem_bytes(RPT_ * kBlock, k, true);                                                       \
                CC_CUDA(cudaFuncSetAttribute(rows_kernel<S_, KW_, true, RPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                const int grid = resident_grid(rows_kernel<S_, KW_, true, RPT_>, kBlock, smem, (nq + RPT_ * kBlock - 1) / (RPT_ * kBlock), g->sm_count); \
                rows_kernel<S_, KW_, true, RPT_><<<grid, kBlock, smem, st>>>(dev_seq, nq, k, nullptr, nullptr, ix, dev_index); })
            if (options().rows_warp) {
                // warp-autonomous form: no CTA barrier between the conversion and the searches, the occupancy hides the line loads
                CC_DISPATCH_SKW(g->h.s, wire_words(k), {
                    const size_t smem = rows_warp_smem_bytes(k, 2);
                    CC_CUDA(cudaFuncSetAttribute(rows_warp_kernel<S_, KW_, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const int grid = resident_grid(rows_warp_kernel<S_, KW_, true, 2>, kBlock, smem, (nq + kBlock - 1) / kBlock, g->sm_count);
                    rows_warp_kernel<S_, KW_, true, 2><<<grid, kBlock, smem, st>>>(dev_seq, nq, k, nullptr, nullptr, kRowsMul, ix, dev_index);
                });
            } else if (k <= 64) { CC_ROWS_FIND(2); } else { CC_ROWS_FIND(1); }     // two lookups in flight per thread when the tile fits
#undef CC_ROWS_FIND
            count_launch();
            CC_CUDA(cudaGetLastError());
            return CC_OK;
        }
        uint64_t *words = nullptr; uint8_t *flags = nullptr;
        CC_CUDA(cudaMallocAsync(&words, nq * g->h.s * sizeof(uint64_t), st));
        CC_CUDA(cudaMallocAsync(&flags, nq, st));
        int rc = launch_pack_windows(dev_seq, 0, g->h.k, words, flags, row_stride, nq, st);
        if (!rc) rc = launch_find_packed(g, words, flags, nq, dev_index, algo, st);
        cudaFreeAsync(words, st); cudaFreeAsync(flags, st);
        return rc;
    }
    SeqJob job = make_job(dev_seq, nq, row_stride, g->h.k);
    const uint64_t ntiles = (nq + job.per_tile - 1) / job.per_tile;
    IndexView ix = view_of(g);
    if (algo == CC_ALGO_BSEARCH) {
        CC_DISPATCH_S(g->h.s, {
            const int grid = resident_grid(seq_kernel<S_, 2 * S_, SEQ_BSEARCH>, kBlock, 0, ntiles, g->sm_count);
            seq_kernel<S_, 2 * S_, SEQ_BSEARCH><<<grid, kBlock, 0, st>>>(job, nullptr, nullptr, ix, dev_index);
        });
    } else {
        CC_DISPATCH_SKW(g->h.s, wire_words(g->h.k), {
            const int grid = resident_grid(seq_kernel<S_, KW_, SEQ_FIND>, kBlock, 0, ntiles, g->sm_count);
            seq_kernel<S_, KW_, SEQ_FIND><<<grid, kBlock, 0, st>>>(job, nullptr, nullptr, ix, dev_index);
        });
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

// Grid and dynamic shared memory of the persistent line-search kernels: the bin table goes to shared memory when the
// batch is large enough to pay for staging it in every CTA.
struct FindLaunch { int grid; size_t smem; int bins_in_smem; };
template <typename Kernel>
int plan_find(Kernel kernel, const cc_graph *g, uint64_t work, FindLaunch &fl) {
    const size_t bins_bytes = (size_t)g->index.nbins * sizeof(uint2);
    fl.bins_in_smem = options().find_bins_smem && work >= (1u << 16) ? 1 : 0;
    fl.smem = fl.bins_in_smem ? bins_bytes : 0;
    if (fl.smem > 48 * 1024) CC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fl.smem));
    fl.grid = resident_grid(kernel, kFindBlock, fl.smem, (work + kFindBlock - 1) / kFindBlock, g->sm_count);
    return CC_OK;
}

int launch_find_small(cc_graph *g, const uint8_t *dev_kmers, uint32_t nq, int64_t *dev_index, uint8_t *dev_raw, cudaStream_t st) {
    if (int rc = check_k(g->h.k)) return rc;
    if (nq == 0) return CC_OK;
    IndexView ix = view_of(g);
    CC_DISPATCH_SKW(g->h.s, wire_words(g->h.k), {
        find_small_kernel<S_, KW_><<<(nq + 7) / 8, 256, 0, st>>>(dev_kmers, nq, g->h.k, ix, g->dev_body, (uint32_t)g->h.record_size, dev_index, dev_raw);
    });
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_find_packed(cc_graph *g, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, int64_t *dev_index,
                       int algo, cudaStream_t st) {
    if (int rc = check_k(g->h.k)) return rc;
    if (nq == 0) return CC_OK;
    IndexView ix = view_of(g);
    const uint32_t s = g->h.s, kw = wire_words(g->h.k);
    const int grid = grid_for(nq, kBlock, g->sm_count, 8);
    if (algo == CC_ALGO_MERGE) {
        // is the batch already in ascending order?  (one pass + one host round trip: this mode is synchronous up to here)
        unsigned int *d_unsorted = nullptr, h_unsorted = 0;
        CC_CUDA(cudaMallocAsync(&d_unsorted, sizeof(unsigned int), st));
        CC_CUDA(cudaMemsetAsync(d_unsorted, 0, sizeof(unsigned int), st));
        CC_DISPATCH_S(s, sorted_check_kernel<S_><<<grid, kBlock, 0, st>>>(dev_words, nq, d_unsorted));
        count_launch();
        CC_CUDA(cudaMemcpyAsync(&h_unsorted, d_unsorted, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
        cudaFreeAsync(d_unsorted, st);
        uint32_t *perm = nullptr;
        if (h_unsorted)
            if (int rc = sort_permutation(dev_words, nq, s, g->h.k, st, &perm)) return rc;
        const uint64_t ntiles = (nq + kMergeTile - 1) / kMergeTile;
        uint64_t *bounds = nullptr;
        CC_CUDA(cudaMallocAsync(&bounds, (ntiles + 1) * sizeof(uint64_t), st));
        const int bgrid = grid_for(ntiles + 1, kBlock, g->sm_count, 8);
        const size_t msmem = (size_t)kMergeWin * s * sizeof(uint64_t);
        CC_DISPATCH_S(s, {
            merge_bounds_kernel<S_><<<bgrid, kBlock, 0, st>>>(dev_words, perm, nq, ntiles, ix.keys, ix.n, bounds);
            if (msmem > 48 * 1024) CC_CUDA(cudaFuncSetAttribute(find_merge_kernel<S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
            const int mgrid = resident_grid(find_merge_kernel<S_>, kBlock, msmem, ntiles, g->sm_count);
            find_merge_kernel<S_><<<mgrid, kBlock, msmem, st>>>(dev_words, dev_flags, perm, nq, ntiles, bounds, ix, dev_index);
        });
        count_launch(2);
        cudaFreeAsync(bounds, st);
        if (perm) cudaFreeAsync(perm, st);
    } else if (algo == CC_ALGO_BSEARCH) {
        CC_DISPATCH_S(s, find_bsearch_kernel<S_><<<grid, kBlock, 0, st>>>(dev_words, dev_flags, nq, ix, dev_index));
        count_launch();
    } else {
        CC_DISPATCH_SKW(s, kw, {
            FindLaunch fl;
            if (int rc = plan_find(find_packed_lines_kernel<S_, KW_>, g, nq, fl)) return rc;
            find_packed_lines_kernel<S_, KW_><<<fl.grid, kFindBlock, fl.smem, st>>>(dev_words, dev_flags, nq, ix, dev_index, fl.bins_in_smem);
        });
        count_launch();
    }
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int build_index(cc_graph *g, int bits_req) {
    if (int rc = check_k(g->h.k)) return rc;
    const uint64_t n = g->h.num_records;
    const uint32_t s = g->h.s, k = g->h.k, kw = wire_words(k);
    if (n >= 0xffffffffull) return fail(CC_ERR_UNSUPPORTED, "more than 2^32-2 records per device shard");
    LookupIndex &ix = g->index;
    if (ix.keys) { cudaFree(ix.keys); ix.keys = nullptr; }
    if (ix.lines) { cudaFree(ix.lines); ix.lines = nullptr; }
    if (ix.bins) { cudaFree(ix.bins); ix.bins = nullptr; }
    ix.built = false;
    ix.nlines = 0;
    cudaStream_t st = g->stream;

    if (int rc = g->scan_ws.ensure(0, 0)) return rc;
    CC_CUDA(cudaMalloc(&ix.keys, std::max<uint64_t>(n * s, 2) * sizeof(uint64_t) + 64));
    if (int rc = launch_decode_columns(g->dev_body, n, s, g->h.c, ix.keys, nullptr, nullptr, g->scan_ws, g->sm_count, st)) return rc;

    // The bins span [first key, last key] of this array, not the whole 2k-bit key space (see key_bin).
    uint64_t first[4] = {0, 0, 0, 0}, last[4] = {0, 0, 0, 0};
    if (n) {
        CC_CUDA(cudaMemcpyAsync(first, ix.keys, s * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaMemcpyAsync(last, ix.keys + (n - 1) * s, s * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
    }
    uint64_t t0 = 0, t1 = 0;
    CC_DISPATCH_S(s, { t0 = key_top64<S_>(first, k); t1 = key_top64<S_>(last, k); });
    const uint64_t span = t1 >= t0 ? t1 - t0 : 0;      // unsorted arrays are rejected below
    // bins: about 256 keys each, a power of two, at most 2^13 (64 KB of shared memory in the search kernels)
    int bits = bits_req > 0 ? bits_req : options().index_bits;
    if (bits <= 0) {
        bits = 0;
        while (bits < (int)kMaxBinsLog2 && (256ull << (bits + 1)) <= n) ++bits;
    }
    bits = std::min<int>(bits, (int)kMaxBinsLog2);
    uint64_t nb = 1ull << bits;
    if (span + 1 != 0 && nb > span + 1) nb = span + 1;  // no finer than the key resolution
    // bin(d) = floor(d * nbins / (span + 1)), evaluated as a 64x64 multiply of the normalised distance (top bit of span at
    // bit 63) with scale = floor(2^64 * nbins / (span' + 1)); the low half of the product is the position inside the bin.
    uint32_t norm = 0;
    while (norm < 63 && !((span << norm) >> 63)) ++norm;
    if (span == 0) norm = 0;
    const unsigned __int128 spanp1 = (unsigned __int128)(span << norm) + 1;
    const unsigned __int128 sc = (((unsigned __int128)nb) << 64) / spanp1;
    ix.base = t0;
    ix.span = span;
    ix.norm = norm;
    ix.scale = sc > (unsigned __int128)~0ull ? ~0ull : (uint64_t)sc;
    ix.nbins = (uint32_t)nb;
    for (int i = 0; i < 8; ++i) ix.pad[i] = 0;
    for (uint32_t j = 0; j < kw; ++j) {                 // wire form of the largest key: the 32-bit words, most significant first
        const uint32_t idx = kw - 1 - j;
        const uint64_t w = last[s - 1 - idx / 2];
        ix.pad[j] = (idx & 1) ? (uint32_t)(w >> 32) : (uint32_t)w;
    }

    unsigned long long *d_unsorted = reinterpret_cast<unsigned long long *>(g->scan_ws.totals + 8);
    unsigned long long *d_total = reinterpret_cast<unsigned long long *>(g->scan_ws.totals + 9);
    const unsigned long long none = ~0ull;
    CC_CUDA(cudaMemcpyAsync(d_unsorted, &none, 8, cudaMemcpyHostToDevice, st));
    const int grid = grid_for(n + 1, 256, g->sm_count, 8);
    CC_DISPATCH_S(s, check_sorted_kernel<S_><<<grid, 256, 0, st>>>(ix.keys, n, d_unsorted));
    count_launch();
    CC_CUDA(cudaMalloc(&ix.bins, ((uint64_t)ix.nbins + 1) * sizeof(uint2)));
    uint32_t *key_start = nullptr;
    CC_CUDA(cudaMallocAsync(&key_start, ((uint64_t)ix.nbins + 2) * sizeof(uint32_t), st));
    IndexView view = view_of(g);
    CC_DISPATCH_S(s, bin_bounds_kernel<S_><<<grid, 256, 0, st>>>(view, key_start));
    const int fill_pct = std::min(100, std::max(5, options().index_fill_pct));
    uint32_t cap_keys = 0;
    CC_DISPATCH_SKW(s, kw, cap_keys = LineLayout<KW_>::CAP);
    const uint32_t fill_x16 = std::max<uint32_t>(1, cap_keys * 16u * (uint32_t)fill_pct / 100u);
    bins_scan_kernel<<<1, 1024, 0, st>>>(key_start, ix.nbins, fill_x16, static_cast<uint2 *>(ix.bins), d_total);
    count_launch(2);
    unsigned long long at = none, nlines = 0;
    CC_CUDA(cudaMemcpyAsync(&at, d_unsorted, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaMemcpyAsync(&nlines, d_total, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    cudaFreeAsync(key_start, st);
    if (nlines >= 0xffffffffull) return fail(CC_ERR_UNSUPPORTED, "lookup index of %llu lines exceeds 2^32-2", nlines);
    ix.nlines = nlines;
    CC_CUDA(cudaMalloc(&ix.lines, std::max<uint64_t>(nlines, 1) * kLineWords * sizeof(uint32_t)));
    view = view_of(g);
    if (at == none) {                                   // an unsorted array has no order-preserving table; lookups are refused
        CC_DISPATCH_SKW(s, kw, fill_lines_kernel<S_, KW_><<<grid, 256, 0, st>>>(view, static_cast<uint32_t *>(ix.lines), nlines));
        count_launch();
    }
    CC_CUDA(cudaGetLastError());
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

uint64_t route_state_size(uint64_t max_q, int nshards) { return route_state_bytes(max_q, nshards); }

int launch_route(const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t k, const uint64_t *dev_splitters, int nshards,
                 int my_rank, uint64_t cap, void *const *peer_inbox, void *const *peer_counts, void *dev_route_state, uint64_t max_q,
                 uint64_t *dev_sent, cudaStream_t st) {
    if (int rc = check_k(k)) return rc;
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    if (my_rank < 0 || my_rank >= nshards) return fail(CC_ERR_ARG, "rank %d out of range", my_rank);
    if (nq >= (1ull << 32) || cap >= (1ull << 32)) return fail(CC_ERR_UNSUPPORTED, "routed batches are limited to 2^32-1 queries per rank");
    const uint32_t s = (k + 31) / 32, kw = wire_words(k);
    PeerPtrs inbox{}, counts{};
    for (int i = 0; i < nshards; ++i) {
        inbox.p[i] = peer_inbox[i]; counts.p[i] = peer_counts[i];
        if ((reinterpret_cast<uintptr_t>(peer_inbox[i]) & 15u) || (cap & 3u))
            return fail(CC_ERR_ARG, "inbox segments must be 16-byte aligned and cap a multiple of 4 keys");
    }
    unsigned long long *cursors = reinterpret_cast<unsigned long long *>(dev_sent);
    CC_CUDA(cudaMemsetAsync(cursors, 0, sizeof(uint64_t) * nshards, st));
    if (nq) {
        if (nq > max_q) return fail(CC_ERR_ARG, "batch of %llu queries exceeds the route state sized for %llu", (unsigned long long)nq, (unsigned long long)max_q);
        const RouteState rs = route_state_of(dev_route_state, max_q, nshards);
        const int block = route_block_for(nshards);
        const uint32_t tile_q = (uint32_t)block * kRouteQ, stage_words = route_stage_words(tile_q, kw, nshards);
        const int depth = std::min(4, std::max(2, options().route_stage_depth));
        const size_t smem = route_smem_bytes(tile_q, kw, nshards, depth);
        const int per_sm = options().route_blocks_per_sm > 0 ? options().route_blocks_per_sm : 32;
#define CC_ROUTE(BLOCK_) CC_DISPATCH_SKW(s, kw, {                                                                                   \
            CC_CUDA(cudaFuncSetAttribute(route_kernel<S_, KW_, BLOCK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
            const int grid = resident_grid(route_kernel<S_, KW_, BLOCK_>, BLOCK_, smem, route_tiles(nq, tile_q), sm_count_now(), per_sm); \
            route_kernel<S_, KW_, BLOCK_><<<grid, BLOCK_, smem, st>>>(dev_words, dev_flags, nq, dev_splitters, nshards, my_rank, cap, inbox, rs, \
                                                                      cursors, stage_words, k, (uint32_t)depth); })
        if (block == 256) { CC_ROUTE(256); } else { CC_ROUTE(128); }
#undef CC_ROUTE
        count_launch();
    }
    publish_counts_kernel<<<1, kMaxShards, 0, st>>>(cursors, nshards, my_rank, cap, counts);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_publish_counts(const uint64_t *dev_sent, int nshards, int my_rank, uint64_t cap, void *const *peer_counts, cudaStream_t st) {
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    PeerPtrs counts{};
    for (int i = 0; i < nshards; ++i) counts.p[i] = peer_counts[i];
    publish_counts_kernel<<<1, kMaxShards, 0, st>>>(reinterpret_cast<const unsigned long long *>(dev_sent), nshards, my_rank, cap, counts);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_find_routed(cc_graph *g, const void *dev_inbox, const uint64_t *dev_counts_in, int world, int vsub, uint64_t cap,
                       void *dev_res, cudaStream_t st) {
    if (int rc = check_k(g->h.k)) return rc;
    if (world < 1 || vsub < 1 || world * vsub > kMaxShards) return fail(CC_ERR_ARG, "world * vsub must be in 1..%d", kMaxShards);
    IndexView ix = view_of(g);
    ix.first_index = 0;                 // the wire carries indices local to the shard; the origin rebases them
    const uint32_t kw = wire_words(g->h.k);
    CC_DISPATCH_SKW(g->h.s, kw, {
        FindLaunch fl;
        if (int rc = plan_find(find_routed_kernel<S_, KW_>, g, 1ull << 20, fl)) return rc;
        if (options().routed_search_blocks_per_sm > 0) fl.grid = std::min(fl.grid, g->sm_count * options().routed_search_blocks_per_sm);
        find_routed_kernel<S_, KW_><<<fl.grid, kFindBlock, fl.smem, st>>>(static_cast<const uint32_t *>(dev_inbox),
                                                                          reinterpret_cast<const unsigned long long *>(dev_counts_in), world, vsub,
                                                                          cap, ix, static_cast<uint32_t *>(dev_res), fl.bins_in_smem);
    });
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_gather_routed(void *const *peer_res, const void *dev_route_state, uint64_t max_q, uint64_t nq, const uint64_t *dev_shard_first,
                         int nshards, uint64_t cap, int64_t *dev_out, cudaStream_t st) {
    if (nshards < 1 || nshards > kMaxShards) return fail(CC_ERR_ARG, "nshards must be in 1..%d", kMaxShards);
    if (nq == 0) return CC_OK;
    const RouteState rs = route_state_of(const_cast<void *>(dev_route_state), max_q, nshards);
    PeerPtrs res{};
    for (int i = 0; i < nshards; ++i) res.p[i] = peer_res[i];
    const int cap_sm = std::max(1, options().gather_blocks_per_sm);
    if (route_block_for(nshards) == 256) {
        const int grid = resident_grid(gather_routed_kernel<256>, 256, 0, route_tiles(nq, 256 * kRouteQ), sm_count_now(), cap_sm);
        gather_routed_kernel<256><<<grid, 256, 0, st>>>(res, rs, nq, dev_shard_first, nshards, cap, dev_out);
    } else {
        const int grid = resident_grid(gather_routed_kernel<128>, 128, 0, route_tiles(nq, 128 * kRouteQ), sm_count_now(), cap_sm);
        gather_routed_kernel<128><<<grid, 128, 0, st>>>(res, rs, nq, dev_shard_first, nshards, cap, dev_out);
    }
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

}  // namespace cc
