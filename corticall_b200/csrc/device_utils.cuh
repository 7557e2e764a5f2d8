// device_utils.cuh -- sm_100a building blocks shared by the kernels: mbarrier + 1-D bulk (TMA) copies,
// relaxed GPU-scope descriptor loads/stores for the decoupled look-back, unaligned shared-memory reads,
// 2-bit k-mer arithmetic.  Hand-written PTX; no CUTLASS/CuTe dependency.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cc {

// ------------------------------------------------------------------ error reporting from device code
// Kernels never spin forever: every wait is bounded by a wall-clock watchdog that records a code in
// *err and traps, so a logic error shows up as a CUDA error instead of a hung GPU.
enum DeviceError : int { DEV_OK = 0, DEV_TIMEOUT_FULL = 1, DEV_TIMEOUT_EMPTY = 2, DEV_TIMEOUT_LOOKBACK = 3 };

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr uint64_t kWatchdogNs = 4000000000ull;   // 4 s

static __device__ __noinline__ void watchdog_fail(int *err, int code) {
    if (err) *reinterpret_cast<volatile int *>(err) = code;      // may live in mapped host memory: a plain store
    __threadfence_system();
    __trap();
}

// ------------------------------------------------------------------ shared-memory addresses
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrival without release semantics: for "I am done READING this buffer" hand-backs, where the values read have
// already been consumed; it does not wait for the warp's outstanding global stores to drain.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t *bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *err, int code) {
    // try_wait suspends the warp in hardware for a bounded time, so the fast path is a handful of
    // instructions; the wall-clock watchdog is consulted only every 4096 failed probes.
#pragma unroll 1
    for (uint32_t spins = 0; spins < 4096u; ++spins) {
        if (mbar_try_wait(bar, parity)) return;
    }
    const uint64_t t0 = globaltimer_ns();
    while (true) {
#pragma unroll 1
        for (uint32_t spins = 0; spins < 4096u; ++spins) {
            if (mbar_try_wait(bar, parity)) return;
        }
        if (globaltimer_ns() - t0 > kWatchdogNs) watchdog_fail(err, code);
    }
}

// ------------------------------------------------------------------ 1-D bulk async copy (TMA engine, SASS UBLKCP)
// dst: shared, 16-byte aligned; src: global, 16-byte aligned; bytes: multiple of 16.
// Streaming data is read exactly once: L2 evict-first policy keeps it from displacing resident tables.
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// ------------------------------------------------------------------ named barrier among a subset of warps
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ look-back descriptors (relaxed, GPU scope)
__device__ __forceinline__ void st_relaxed_gpu(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_gpu(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------ unaligned shared-memory reads
// Reads may touch up to 7 bytes past the last requested byte; stage buffers carry slack for that.
template <bool ALIGNED4>
__device__ __forceinline__ uint32_t lds_u32(const uint8_t *p) {
    if (ALIGNED4) return *reinterpret_cast<const uint32_t *>(p);
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    return __funnelshift_r(w[0], w[1], sh);
}
template <bool ALIGNED4>
__device__ __forceinline__ uint64_t lds_u64(const uint8_t *p) {
    if (ALIGNED4) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
        return (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    }
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t lo = __funnelshift_r(w[0], w[1], sh);
    uint32_t hi = __funnelshift_r(w[1], w[2], sh);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

// ------------------------------------------------------------------ 2-bit k-mer arithmetic
// Reverse the order of the 32 two-bit groups of a 64-bit word.
__device__ __forceinline__ uint64_t rev2(uint64_t x) {
    x = __brevll(x);                                               // reverses single bits
    return ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);   // un-swap inside each pair
}

constexpr int kMaxWords = 8;   // k <= 256

// Reverse complement of a right-aligned k-mer held in w[0..s-1] (w[0] most significant).
template <int S>
__device__ __forceinline__ void revcomp_words(const uint64_t (&w)[S], uint64_t (&rc)[S], uint32_t k) {
    uint64_t t[S];
#pragma unroll
    for (int i = 0; i < S; ++i) t[i] = rev2(~w[S - 1 - i]);        // sequence now left-aligned in 64*S bits
    const uint32_t sh = 64u * S - 2u * k;                           // 0..62, even
    if (sh == 0) {
#pragma unroll
        for (int i = 0; i < S; ++i) rc[i] = t[i];
    } else {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            uint64_t v = t[i] >> sh;
            if (i > 0) v |= t[i - 1] << (64 - sh);
            rc[i] = v;
        }
    }
}

template <int S>
__device__ __forceinline__ bool words_less(const uint64_t (&a)[S], const uint64_t (&b)[S]) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
        if (a[i] != b[i]) return a[i] < b[i];
    }
    return false;
}
template <int S>
__device__ __forceinline__ bool words_equal(const uint64_t (&a)[S], const uint64_t (&b)[S]) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < S; ++i) e &= (a[i] == b[i]);
    return e;
}

// ASCII -> 2-bit code for ACGT / acgt: ((c>>1)&3) gives A0 C1 G3 T2; x ^ (x>>1) fixes G/T.
__device__ __forceinline__ uint32_t base_code(uint32_t c) {
    uint32_t x = (c >> 1) & 3u;
    return x ^ (x >> 1);
}
__device__ __forceinline__ bool is_upper_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
__device__ __forceinline__ bool is_lower_acgt(uint32_t c) { return c == 'a' || c == 'c' || c == 'g' || c == 't'; }

// SequenceUtils.complement (SequenceUtils.java:61-86): ACGT/acgt swap, everything else maps to itself.
__device__ __forceinline__ uint8_t complement_ascii(uint8_t b) {
    switch (b) {
        case 'A': return 'T'; case 'a': return 't';
        case 'C': return 'G'; case 'c': return 'g';
        case 'G': return 'C'; case 'g': return 'c';
        case 'T': return 'A'; case 't': return 'a';
        default: return b;
    }
}

}  // namespace cc
