// Read-only HBM bandwidth calibration on B200: (a) grid-stride 16-byte loads, (b) 1-D bulk (TMA) copies into a
// shared-memory ring with nothing consuming the data.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw membw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../corticall_b200/csrc/device_utils.cuh"
using namespace cc;

__global__ void ldg_read(const uint4 *p, size_t n, unsigned long long *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll 8
    for (; i < n; i += stride) {
        uint4 v = __ldcs(p + i);
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) atomicAdd(sink, 1ull);
}

// each CTA: one thread issues bulk copies of `tile` bytes round-robin over `stages` buffers; static tile assignment
__global__ void tma_read(const uint8_t *p, size_t bytes, uint32_t tile, uint32_t stages, int use_hint) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *buf = smem + 1024;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < stages; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const uint64_t pol = make_evict_first_policy();
    const size_t ntiles = bytes / tile;
    uint32_t it = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t st = it % stages, ph = (it / stages) & 1u;
        if (it >= stages) mbar_wait(&bars[st], ph ^ 1u, nullptr, 0);      // previous copy into this buffer has landed
        mbar_arrive_expect_tx(&bars[st], tile);
        if (use_hint) bulk_g2s(buf + (size_t)st * tile, p + t * tile, tile, &bars[st], pol);
        else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                          ::"r"(smem_u32(buf + (size_t)st * tile)), "l"(p + t * tile), "r"(tile), "r"(smem_u32(&bars[st])) : "memory");
    }
    // drain
    for (uint32_t s = 0; s < stages && s < it; ++s) {
        const uint32_t j = it - 1 - s, st = j % stages, ph = (j / stages) & 1u;
        mbar_wait(&bars[st], ph, nullptr, 0);
    }
}

int main() {
    const size_t bytes = 900ull * 1000 * 1000 / 16 * 16;
    uint8_t *d; unsigned long long *sink;
    cudaMalloc(&d, bytes + 65536); cudaMalloc(&sink, 8);
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int bpsm : {2, 4, 8, 16, 32}) for (int threads : {256, 512}) {
        int grid = 148 * bpsm;
        for (int i = 0; i < 3; ++i) ldg_read<<<grid, threads>>>((const uint4 *)d, bytes / 16, sink);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) ldg_read<<<grid, threads>>>((const uint4 *)d, bytes / 16, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("ldg   grid=148x%-2d threads=%d  %.3f ms  %.0f GB/s\n", bpsm, threads, ms / 20, bytes / (ms / 20) / 1e6);
    }
    cudaFuncSetAttribute(tma_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    for (int hint : {1, 0}) for (uint32_t tile : {8192u, 16384u, 32768u, 65536u}) for (uint32_t stages : {2u, 3u, 4u, 6u}) for (int ctas : {1, 2, 4}) {
        size_t smem = 1024 + (size_t)tile * stages;
        if (smem * ctas > 226 * 1024) continue;
        int grid = 148 * ctas;
        for (int i = 0; i < 3; ++i) tma_read<<<grid, 32, smem>>>(d, bytes, tile, stages, hint);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) tma_read<<<grid, 32, smem>>>(d, bytes, tile, stages, hint);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        printf("tma   hint=%d tile=%-6u stages=%u ctas/sm=%d  %.3f ms  %.0f GB/s %s\n", hint, tile, stages, ctas, ms / 20, bytes / (ms / 20) / 1e6,
               err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
