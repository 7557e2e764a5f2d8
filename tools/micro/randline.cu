// Random 64-byte-line reads (the access pattern of the K4 line index): how many lines per second does HBM deliver,
// for per-thread lines (4 x LDG.128 per thread) vs quad-cooperative lines (one LDG.128 per lane, 8 lines per warp load)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randline randline.cu && ./randline
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
template <int LINE16>   // line size in 16-byte units: 2 = 32 B, 4 = 64 B, 8 = 128 B
__global__ void __launch_bounds__(512, 2) coop_kernel(const uint4 *lines, uint64_t nlines, uint64_t nq, uint32_t *out) {
    const uint32_t lane = threadIdx.x & 31, j = lane % LINE16;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * 512 + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * 512) {
        const uint64_t l = mix(i) % nlines;
        uint4 v[LINE16];
#pragma unroll
        for (int r = 0; r < LINE16; ++r) {
            const uint64_t lr = __shfl_sync(0xffffffffu, l, r, LINE16);
            v[r] = __ldg(lines + lr * LINE16 + j);
        }
#pragma unroll
        for (int r = 0; r < LINE16; ++r) acc += v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// the same with an explicit L2 prefetch-size qualifier on the loads (PF = 64, 128 or 256 bytes)
template <int PF>
__device__ __forceinline__ uint4 ld_pf(const uint4 *p) {
    uint4 v;
    if (PF == 64) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (PF == 128) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int PF>
__global__ void __launch_bounds__(512, 2) coop_pf_kernel(const uint4 *lines, uint64_t nlines, uint64_t nq, uint32_t *out) {
    const uint32_t lane = threadIdx.x & 31, j = lane % 4;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * 512 + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * 512) {
        const uint64_t l = mix(i) % nlines;
        uint4 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint64_t lr = __shfl_sync(0xffffffffu, l, r, 4);
            v[r] = ld_pf<PF>(lines + lr * 4 + j);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) acc += v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int LINE16>
__global__ void __launch_bounds__(512, 2) thread_kernel(const uint4 *lines, uint64_t nlines, uint64_t nq, uint32_t *out) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * 512 + threadIdx.x; i < nq; i += (uint64_t)gridDim.x * 512) {
        const uint64_t l = mix(i) % nlines;
        uint4 v[LINE16];
#pragma unroll
        for (int r = 0; r < LINE16; ++r) v[r] = __ldg(lines + l * LINE16 + r);
#pragma unroll
        for (int r = 0; r < LINE16; ++r) acc += v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <typename K>
void run(const char *name, K kern, int line16, const uint4 *buf, uint64_t bytes, uint64_t nq, uint32_t *out) {
    const uint64_t nlines = bytes / (16ull * line16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<148 * 2, 512>>>(buf, nlines, nq, out);
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) kern<<<148 * 2, 512>>>(buf, nlines, nq, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("%-22s line %3d B  table %.2f GB  %.3g lines/s  %.0f GB/s of lines\n", name, 16 * line16, bytes / 1e9, nq / ms * 1e3, nq / ms * 1e3 * 16 * line16 / 1e9);
}
int main() {
    const uint64_t bytes = 2560ull << 20, nq = 1ull << 28;
    uint4 *buf; uint32_t *out;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes); cudaMalloc(&out, 4);
    // does the L2 fetch-granularity hint change what a random line costs?
    for (size_t gran : {(size_t)128, (size_t)64, (size_t)32}) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity %zu -> %s, now %zu\n", gran, cudaGetErrorString(e), got);
        run("coop", coop_kernel<2>, 2, buf, bytes, nq, out);
        run("coop", coop_kernel<4>, 4, buf, bytes, nq, out);
    }
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128);
    run("coop", coop_kernel<2>, 2, buf, bytes, nq, out);
    run("coop", coop_kernel<4>, 4, buf, bytes, nq, out);
    run("coop", coop_kernel<8>, 8, buf, bytes, nq, out);
    run("thread", thread_kernel<2>, 2, buf, bytes, nq, out);
    run("thread", thread_kernel<4>, 4, buf, bytes, nq, out);
    run("thread", thread_kernel<8>, 8, buf, bytes, nq, out);
    run("coop L2::64B", coop_pf_kernel<64>, 4, buf, bytes, nq, out);
    run("coop L2::128B", coop_pf_kernel<128>, 4, buf, bytes, nq, out);
    run("coop L2::256B", coop_pf_kernel<256>, 4, buf, bytes, nq, out);
    run("coop 320MB (8-GPU shard)", coop_kernel<4>, 4, buf, 320ull << 20, nq, out);
    run("coop 40MB (L2)", coop_kernel<4>, 4, buf, 40ull << 20, nq, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
