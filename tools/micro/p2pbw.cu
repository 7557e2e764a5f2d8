// Peer-to-peer bandwidth over NVLink from inside a kernel (2+ GPUs): what can the route leg (stores into a peer's inbox) and
// the gather leg (loads from a peer's results) reach, with plain vector accesses and with the bulk-copy (TMA) engine?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2pbw p2pbw.cu && ./p2pbw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) store_plain(uint4 *dst, uint64_t n16) {
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * 256) dst[i] = v;
}
__global__ void __launch_bounds__(256) load_plain(const uint4 *src, uint64_t n16, uint32_t *sink) {
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * 256 * 4) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t at = i + (uint64_t)j * gridDim.x * 256;
            v[j] = at < n16 ? __ldcv(src + at) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j].x ^ v[j].w;
    }
    if (acc == 0x1234567u) *sink = acc;
}
// every CTA owns a shared-memory buffer of CHUNK bytes and pushes it to consecutive CHUNK-sized pieces of dst with
// cp.async.bulk (shared -> global), keeping DEPTH copies in flight
template <int CHUNK, int DEPTH>
__global__ void __launch_bounds__(128) store_bulk(uint8_t *dst, uint64_t bytes) {
    extern __shared__ __align__(128) uint8_t buf[];
    for (int i = threadIdx.x; i < CHUNK / 4; i += 128) reinterpret_cast<uint32_t *>(buf)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t nchunks = bytes / CHUNK;
        int inflight = 0;
        for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CHUNK), "r"(smem_u32(buf)), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= DEPTH) { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory"); --inflight; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
template <typename F>
float timed(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}
int main() {
    int nd = 0; cudaGetDeviceCount(&nd);
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const uint64_t bytes = 2ull << 30;
    uint8_t *remote, *local; uint32_t *sink;
    cudaSetDevice(1); cudaMalloc(&remote, bytes); cudaMemset(remote, 1, bytes);
    cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0); cudaMalloc(&local, bytes); cudaMalloc(&sink, 4);
    for (int pass = 0; pass < 2; ++pass) {
        uint8_t *dst = pass ? remote : local;
        const char *where = pass ? "peer (NVLink)" : "local HBM";
        for (int per_sm : {1, 2, 4, 8}) {
            float ms = timed([&] { store_plain<<<148 * per_sm, 256>>>((uint4 *)dst, bytes / 16); });
            printf("%-14s store plain 16 B/lane  %d CTAs/SM  %.0f GB/s\n", where, per_sm, bytes / ms / 1e6);
        }
        for (int per_sm : {1, 4, 8}) {
            float ms = timed([&] { load_plain<<<148 * per_sm, 256>>>((const uint4 *)dst, bytes / 16, sink); });
            printf("%-14s load  plain 16 B/lane  %d CTAs/SM  %.0f GB/s\n", where, per_sm, bytes / ms / 1e6);
        }
        for (int per_sm : {1, 2, 4}) {
            float ms = timed([&] { store_bulk<2048, 4><<<148 * per_sm, 128, 2048>>>(dst, bytes); });
            printf("%-14s store bulk 2 KB x4     %d CTAs/SM  %.0f GB/s\n", where, per_sm, bytes / ms / 1e6);
            ms = timed([&] { store_bulk<8192, 4><<<148 * per_sm, 128, 8192>>>(dst, bytes); });
            printf("%-14s store bulk 8 KB x4     %d CTAs/SM  %.0f GB/s\n", where, per_sm, bytes / ms / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
