// All-to-all store bandwidth from inside kernels on N GPUs of one process (the traffic pattern of the route leg): every GPU
// stores `bytes` to EACH of its N-1 peers at the same time, in chunks of `chunk` bytes that rotate over the peers, with plain
// 16-byte stores by a whole CTA or with one bulk copy (TMA engine) per chunk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o a2abw a2abw.cu && ./a2abw
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

struct Peers { uint8_t *p[8]; };
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) a2a_plain(Peers dst, int npeers, int me, uint64_t bytes, uint32_t chunk) {
    const uint64_t per_peer = bytes / chunk, total = per_peer * npeers;
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, me, 7);
    for (uint64_t j = blockIdx.x; j < total; j += gridDim.x) {
        uint4 *d = reinterpret_cast<uint4 *>(dst.p[j % npeers] + (uint64_t)me * bytes + (j / npeers) * chunk);
        for (uint32_t i = threadIdx.x; i < chunk / 16; i += 128) d[i] = v;
    }
}
template <int DEPTH>
__global__ void __launch_bounds__(128) a2a_bulk(Peers dst, int npeers, int me, uint64_t bytes, uint32_t chunk) {
    extern __shared__ __align__(128) uint8_t buf[];
    for (uint32_t i = threadIdx.x; i < chunk / 4; i += 128) reinterpret_cast<uint32_t *>(buf)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t per_peer = bytes / chunk, total = per_peer * npeers;
        int inflight = 0;
        for (uint64_t j = blockIdx.x; j < total; j += gridDim.x) {
            uint8_t *d = dst.p[j % npeers] + (uint64_t)me * bytes + (j / npeers) * chunk;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(smem_u32(buf)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= DEPTH) { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory"); --inflight; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    int nd = 0; cudaGetDeviceCount(&nd);
    if (nd < 2) { printf("needs 2+ GPUs\n"); return 0; }
    if (nd > 8) nd = 8;
    const uint64_t bytes = 192ull << 20;                     // per (source, destination) pair
    std::vector<uint8_t *> recv(nd);
    std::vector<cudaStream_t> st(nd);
    std::vector<cudaEvent_t> e0(nd), e1(nd);
    for (int d = 0; d < nd; ++d) {
        cudaSetDevice(d);
        for (int o = 0; o < nd; ++o) if (o != d) cudaDeviceEnablePeerAccess(o, 0);
        cudaMalloc(&recv[d], bytes * nd);
        cudaStreamCreate(&st[d]); cudaEventCreate(&e0[d]); cudaEventCreate(&e1[d]);
    }
    cudaGetLastError();
    auto run = [&](const char *name, int mode, uint32_t chunk, int per_sm) {
        for (int rep = 0; rep < 2; ++rep) {
            for (int d = 0; d < nd; ++d) {
                cudaSetDevice(d);
                Peers p{}; int np = 0;
                for (int o = 0; o < nd; ++o) if (o != d) p.p[np++] = recv[o];
                cudaEventRecord(e0[d], st[d]);
                if (mode == 0) a2a_plain<<<148 * per_sm, 128, 0, st[d]>>>(p, np, d, bytes, chunk);
                else a2a_bulk<4><<<148 * per_sm, 128, chunk, st[d]>>>(p, np, d, bytes, chunk);
                cudaEventRecord(e1[d], st[d]);
            }
            for (int d = 0; d < nd; ++d) { cudaSetDevice(d); cudaStreamSynchronize(st[d]); }
        }
        float worst = 0, best = 1e9;
        for (int d = 0; d < nd; ++d) { float ms; cudaEventElapsedTime(&ms, e0[d], e1[d]); worst = ms > worst ? ms : worst; best = ms < best ? ms : best; }
        printf("%-6s chunk %6u B  %d CTAs/SM  %d GPUs: %.0f GB/s per GPU out (slowest), %.0f (fastest)\n", name, chunk, per_sm, nd,
               bytes * (nd - 1) / worst / 1e6, bytes * (nd - 1) / best / 1e6);
    };
    for (uint32_t chunk : {512u, 1536u, 4096u, 16384u}) {
        run("plain", 0, chunk, 4);
        run("plain", 0, chunk, 8);
        if (chunk % 16 == 0) { run("bulk", 1, chunk, 2); run("bulk", 1, chunk, 4); }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
