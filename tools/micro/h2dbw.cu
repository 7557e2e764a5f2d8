// Host -> device copy bandwidth with 1, 2, 4, 8 GPUs copying at the same time from their own pinned buffers (one process):
// the ceiling of the box for the host-streamed scan (bench.py e2e), independent of this library.
//   nvcc -O3 -o h2dbw h2dbw.cu && ./h2dbw
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
int main() {
    int nd = 0; cudaGetDeviceCount(&nd);
    if (nd > 8) nd = 8;
    const size_t bytes = 900000000;
    std::vector<void *> host(nd), dev(nd);
    std::vector<cudaStream_t> st(nd);
    std::vector<cudaEvent_t> e0(nd), e1(nd);
    for (int d = 0; d < nd; ++d) {
        cudaSetDevice(d);
        cudaHostAlloc(&host[d], bytes, cudaHostAllocDefault);
        cudaMalloc(&dev[d], bytes);
        cudaStreamCreate(&st[d]); cudaEventCreate(&e0[d]); cudaEventCreate(&e1[d]);
    }
    for (int active = 1; active <= nd; active *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            for (int d = 0; d < active; ++d) {
                cudaSetDevice(d);
                cudaEventRecord(e0[d], st[d]);
                for (int i = 0; i < 5; ++i) cudaMemcpyAsync(dev[d], host[d], bytes, cudaMemcpyHostToDevice, st[d]);
                cudaEventRecord(e1[d], st[d]);
            }
            for (int d = 0; d < active; ++d) { cudaSetDevice(d); cudaStreamSynchronize(st[d]); }
        }
        float worst = 0, best = 1e9;
        for (int d = 0; d < active; ++d) { float ms; cudaEventElapsedTime(&ms, e0[d], e1[d]); worst = ms > worst ? ms : worst; best = ms < best ? ms : best; }
        printf("%d GPUs copying: %.1f GB/s per GPU (slowest), %.1f (fastest), %.0f GB/s in total\n", active, 5.0 * bytes / worst / 1e6, 5.0 * bytes / best / 1e6,
               active * 5.0 * bytes / worst / 1e6);
    }
    return 0;
}
