// prefilter.cu -- scan-shaped pre-filters and recovery around the novelty step.
//
// Reference semantics reproduced (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
//   FindLowCoverage        S/commands/prefilter/FindLowCoverage.java:33-66    records with coverage[0] < minCoverage (Java int compare)
//   FindShared             S/commands/prefilter/FindShared.java:40-118        ROI records whose k-mer has coverage > 0 in a colour of the
//                                                                              pedigree graph that is neither child, parent nor ignored
//   RecoverExcludedKmers   S/commands/discover/recover/RecoverExcludedKmers.java:31-106
//   CovStats               S/commands/utils/CovStats.java:33-72               histogram of child coverage over records shared with parents
//                                                                              and with other samples
// All coverage comparisons are on Java ints (uint32 on disk reinterpreted as signed, BinaryUtils.java:6-17).
//
// B200 design.  These are column problems: the record array is streamed ONCE by decode_columns_kernel (the TMA ring of scan.cu)
// into a coalesced int32 coverage matrix (and the key column where lookups follow); the predicates below then run as
// fully coalesced column kernels, the selection is an index list (cub::DeviceSelect, plumbing), and one gather kernel
// writes the projected records.  The lookups inside FindShared / RecoverExcludedKmers are the K4 kernels of lookup.cu.
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "cc_internal.hpp"
#include "device_utils.cuh"

namespace cc {

namespace {

constexpr int kPBlock = 256;

__device__ __forceinline__ int32_t read_cov(const uint8_t *rec, uint32_t s, uint32_t color) {
    const uint8_t *p = rec + 8u * s + 4u * color;              // records are byte-aligned only
    return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}
__device__ __forceinline__ bool in_mask(const uint32_t *mask, uint32_t c) { return (mask[c >> 5] >> (c & 31u)) & 1u; }

// FindLowCoverage.java:48-57: `if (cr.getCoverage(0) >= MIN_COVERAGE) kept else written`
__global__ void lowcov_flags_kernel(const int32_t *__restrict__ cov, uint64_t n, uint32_t c, int32_t min_cov, uint8_t *__restrict__ flags) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        flags[i] = cov[i * c] < min_cov ? 1 : 0;
}

// Remove.java:47-55: a merged record is dropped when a colour of a secondary graph (colours c_primary .. c-1 of the
// collection) has coverage > 0 (signed); flags = 1 for the records that are KEPT.
__global__ void remove_flags_kernel(const int32_t *__restrict__ cov, uint64_t n, uint32_t c, uint32_t c_primary, uint8_t *__restrict__ flags) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t *row = cov + i * c;
        bool found = false;
        for (uint32_t cc = c_primary; cc < c; ++cc) found |= row[cc] > 0;
        flags[i] = found ? 0 : 1;
    }
}

// RecoverExcludedKmers.java:50-62: class 1 = child coverage > 0 (written as is); class 2 = candidate (child <= 0 and
// another colour > 0: looked up in the dirty graph); 0 = dropped.  find_flags: 0 for candidates, 2 (skip) otherwise.
__global__ void recover_classes_kernel(const int32_t *__restrict__ cov, uint64_t n, uint32_t c, uint32_t child, uint8_t *__restrict__ cls,
                                       uint8_t *__restrict__ find_flags) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t *row = cov + i * c;
        uint8_t k = 0;
        if (row[child] > 0) k = 1;
        else {
            for (uint32_t cc = 0; cc < c; ++cc)
                if (cc != child && row[cc] > 0) { k = 2; break; }
        }
        cls[i] = k;
        find_flags[i] = k == 2 ? 0 : 2;
    }
}

// RecoverExcludedKmers.java:63-86: a candidate is recovered when the dirty graph holds its k-mer with coverage(0) > 0; its
// child coverage becomes the dirty one.  flags: 0 dropped, 1 written as is, 3 written with patch[i].
__global__ void recover_finalize_kernel(const uint8_t *__restrict__ cls, const int64_t *__restrict__ idx, const uint8_t *__restrict__ dirty_body,
                                        uint32_t dirty_S, uint32_t s, uint64_t dirty_first, uint64_t n, uint8_t *__restrict__ flags,
                                        int32_t *__restrict__ patch, unsigned long long *recovered) {
    unsigned long long mine = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t f = cls[i] == 1 ? 1 : 0;
        if (cls[i] == 2 && idx[i] >= 0) {
            const int32_t dc = read_cov(dirty_body + (uint64_t)(idx[i] - (int64_t)dirty_first) * dirty_S, s, 0);
            if (dc > 0) { f = 3; patch[i] = dc; ++mine; }
        }
        flags[i] = f;
    }
    if (mine) atomicAdd(recovered, mine);
}

// FindShared.java:64-77: shared when a colour outside {child, parents, ignored} has coverage > 0 in the pedigree graph's
// record.  A ROI k-mer absent from the graph is a NullPointerException in the reference: reported through missing_at.
__global__ void shared_flags_kernel(const int64_t *__restrict__ idx, const uint8_t *__restrict__ graph_body, uint32_t S, uint32_t s, uint32_t c,
                                    uint64_t graph_first, const uint32_t *__restrict__ excluded, uint64_t nroi, uint8_t *__restrict__ flags,
                                    unsigned long long *missing_at) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nroi; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t f = 0;
        if (idx[i] < 0) atomicMin(missing_at, (unsigned long long)i);
        else {
            const uint8_t *rec = graph_body + (uint64_t)(idx[i] - (int64_t)graph_first) * S;
            for (uint32_t cc = 0; cc < c; ++cc)
                if (!in_mask(excluded, cc) && read_cov(rec, s, cc) > 0) { f = 1; break; }
        }
        flags[i] = f;
    }
}

// CovStats.java:46-66: for records present in the child, in >= 1 parent and in >= 1 other sample:
// hist[childCov] += numberOfParents + numberOfChildren.  key 0 = record does not count (childCov > 0 whenever it does).
__global__ void covstats_pairs_kernel(const int32_t *__restrict__ cov, uint64_t n, uint32_t c, int32_t child, const uint32_t *__restrict__ parents,
                                      uint32_t *__restrict__ key, long long *__restrict__ weight) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t *row = cov + i * c;
        const bool in_child = child >= 0 && (uint32_t)child < c && row[child] > 0;
        uint32_t np = 0, nc = 0;
        for (uint32_t cc = 0; cc < c; ++cc) {
            if (row[cc] > 0 && (int32_t)cc != child) {
                if (in_mask(parents, cc)) ++np; else ++nc;
            }
        }
        const bool counts = in_child && np > 0 && nc > 0;
        key[i] = counts ? (uint32_t)row[child] : 0u;
        weight[i] = counts ? (long long)(np + nc) : 0ll;
    }
}

// CovStats as ONE pass over the record array: no coverage matrix, no sort.  Child coverages below kCovHistSmem go to a
// per-CTA shared-memory histogram (flushed once), below kCovHistGlobal to a global table of 64-bit counters, anything
// larger (never seen in a real graph; possible in a file) to a short overflow list the host aggregates.  Weights are
// summed exactly; the shared counters are 32-bit and the launcher bounds what one CTA can add.
constexpr uint32_t kCovHistSmem = 4096, kCovHistGlobal = 65536, kCovOverflowCap = 1u << 20;

__global__ void __launch_bounds__(kPBlock) covstats_hist_kernel(const uint8_t *__restrict__ body, uint64_t n, uint32_t s, uint32_t c, uint32_t child,
                                                                const uint32_t *__restrict__ parents, unsigned long long *__restrict__ hist_global,
                                                                uint2 *__restrict__ overflow, unsigned int *__restrict__ overflow_n) {
    __shared__ uint32_t hist[kCovHistSmem];
    for (uint32_t b = threadIdx.x; b < kCovHistSmem; b += kPBlock) hist[b] = 0;
    __syncthreads();
    const uint32_t S = 8u * s + 5u * c;
    for (uint64_t i = (uint64_t)blockIdx.x * kPBlock + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kPBlock) {
        const uint8_t *rec = body + i * S;
        const int32_t cc_child = read_cov(rec, s, child);
        if (cc_child <= 0) continue;
        uint32_t np = 0, nc = 0;
        for (uint32_t cc = 0; cc < c; ++cc) {
            if (cc != child && read_cov(rec, s, cc) > 0) {
                if (in_mask(parents, cc)) ++np; else ++nc;
            }
        }
        if (np == 0 || nc == 0) continue;
        const uint32_t w = np + nc, v = (uint32_t)cc_child;
        if (v < kCovHistSmem) atomicAdd(&hist[v], w);
        else if (v < kCovHistGlobal) atomicAdd(&hist_global[v], (unsigned long long)w);
        else {
            const unsigned int slot = atomicAdd(overflow_n, 1u);
            if (slot < kCovOverflowCap) overflow[slot] = make_uint2(v, w);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < kCovHistSmem; b += kPBlock)
        if (hist[b]) atomicAdd(&hist_global[b], (unsigned long long)hist[b]);
}

// Output record j = the first c_out colours of input record sel[j] (CortexGraphWriter.addRecord :106-138 writes
// header.getNumColors() coverages and edges of whatever record it is given), coverage[patch_color] replaced where flags == 3.
// A warp writes 32 consecutive output records as one contiguous byte stream.
__global__ void project_records_kernel(const uint8_t *__restrict__ body, uint32_t s, uint32_t c_in, const uint32_t *__restrict__ sel, uint64_t m,
                                       uint32_t c_out, const uint8_t *__restrict__ flags, const int32_t *__restrict__ patch, uint32_t patch_color,
                                       uint8_t *__restrict__ out) {
    const uint32_t Si = 8u * s + 5u * c_in, So = 8u * s + 5u * c_out;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = warp; g * 32 < m; g += nwarps) {
        const uint64_t j0 = g * 32;
        const uint32_t rows = (uint32_t)min((uint64_t)32, m - j0);
        const uint64_t mine = lane < rows ? sel[j0 + lane] : 0;
        const bool patched = lane < rows && flags && flags[mine] == 3 && patch_color < c_out;
        const int32_t pv = patched ? patch[mine] : 0;
        for (uint32_t b0 = 0; b0 < rows * So; b0 += 32) {       // warp-uniform trip count: every lane takes part in the shuffles
            const uint32_t b = b0 + lane;
            const bool active = b < rows * So;
            const uint32_t r = active ? b / So : 0u, off = b - r * So;
            const uint64_t src_rec = __shfl_sync(0xffffffffu, mine, r);
            const bool src_patched = __shfl_sync(0xffffffffu, patched ? 1 : 0, r) != 0;
            const int32_t src_pv = __shfl_sync(0xffffffffu, pv, r);
            uint32_t in_off = off;                                           // words and the first c_out coverages sit at the same offsets
            if (off >= 8u * s + 4u * c_out) in_off = 8u * s + 4u * c_in + (off - 8u * s - 4u * c_out);   // edges
            if (!active) continue;
            uint8_t v = body[src_rec * Si + in_off];
            const uint32_t p0 = 8u * s + 4u * patch_color;
            if (src_patched && off >= p0 && off < p0 + 4u) v = (uint8_t)((uint32_t)src_pv >> (8u * (off - p0)));
            out[j0 * So + b] = v;
        }
    }
}

int grid_of(uint64_t n, int sm_count) {
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + kPBlock - 1) / kPBlock, (uint64_t)sm_count * 8));
}

}  // namespace

int launch_lowcov_flags(const int32_t *cov, uint64_t n, uint32_t c, int32_t min_cov, uint8_t *flags, int sm_count, cudaStream_t st) {
    if (n == 0) return CC_OK;
    lowcov_flags_kernel<<<grid_of(n, sm_count), kPBlock, 0, st>>>(cov, n, c, min_cov, flags);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_remove_flags(const int32_t *cov, uint64_t n, uint32_t c, uint32_t c_primary, uint8_t *flags, int sm_count, cudaStream_t st) {
    if (n == 0) return CC_OK;
    remove_flags_kernel<<<grid_of(n, sm_count), kPBlock, 0, st>>>(cov, n, c, c_primary, flags);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_recover_classes(const int32_t *cov, uint64_t n, uint32_t c, uint32_t child, uint8_t *cls, uint8_t *find_flags, int sm_count,
                           cudaStream_t st) {
    if (n == 0) return CC_OK;
    recover_classes_kernel<<<grid_of(n, sm_count), kPBlock, 0, st>>>(cov, n, c, child, cls, find_flags);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_recover_finalize(const uint8_t *cls, const int64_t *idx, const uint8_t *dirty_body, uint32_t dirty_S, uint32_t s, uint64_t dirty_first,
                            uint64_t n, uint8_t *flags, int32_t *patch, unsigned long long *recovered, int sm_count, cudaStream_t st) {
    if (n == 0) return CC_OK;
    recover_finalize_kernel<<<grid_of(n, sm_count), kPBlock, 0, st>>>(cls, idx, dirty_body, dirty_S, s, dirty_first, n, flags, patch, recovered);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

int launch_shared_flags(const int64_t *idx, const uint8_t *graph_body, uint32_t S, uint32_t s, uint32_t c, uint64_t graph_first,
                        const uint32_t *excluded, uint64_t nroi, uint8_t *flags, unsigned long long *missing_at, int sm_count, cudaStream_t st) {
    if (nroi == 0) return CC_OK;
    shared_flags_kernel<<<grid_of(nroi, sm_count), kPBlock, 0, st>>>(idx, graph_body, S, s, c, graph_first, excluded, nroi, flags, missing_at);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// Indices of the records with flags != 0, ascending (= input order).  *sel_out is stream-ordered memory (cudaFreeAsync).
int select_flagged(const uint8_t *flags, uint64_t n, uint32_t **sel_out, uint64_t *m_out, cudaStream_t st) {
    *sel_out = nullptr;
    *m_out = 0;
    if (n >= (1ull << 32)) return fail(CC_ERR_UNSUPPORTED, "selection is limited to 2^32-1 records");
    uint32_t *sel = nullptr;
    unsigned long long *d_m = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    CC_CUDA(cudaMallocAsync(&sel, std::max<uint64_t>(n, 1) * 4, st));
    CC_CUDA(cudaMallocAsync(&d_m, 8, st));
    cub::CountingInputIterator<uint32_t> iota(0);
    CC_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, flags, sel, d_m, (int64_t)n, st));
    CC_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 16), st));
    cudaError_t e = cub::DeviceSelect::Flagged(tmp, tmp_bytes, iota, flags, sel, d_m, (int64_t)n, st);
    count_launch(2);
    unsigned long long m = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&m, d_m, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(tmp, st);
    cudaFreeAsync(d_m, st);
    if (e != cudaSuccess) { cudaFreeAsync(sel, st); return cuda_fail(e, "cub::DeviceSelect::Flagged", __FILE__, __LINE__); }
    *sel_out = sel;
    *m_out = m;
    return CC_OK;
}

int launch_project_records(const uint8_t *body, uint32_t s, uint32_t c_in, const uint32_t *sel, uint64_t m, uint32_t c_out,
                           const uint8_t *flags, const int32_t *patch, uint32_t patch_color, uint8_t *out, int sm_count, cudaStream_t st) {
    if (m == 0) return CC_OK;
    const uint64_t groups = (m + 31) / 32;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + kPBlock / 32 - 1) / (kPBlock / 32), (uint64_t)sm_count * 8));
    project_records_kernel<<<grid, kPBlock, 0, st>>>(body, s, c_in, sel, m, c_out, flags, patch, patch_color, out);
    count_launch();
    CC_CUDA(cudaGetLastError());
    return CC_OK;
}

// CovStats on the device: (childCov, weight) per record, radix sort by childCov, reduce by key.  Returns the host table
// (ascending coverage, key 0 dropped); counts are summed in 64 bits (the caller wraps them to Java ints).
int cov_stats(const int32_t *cov, uint64_t n, uint32_t c, int32_t child, const uint32_t *dev_parent_mask, int sm_count, cudaStream_t st,
              std::vector<int32_t> &out_cov, std::vector<long long> &out_count) {
    out_cov.clear();
    out_count.clear();
    if (n == 0) return CC_OK;
    if (n >= (1ull << 31)) return fail(CC_ERR_UNSUPPORTED, "CovStats is limited to 2^31-1 records per device");
    uint32_t *key_a = nullptr, *key_b = nullptr, *ukeys = nullptr;
    long long *w_a = nullptr, *w_b = nullptr, *sums = nullptr;
    unsigned long long *d_runs = nullptr;
    void *tmp = nullptr;
    struct Free {
        cudaStream_t st; uint32_t *&a, *&b, *&u; long long *&c, *&d, *&s; unsigned long long *&r; void *&t;
        ~Free() { cudaFreeAsync(a, st); cudaFreeAsync(b, st); cudaFreeAsync(u, st); cudaFreeAsync(c, st); cudaFreeAsync(d, st);
                  cudaFreeAsync(s, st); cudaFreeAsync(r, st); cudaFreeAsync(t, st); }
    } fr{st, key_a, key_b, ukeys, w_a, w_b, sums, d_runs, tmp};
    CC_CUDA(cudaMallocAsync(&key_a, n * 4, st)); CC_CUDA(cudaMallocAsync(&key_b, n * 4, st)); CC_CUDA(cudaMallocAsync(&ukeys, n * 4, st));
    CC_CUDA(cudaMallocAsync(&w_a, n * 8, st)); CC_CUDA(cudaMallocAsync(&w_b, n * 8, st)); CC_CUDA(cudaMallocAsync(&sums, n * 8, st));
    CC_CUDA(cudaMallocAsync(&d_runs, 8, st));
    covstats_pairs_kernel<<<grid_of(n, sm_count), kPBlock, 0, st>>>(cov, n, c, child, dev_parent_mask, key_a, w_a);
    count_launch();
    size_t b1 = 0, b2 = 0;
    CC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, key_a, key_b, w_a, w_b, (int64_t)n, 0, 32, st));
    CC_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, b2, key_b, ukeys, w_b, sums, d_runs, cub::Sum(), (int64_t)n, st));
    CC_CUDA(cudaMallocAsync(&tmp, std::max(b1, b2) + 16, st));
    size_t tb = std::max(b1, b2) + 16;
    CC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key_a, key_b, w_a, w_b, (int64_t)n, 0, 32, st));
    tb = std::max(b1, b2) + 16;
    CC_CUDA(cub::DeviceReduce::ReduceByKey(tmp, tb, key_b, ukeys, w_b, sums, d_runs, cub::Sum(), (int64_t)n, st));
    count_launch(10);
    unsigned long long runs = 0;
    CC_CUDA(cudaMemcpyAsync(&runs, d_runs, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> hk(runs);
    std::vector<long long> hs(runs);
    if (runs) {
        CC_CUDA(cudaMemcpyAsync(hk.data(), ukeys, runs * 4, cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaMemcpyAsync(hs.data(), sums, runs * 8, cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
    }
    for (uint64_t i = 0; i < runs; ++i) {
        if (hk[i] == 0) continue;
        out_cov.push_back((int32_t)hk[i]);
        out_count.push_back(hs[i]);
    }
    return CC_OK;
}

// CovStats straight from the record array (covstats_hist_kernel).  *fell_back = true when the overflow list did not hold
// every outlier (or the 32-bit shared counters could wrap): the caller then takes the sort-based path.
int cov_stats_fused(const uint8_t *body, uint64_t n, uint32_t s, uint32_t c, int32_t child, const uint32_t *dev_parent_mask, int sm_count,
                    cudaStream_t st, std::vector<int32_t> &out_cov, std::vector<long long> &out_count, bool *fell_back) {
    out_cov.clear();
    out_count.clear();
    *fell_back = false;
    if (n == 0) return CC_OK;
    const int grid = grid_of(n, sm_count);
    if (((n + (uint64_t)grid * kPBlock - 1) / ((uint64_t)grid * kPBlock)) * kPBlock * (uint64_t)c >= (1ull << 32)) { *fell_back = true; return CC_OK; }
    unsigned long long *hist = nullptr;
    uint2 *overflow = nullptr;
    unsigned int *d_n = nullptr;
    struct Free {
        cudaStream_t st; unsigned long long *&h; uint2 *&o; unsigned int *&n;
        ~Free() { cudaFreeAsync(h, st); cudaFreeAsync(o, st); cudaFreeAsync(n, st); }
    } fr{st, hist, overflow, d_n};
    CC_CUDA(cudaMallocAsync(&hist, kCovHistGlobal * sizeof(unsigned long long), st));
    CC_CUDA(cudaMallocAsync(&overflow, (size_t)kCovOverflowCap * sizeof(uint2), st));
    CC_CUDA(cudaMallocAsync(&d_n, sizeof(unsigned int), st));
    CC_CUDA(cudaMemsetAsync(hist, 0, kCovHistGlobal * sizeof(unsigned long long), st));
    CC_CUDA(cudaMemsetAsync(d_n, 0, sizeof(unsigned int), st));
    covstats_hist_kernel<<<grid, kPBlock, 0, st>>>(body, n, s, c, (uint32_t)child, dev_parent_mask, hist, overflow, d_n);
    count_launch();
    CC_CUDA(cudaGetLastError());
    std::vector<unsigned long long> h(kCovHistGlobal);
    unsigned int nover = 0;
    CC_CUDA(cudaMemcpyAsync(h.data(), hist, kCovHistGlobal * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaMemcpyAsync(&nover, d_n, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    if (nover > kCovOverflowCap) { *fell_back = true; return CC_OK; }
    for (uint32_t v = 1; v < kCovHistGlobal; ++v) {
        if (h[v]) { out_cov.push_back((int32_t)v); out_count.push_back((long long)h[v]); }
    }
    if (nover) {
        std::vector<uint2> ov(nover);
        CC_CUDA(cudaMemcpyAsync(ov.data(), overflow, (size_t)nover * sizeof(uint2), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
        std::sort(ov.begin(), ov.end(), [](const uint2 &a, const uint2 &b) { return a.x < b.x; });
        for (size_t i = 0; i < ov.size();) {
            size_t j = i;
            long long sum = 0;
            for (; j < ov.size() && ov[j].x == ov[i].x; ++j) sum += ov[j].y;
            out_cov.push_back((int32_t)ov[i].x);
            out_count.push_back(sum);
            i = j;
        }
    }
    return CC_OK;
}

}  // namespace cc
