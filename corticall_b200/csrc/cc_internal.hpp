// cc_internal.hpp -- host-side internals of libcorticall_cuda (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/corticall_cuda.h"

namespace cc {

// ------------------------------------------------------------------ errors
void set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
int fail(int status, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CC_CUDA(expr)                                                        \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) return ::cc::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint32_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------ options (cc_set_option)
struct Options {
    int scan_stages = 3;
    int scan_tile_bytes = 32768;
    int scan_ctas_per_sm = 2;
    int index_bits = 0;          // log2 of the number of bins of the lookup index (0 = auto: ~256 keys per bin, at most 2^13)
    int index_fill_pct = 50;     // average fill of a bucket line, in percent of its capacity
    int rows_rpt2_max_k = 32;    // cc_pack_kmers: two rows per thread (512-row tiles) up to this k, else one
    int lookup_l2_hints = -1;    // line loads: -1 = auto (plain up to 1 GB of lines, evict-first beyond), 0 plain, 1 evict-first, 2 evict-normal, 3 evict-last
    int find_bins_smem = 1;      // packed / routed search: stage the bin table in shared memory
    int rows_warp = 2;           // cc_pack_kmers on independent rows: 2 / 3 = warp-autonomous kernel with that many raw buffers per warp, 0 = CTA tiles
    int join_tiled = 1;          // Join / Remove: 1 = tiled union through shared memory, 0 = one global merge-path search per thread (round-1 form)
    int join_tile_kb = 48;       // tiled union: shared-memory budget per CTA in KB for the key slices (decides the tile size: 2048 elements at k <= 64)
    int covstats_fused = 1;      // CovStats: 1 = one histogram pass over the records, 0 = coverage matrix + radix sort + reduce by key (the fallback)
    int rows_fused = 1;          // ASCII query lists: 1 = pack + search in one kernel, 0 = pack, then search
    int route_stage_depth = 2;            // staging areas per route CTA (2..4): tiles whose bulk copies may still be reading shared memory
    int route_blocks_per_sm = 0;          // 0 = as many as fit; the overlapped pipeline uses 1
    int routed_search_blocks_per_sm = 0;  // cap on the CTAs of find_routed_kernel per SM (0 = all resident CTAs)
    int gather_blocks_per_sm = 16;  // cap on resident gather CTAs per SM
    int host_chunk_mb = 64;      // cc_find_novel_host chunk size
    int scan_fast = 1;           // 1 = chunked deferred-look-back kernel first, general kernel only on overflow
    int scan_chunk_tiles = 16;   // tiles per chunk of the fast kernel (power of two, <= 16)
    int scan_stage_buf_bytes = 4096;   // fast kernel: staging bytes (global scratch) per consumer warp per chunk parity
    int scan_pdl = 1;            // launch the rewrite kernel with programmatic stream serialisation (overlaps its launch with the scan's tail)
    int scan_debug = 0;          // diagnosis only: bit0 = skip the look-back (WRONG output positions)
};
Options &options();

// ------------------------------------------------------------------ header model (ctx_spec.md tables 1-3)
struct ColorMeta {
    std::string sample_name;
    std::string graph_name;
    cc_color_info info{};
};
struct Header {
    uint32_t version = 0, k = 0, s = 0, c = 0;
    uint64_t data_offset = 0, record_size = 0, num_records = 0;
    std::vector<ColorMeta> colors;
};
// Parses a header image; returns cc_status and fills h (num_records computed from total_size).
int parse_header(const uint8_t *buf, uint64_t avail, uint64_t total_size, const char *path_for_msg, Header &h);
// FindROIs.makeCortexHeader + CortexGraphWriter.initialize: the 1-colour ROI header.
std::vector<uint8_t> make_roi_header(uint32_t k, uint32_t s, const std::string &sample_name);

// ------------------------------------------------------------------ device workspace for the scan
struct ScanWorkspace {
    uint64_t *tile_state = nullptr;   // look-back descriptors
    uint8_t *scratch = nullptr;       // global staging lists of the fast scan kernel
    size_t scratch_bytes = 0;
    uint64_t *dirty_list = nullptr;   // {chunk, prefix} pairs queued by the fast scan for the rewrite kernel
    uint64_t tile_state_cap = 0;
    uint32_t *tile_counter = nullptr; // dynamic tile ticket
    uint64_t *totals = nullptr;       // [0] in, [1] out (ping-pong for chunked scans)
    int32_t *parents = nullptr;       // device copy of the parent colour list
    uint32_t parents_cap = 0;
    int *dev_error = nullptr;         // watchdog code (device view)
    int *host_error = nullptr;        // the same word in mapped host memory (null: plain device memory)
    bool poisoned = false;            // a launch failed after touching the counters: reset everything on the next ensure()
    uint32_t epoch = 0;
    uint32_t ticket_base = 0;         // tickets drawn so far from tile_counter (host mirror)
    int ensure(uint64_t ntiles, uint32_t nparents);
    int ensure_scratch(size_t bytes);
    void release();
};

// ------------------------------------------------------------------ lookup index
struct LookupIndex {
    uint64_t *keys = nullptr;     // [n*s] native words, word 0 first
    void *lines = nullptr;        // [nlines] 64-byte bucket lines (lookup.cu: "the lookup index")
    void *bins = nullptr;         // [nbins] {first line, lines} per equal slice of the array's own key range
    uint64_t nlines = 0;
    uint64_t base = 0;            // top 64 bits of the first key
    uint64_t span = 0;            // top 64 bits of the last key - base
    uint64_t scale = 0;           // see key_bin
    uint32_t norm = 0, nbins = 1;
    uint32_t pad[8] = {};         // wire form of the largest key (fills unused slots)
    bool built = false;
    bool sorted = true;
    uint64_t unsorted_at = 0;
};

}  // namespace cc

struct cc_graph {
    int device = 0;
    cc::Header h;
    std::string path;
    // host image (mmap or caller copy) -- only for cc_get_records and the upload
    const uint8_t *host_image = nullptr;
    uint64_t host_size = 0;
    void *map_base = nullptr;
    uint64_t map_len = 0;
    std::vector<uint8_t> owned_image;
    // device body
    const uint8_t *dev_body = nullptr;   // record 0
    void *dev_alloc = nullptr;           // owned allocation (null when wrapping a caller buffer)
    bool dev_alloc_pooled = false;       // dev_alloc came from the stream-ordered pool (cudaFreeAsync)
    uint64_t first_index = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cc::ScanWorkspace scan_ws;
    cc::LookupIndex index;
    cc_stats stats{};
    int sm_count = 148;
    // staging for the host-output novelty scan (grown on demand, kept across calls)
    void *novel_buf = nullptr;
    void *novel_idx = nullptr;
    uint64_t novel_cap = 0;
    std::vector<int32_t> parents_cached;   // what scan_ws.parents currently holds
    // device staging of the host-buffer lookups (cc_find_ascii / windows / packed): two chunks in flight, kept across calls
    void *look_in[2] = {nullptr, nullptr}, *look_flag[2] = {nullptr, nullptr}, *look_out[2] = {nullptr, nullptr};
    size_t look_in_cap = 0, look_q_cap = 0;
    cudaStream_t look_stream = nullptr;        // second stream: the copies of one chunk overlap the search of the other
    cudaEvent_t look_done = nullptr;
    // mapped pinned staging of cc_find_records (the low-latency findRecord path): queries | indices | record bytes
    uint8_t *small_host = nullptr, *small_dev = nullptr;
    size_t small_bytes = 0;
};

namespace cc {

// ------------------------------------------------------------------ kernel launchers (defined in the .cu files)
struct ScanArgs {
    const uint8_t *body;   // device, record 0 of this launch
    uint64_t n;            // records in this launch
    uint64_t index_base;   // global index of record 0 (for out_index)
    uint32_t k, s, c;
    int32_t child;
    int nparents;          // device list in ws.parents
    const int32_t *parent_list;   // host copy of the same list
    uint8_t *out_records;  // device
    uint64_t *out_index;   // device or null
    uint64_t cap;
    const uint64_t *total_in;   // device or null (=0)
    uint64_t *total_out;        // device
};
int launch_scan_novel(const ScanArgs &a, ScanWorkspace &ws, int sm_count, cudaStream_t st);
// Number of tiles launch_scan_novel / launch_decode_columns will cut n records into (current options).
uint64_t scan_tiles_for(uint64_t n, uint32_t s, uint32_t c);
int launch_decode_columns(const uint8_t *dev_body, uint64_t n, uint32_t s, uint32_t c,
                          uint64_t *dev_words, int32_t *dev_cov, uint8_t *dev_edges, ScanWorkspace &ws, int sm_count,
                          cudaStream_t st);

int build_index(cc_graph *g, int bits);
int launch_pack_windows(const uint8_t *dev_seq, uint64_t len, uint32_t k, uint64_t *dev_words, uint8_t *dev_flags,
                        uint64_t row_stride /*1 for sliding windows, k for independent rows*/, uint64_t nq, cudaStream_t st);
int launch_find_packed(cc_graph *g, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, int64_t *dev_index,
                       int algo, cudaStream_t st);
int launch_find_small(cc_graph *g, const uint8_t *dev_kmers, uint32_t nq, int64_t *dev_index, uint8_t *dev_raw, cudaStream_t st);
int launch_find_seq(cc_graph *g, const uint8_t *dev_seq, uint64_t len, uint64_t row_stride, uint64_t nq, int64_t *dev_index,
                    int algo, cudaStream_t st);
int launch_bucket_by_owner(const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t s,
                           const uint64_t *dev_splitters, int nshards, uint64_t *dev_counts, uint64_t *dev_sorted_words,
                           uint32_t *dev_slots, cudaStream_t st);
int launch_scatter_results(const int64_t *dev_values, const uint32_t *dev_slots, uint64_t n, int64_t *dev_out, cudaStream_t st);
int join_pair(const uint8_t *body_a, const uint64_t *keys_a, uint64_t na, uint32_t ca, const uint8_t *body_b, const uint64_t *keys_b,
              uint64_t nb, uint32_t cb, uint32_t s, cudaStream_t st, void **out_body, uint64_t *out_n, uint64_t **out_keys = nullptr);
int sort_permutation(const uint64_t *dev_words, uint64_t n, uint32_t s, uint32_t k, cudaStream_t st, uint32_t **perm_out);
int launch_gather_records(const uint8_t *body, const uint32_t *perm, uint64_t n, uint32_t S, uint8_t *out, cudaStream_t st);
// prefilter.cu
int launch_lowcov_flags(const int32_t *cov, uint64_t n, uint32_t c, int32_t min_cov, uint8_t *flags, int sm_count, cudaStream_t st);
int launch_remove_flags(const int32_t *cov, uint64_t n, uint32_t c, uint32_t c_primary, uint8_t *flags, int sm_count, cudaStream_t st);
int launch_recover_classes(const int32_t *cov, uint64_t n, uint32_t c, uint32_t child, uint8_t *cls, uint8_t *find_flags, int sm_count,
                           cudaStream_t st);
int launch_recover_finalize(const uint8_t *cls, const int64_t *idx, const uint8_t *dirty_body, uint32_t dirty_S, uint32_t s, uint64_t dirty_first,
                            uint64_t n, uint8_t *flags, int32_t *patch, unsigned long long *recovered, int sm_count, cudaStream_t st);
int launch_shared_flags(const int64_t *idx, const uint8_t *graph_body, uint32_t S, uint32_t s, uint32_t c, uint64_t graph_first,
                        const uint32_t *excluded, uint64_t nroi, uint8_t *flags, unsigned long long *missing_at, int sm_count, cudaStream_t st);
int select_flagged(const uint8_t *flags, uint64_t n, uint32_t **sel_out, uint64_t *m_out, cudaStream_t st);
int launch_project_records(const uint8_t *body, uint32_t s, uint32_t c_in, const uint32_t *sel, uint64_t m, uint32_t c_out,
                           const uint8_t *flags, const int32_t *patch, uint32_t patch_color, uint8_t *out, int sm_count, cudaStream_t st);
int cov_stats_fused(const uint8_t *body, uint64_t n, uint32_t s, uint32_t c, int32_t child, const uint32_t *dev_parent_mask, int sm_count,
                    cudaStream_t st, std::vector<int32_t> &out_cov, std::vector<long long> &out_count, bool *fell_back);
int cov_stats(const int32_t *cov, uint64_t n, uint32_t c, int32_t child, const uint32_t *dev_parent_mask, int sm_count, cudaStream_t st,
              std::vector<int32_t> &out_cov, std::vector<long long> &out_count);
uint64_t route_state_size(uint64_t max_q, int nshards);
int launch_route(const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t k, const uint64_t *dev_splitters, int nshards,
                 int my_rank, uint64_t cap, void *const *peer_inbox, void *const *peer_counts, void *dev_route_state, uint64_t max_q,
                 uint64_t *dev_sent, cudaStream_t st);
int launch_publish_counts(const uint64_t *dev_sent, int nshards, int my_rank, uint64_t cap, void *const *peer_counts, cudaStream_t st);
int launch_find_routed(cc_graph *g, const void *dev_inbox, const uint64_t *dev_counts_in, int world, int vsub, uint64_t cap,
                       void *dev_res, cudaStream_t st);
int launch_gather_routed(void *const *peer_res, const void *dev_route_state, uint64_t max_q, uint64_t nq, const uint64_t *dev_shard_first,
                         int nshards, uint64_t cap, int64_t *dev_out, cudaStream_t st);

}  // namespace cc
