// api.cu -- the extern "C" surface of libcorticall_cuda (include/corticall_cuda.h): handle lifecycle,
// .ctx header parsing, host<->device marshalling around the kernels of scan.cu / lookup.cu.
//
// Reference behaviour mirrored here (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
//   header parse + error texts   S/utils/io/graph/cortex/CortexGraph.java:66-168
//   numRecords floors            S/utils/io/graph/cortex/CortexGraph.java:148-149
//   getColorForSampleName        S/utils/io/graph/cortex/CortexGraph.java:335-354
//   ROI header                   S/commands/discover/roi/FindROIs.java:85-105 + CortexGraphWriter.java:31-104
//
// There is no CPU implementation of any compute entry point in this library: without a CUDA device they
// fail with CC_ERR_CUDA.
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <mutex>

#include "cc_internal.hpp"

namespace cc {

// ------------------------------------------------------------------ errors / options / counters
static thread_local std::string t_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}
int fail(int status, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
    return status;
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, base ? base + 1 : file, line);
    return CC_ERR_CUDA;
}

std::atomic<uint64_t> g_launches{0};

Options &options() {
    static Options o;
    return o;
}

// ------------------------------------------------------------------ header (ctx_spec.md tables 1-3)
namespace {

struct Cursor {
    const uint8_t *p;
    uint64_t avail, pos = 0;
    bool ok = true;
    bool need(uint64_t n) {
        if (!ok || pos + n > avail) { ok = false; return false; }
        return true;
    }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t v;
        memcpy(&v, p + pos, 4);
        pos += 4;
        return v;
    }
    uint64_t u64() {
        if (!need(8)) return 0;
        uint64_t v;
        memcpy(&v, p + pos, 8);
        pos += 8;
        return v;
    }
    uint8_t u8() {
        if (!need(1)) return 0;
        return p[pos++];
    }
    // fixStringsWithEarlyTerminators (CortexGraph.java:50-64): cut at the first NUL
    std::string str(uint64_t n) {
        if (!need(n)) return std::string();
        const char *b = reinterpret_cast<const char *>(p + pos);
        pos += n;
        return std::string(b, strnlen(b, n));
    }
};

bool magic_ok(const uint8_t *p) { return strncasecmp(reinterpret_cast<const char *>(p), "CORTEX", 6) == 0; }

}  // namespace

int parse_header(const uint8_t *buf, uint64_t avail, uint64_t total_size, const char *path, Header &h) {
    Cursor c{buf, avail};
    if (avail < 6 || !magic_ok(buf))
        return fail(CC_ERR_NOT_CORTEX, "The file '%s' does not appear to be a Cortex graph", path);
    c.pos = 6;
    h.version = c.u32();
    if (!c.ok) return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': truncated header", path);
    if (h.version != 6) return fail(CC_ERR_BAD_VERSION, "The file '%s' is not a version 6 Cortex graph", path);
    h.k = c.u32();
    h.s = c.u32();
    h.c = c.u32();
    if (!c.ok || (uint64_t)h.c * 12 > avail)
        return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': truncated header", path);
    h.colors.assign(h.c, ColorMeta());
    for (uint32_t i = 0; i < h.c; ++i) h.colors[i].info.mean_read_length = c.u32();
    for (uint32_t i = 0; i < h.c; ++i) h.colors[i].info.total_sequence = c.u64();
    for (uint32_t i = 0; i < h.c; ++i) {
        const uint32_t L = c.u32();
        h.colors[i].sample_name = c.str(L);
    }
    for (uint32_t i = 0; i < h.c; ++i) { c.need(16); c.pos += c.ok ? 16 : 0; }   // long double error rates: skipped like the reference
    for (uint32_t i = 0; i < h.c; ++i) {
        cc_color_info &ci = h.colors[i].info;
        ci.tip_clipping = c.u8();
        ci.low_covg_supernodes_removed = c.u8();
        ci.low_covg_kmers_removed = c.u8();
        ci.cleaned_against_graph = c.u8();
        ci.low_cov_supernodes_threshold = c.u32();
        ci.low_cov_kmer_threshold = c.u32();
        const uint32_t G = c.u32();
        h.colors[i].graph_name = c.str(G);
    }
    if (!c.need(6)) return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': truncated header", path);
    if (!magic_ok(buf + c.pos))
        return fail(CC_ERR_BAD_TRAILER, "We didn't see a proper header terminator at the expected place in Cortex graph '%s'", path);
    c.pos += 6;
    h.data_offset = c.pos;
    // kmer_bits must be the word count of kmer_size (CortexRecord.getKmerBits :309-311); a header that disagrees would make every
    // shift of the 2-bit arithmetic undefined
    if (h.k == 0 || h.s != (h.k + 31) / 32)
        return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': kmer_bits %u does not match kmer_size %u", path, h.s, h.k);
    h.record_size = 8ull * h.s + 5ull * h.c;
    if (h.record_size == 0) return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': zero-sized records", path);
    h.num_records = (total_size - h.data_offset) / h.record_size;       // floors: trailing partial record ignored
    return CC_OK;
}

std::vector<uint8_t> make_roi_header(uint32_t k, uint32_t s, const std::string &name) {
    static const uint8_t err[16] = {0, 0xd8, 0xa3, 0x70, 0x3d, 0x0a, 0xd7, 0xa3, 0xf8, 0x3f, 0, 0, 0, 0, 0, 0};   // CortexGraphWriter.java:76
    std::vector<uint8_t> out;
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; ++i) out.push_back((uint8_t)(v >> (8 * i))); };
    auto raw = [&](const void *p, size_t n) { out.insert(out.end(), (const uint8_t *)p, (const uint8_t *)p + n); };
    raw("CORTEX", 6);
    u32(6); u32(k); u32(s); u32(1);
    u32(0);                 // mean read length
    u32(0); u32(0);         // total sequence
    u32((uint32_t)name.size()); raw(name.data(), name.size());
    raw(err, 16);
    u32(0);                 // four cleaning booleans
    u32(0); u32(0);         // thresholds
    u32(0);                 // cleaned-against graph name: empty
    raw("CORTEX", 6);
    return out;
}

namespace {

// ------------------------------------------------------------------ handle plumbing
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CC_ERR_CUDA, "no CUDA device available (%s); libcorticall_cuda has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(CC_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    return CC_OK;
}

int init_handle(cc_graph *g, int device) {
    g->device = device;
    CC_CUDA(cudaSetDevice(device));
    CC_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    CC_CUDA(cudaEventCreate(&g->ev0));
    CC_CUDA(cudaEventCreate(&g->ev1));
    CC_CUDA(cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device));
    // Scratch for batched lookups comes from the stream-ordered pool; keep freed blocks cached instead of
    // returning them to the driver at every synchronisation (re-mapping gigabytes per call costs milliseconds).
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    return CC_OK;
}

int upload_body(cc_graph *g) {
    const uint64_t bytes = g->h.num_records * g->h.record_size;
    void *d = nullptr;
    CC_CUDA(cudaMalloc(&d, bytes + 256));            // slack: tiles are fetched as 16-byte-aligned supersets
    g->dev_alloc = d;
    g->dev_body = static_cast<const uint8_t *>(d);
    const uint8_t *src = g->host_image + g->h.data_offset;
    const uint64_t chunk = 256ull << 20;
    for (uint64_t off = 0; off < bytes; off += chunk) {
        const uint64_t nb = std::min(chunk, bytes - off);
        CC_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(d) + off, src + off, nb, cudaMemcpyHostToDevice, g->stream));
    }
    CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(d) + bytes, 0, 256, g->stream));
    CC_CUDA(cudaStreamSynchronize(g->stream));
    return CC_OK;
}

void destroy(cc_graph *g) {
    if (!g) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    g->scan_ws.release();
    if (g->index.keys) cudaFree(g->index.keys);
    if (g->index.lines) cudaFree(g->index.lines);
    if (g->index.bins) cudaFree(g->index.bins);
    if (g->novel_buf) cudaFree(g->novel_buf);
    if (g->novel_idx) cudaFree(g->novel_idx);
    if (g->small_host) cudaFreeHost(g->small_host);
    for (int b = 0; b < 2; ++b) {
        if (g->look_in[b]) cudaFree(g->look_in[b]);
        if (g->look_flag[b]) cudaFree(g->look_flag[b]);
        if (g->look_out[b]) cudaFree(g->look_out[b]);
    }
    if (g->look_stream) cudaStreamDestroy(g->look_stream);
    if (g->look_done) cudaEventDestroy(g->look_done);
    if (g->dev_alloc) {
        if (g->dev_alloc_pooled && g->stream) { cudaFreeAsync(g->dev_alloc, g->stream); cudaStreamSynchronize(g->stream); }
        else cudaFree(g->dev_alloc);
    }
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->stream) cudaStreamDestroy(g->stream);
    if (g->map_base) munmap(g->map_base, g->map_len);
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
    delete g;
}

int finish_open(std::unique_ptr<cc_graph, void (*)(cc_graph *)> &g, int device, cc_graph **out) {
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    if (int rc = init_handle(g.get(), device)) return rc;
    if (int rc = upload_body(g.get())) return rc;
    *out = g.release();
    return CC_OK;
}

#define CC_REQUIRE(cond, ...)                                   \
    do {                                                        \
        if (!(cond)) return ::cc::fail(CC_ERR_ARG, __VA_ARGS__); \
    } while (0)

int check_colors(const cc_graph *g, int32_t child, const int32_t *parents, int nparents) {
    // The reference indexes int[] coverages with these: out of range = ArrayIndexOutOfBoundsException
    // (e.g. getColorForSampleName returned -1).
    if (child < 0 || (uint32_t)child >= g->h.c)
        return fail(CC_ERR_ARG, "child colour %d out of range (graph has %u colours)", child, g->h.c);
    if (nparents < 0 || (nparents > 0 && !parents)) return fail(CC_ERR_ARG, "bad parent list");
    for (int i = 0; i < nparents; ++i)
        if (parents[i] < 0 || (uint32_t)parents[i] >= g->h.c)
            return fail(CC_ERR_ARG, "parent colour %d out of range (graph has %u colours)", parents[i], g->h.c);
    return CC_OK;
}

// Copies the parent list to the workspace (skipped when unchanged) and sizes the look-back state.
int prepare_scan(cc_graph *g, uint64_t n_per_launch, const int32_t *parents, int nparents, cudaStream_t st) {
    if (int rc = g->scan_ws.ensure(scan_tiles_for(n_per_launch, g->h.s, g->h.c), (uint32_t)nparents)) return rc;
    if (nparents > 0) {
        if (g->parents_cached.size() != (size_t)nparents || memcmp(g->parents_cached.data(), parents, 4 * (size_t)nparents) != 0) {
            g->parents_cached.assign(parents, parents + nparents);
            CC_CUDA(cudaMemcpyAsync(g->scan_ws.parents, g->parents_cached.data(), 4 * (size_t)nparents, cudaMemcpyHostToDevice, st));
        }
    }
    return CC_OK;
}

int sync_stream(cc_graph *g, cudaStream_t st) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        const int code = (g && g->scan_ws.host_error) ? *static_cast<volatile int *>(g->scan_ws.host_error) : 0;
        if (g) g->scan_ws.poisoned = true;
        cuda_fail(e, "cudaStreamSynchronize", __FILE__, __LINE__);
        if (code) set_error("%s (device watchdog code %d: %s)", t_error.c_str(), code,
                            code == 1 ? "tile never arrived" : code == 2 ? "stage never released" : "look-back never resolved");
        return CC_ERR_CUDA;
    }
    return CC_OK;
}

int ensure_index(cc_graph *g) {
    if (!g->index.built) {
        if (int rc = build_index(g, 0)) return rc;
    }
    if (!g->index.sorted)
        return fail(CC_ERR_UNSORTED, "Records are not sorted (record %llu sorts before its predecessor)",
                    (unsigned long long)g->index.unsorted_at);
    return CC_OK;
}

// Scratch from the device's stream-ordered pool (cached across calls: init_handle raises the release threshold).
struct PoolBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~PoolBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t n, cudaStream_t stream) {
        st = stream;
        CC_CUDA(cudaMallocAsync(&p, std::max<size_t>(n, 16), stream));
        return CC_OK;
    }
    template <class T> T *as() { return static_cast<T *>(p); }
};

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        CC_CUDA(cudaMalloc(&p, std::max<size_t>(n, 16)));
        return CC_OK;
    }
    template <class T> T *as() { return static_cast<T *>(p); }
};

}  // namespace
}  // namespace cc

using namespace cc;

// ==================================================================== library
extern "C" {

const char *cc_last_error(void) { return t_error.c_str(); }
const char *cc_version(void) { return "corticall_cuda 0.1 (sm_100a)"; }

int cc_device_count(int *out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (out) *out = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CC_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(e));
    }
    if (n == 0) return fail(CC_ERR_CUDA, "no CUDA device available (device count is 0)");
    return CC_OK;
}

uint64_t cc_launch_count(void) { return g_launches.load(); }

int cc_set_option(const char *name, int64_t value) {
    if (!name) return fail(CC_ERR_ARG, "option name is null");
    Options &o = options();
    if (!strcmp(name, "scan_stages")) o.scan_stages = (int)value;
    else if (!strcmp(name, "scan_tile_bytes")) o.scan_tile_bytes = (int)value;
    else if (!strcmp(name, "scan_ctas_per_sm")) o.scan_ctas_per_sm = (int)value;
    else if (!strcmp(name, "index_bits")) o.index_bits = (int)value;
    else if (!strcmp(name, "index_fill_pct")) o.index_fill_pct = (int)value;
    else if (!strcmp(name, "find_bins_smem")) o.find_bins_smem = (int)value;
    else if (!strcmp(name, "rows_fused")) o.rows_fused = (int)value;
    else if (!strcmp(name, "rows_warp")) o.rows_warp = (int)value;
    else if (!strcmp(name, "covstats_fused")) o.covstats_fused = (int)value;
    else if (!strcmp(name, "join_tiled")) o.join_tiled = (int)value;
    else if (!strcmp(name, "join_tile_kb")) o.join_tile_kb = (int)value;
    else if (!strcmp(name, "rows_rpt2_max_k")) o.rows_rpt2_max_k = (int)value;
    else if (!strcmp(name, "route_blocks_per_sm")) o.route_blocks_per_sm = (int)value;
    else if (!strcmp(name, "route_stage_depth")) o.route_stage_depth = (int)value;
    else if (!strcmp(name, "routed_search_blocks_per_sm")) o.routed_search_blocks_per_sm = (int)value;
    else if (!strcmp(name, "gather_blocks_per_sm")) o.gather_blocks_per_sm = (int)value;
    else if (!strcmp(name, "host_chunk_mb")) o.host_chunk_mb = (int)value;
    else if (!strcmp(name, "scan_debug")) o.scan_debug = (int)value;
    else if (!strcmp(name, "scan_pdl")) o.scan_pdl = (int)value;
    else if (!strcmp(name, "scan_fast")) o.scan_fast = (int)value;
    else if (!strcmp(name, "scan_stage_buf_bytes")) o.scan_stage_buf_bytes = (int)value;
    else if (!strcmp(name, "scan_chunk_tiles")) o.scan_chunk_tiles = (int)value;
    else if (!strcmp(name, "lookup_l2_hints")) o.lookup_l2_hints = (int)value;
    else return fail(CC_ERR_ARG, "unknown option '%s'", name);
    return CC_OK;
}

// ==================================================================== lifecycle
int cc_open(const char *path, int device, cc_graph **out) {
    if (!path || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> g(new cc_graph(), destroy);
    g->path = path;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        if (errno == ENOENT) return fail(CC_ERR_IO, "Cortex graph file '%s' not found: %s", path, strerror(errno));
        return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': %s", path, strerror(errno));
    }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': %s", path, strerror(errno)); }
    const uint64_t size = (uint64_t)sb.st_size;
    if (size == 0) { close(fd); return fail(CC_ERR_NOT_CORTEX, "The file '%s' does not appear to be a Cortex graph", path); }
    void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail(CC_ERR_IO, "Error while parsing Cortex graph file '%s': mmap: %s", path, strerror(errno));
    g->map_base = m;
    g->map_len = size;
    g->host_image = static_cast<const uint8_t *>(m);
    g->host_size = size;
    if (int rc = parse_header(g->host_image, size, size, path, g->h)) return rc;
    return finish_open(g, device, out);
}

int cc_open_memory(const void *file_image, uint64_t size, int device, cc_graph **out) {
    if (!file_image || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> g(new cc_graph(), destroy);
    g->path = "<memory>";
    g->owned_image.assign(static_cast<const uint8_t *>(file_image), static_cast<const uint8_t *>(file_image) + size);
    g->host_image = g->owned_image.data();
    g->host_size = size;
    if (int rc = parse_header(g->host_image, size, size, "<memory>", g->h)) return rc;
    return finish_open(g, device, out);
}

int cc_open_device(const void *dev_body, uint32_t k, uint32_t s, uint32_t c, uint64_t n, uint64_t first_index, int device,
                   cc_graph **out) {
    if (!out || (!dev_body && n)) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    if (k == 0 || s != (k + 31) / 32) return fail(CC_ERR_ARG, "kmer_bits %u does not match kmer_size %u", s, k);
    if (int rc = check_device(device)) return rc;
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> g(new cc_graph(), destroy);
    g->path = "<device>";
    g->h.version = 6; g->h.k = k; g->h.s = s; g->h.c = c;
    g->h.record_size = 8ull * s + 5ull * c;
    g->h.num_records = n;
    g->h.colors.assign(c, ColorMeta());
    for (uint32_t i = 0; i < c; ++i) g->h.colors[i].sample_name = std::to_string(i);
    g->dev_body = static_cast<const uint8_t *>(dev_body);
    g->first_index = first_index;
    DeviceGuard guard(device);
    if (int rc = init_handle(g.get(), device)) return rc;
    // The caller's buffer may still be being written on another stream (the handle's own stream is
    // non-blocking): settle the device once here so that every later call sees the finished array.
    CC_CUDA(cudaDeviceSynchronize());
    *out = g.release();
    return CC_OK;
}

void cc_dispose(cc_graph *g) { destroy(g); }

// ==================================================================== header / colours
int cc_header(const cc_graph *g, uint32_t *version, uint32_t *kmer_size, uint32_t *kmer_bits, uint32_t *num_colors,
              uint64_t *num_records, uint64_t *data_offset, uint64_t *record_size) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    if (version) *version = g->h.version;
    if (kmer_size) *kmer_size = g->h.k;
    if (kmer_bits) *kmer_bits = g->h.s;
    if (num_colors) *num_colors = g->h.c;
    if (num_records) *num_records = g->h.num_records;
    if (data_offset) *data_offset = g->h.data_offset;
    if (record_size) *record_size = g->h.record_size;
    return CC_OK;
}

static int copy_string(const std::string &s, char *buf, size_t cap) {
    if (!buf || cap < s.size() + 1) return fail(CC_ERR_ARG, "buffer too small (%zu needed)", s.size() + 1);
    memcpy(buf, s.data(), s.size());
    buf[s.size()] = 0;
    return CC_OK;
}

int cc_color_name(const cc_graph *g, uint32_t color, char *buf, size_t cap) {
    if (!g || color >= g->h.c) return fail(CC_ERR_ARG, "colour %u out of range", color);
    return copy_string(g->h.colors[color].sample_name, buf, cap);
}
int cc_color_graph_name(const cc_graph *g, uint32_t color, char *buf, size_t cap) {
    if (!g || color >= g->h.c) return fail(CC_ERR_ARG, "colour %u out of range", color);
    return copy_string(g->h.colors[color].graph_name, buf, cap);
}
int cc_color_info_get(const cc_graph *g, uint32_t color, cc_color_info *out) {
    if (!g || !out || color >= g->h.c) return fail(CC_ERR_ARG, "colour %u out of range", color);
    *out = g->h.colors[color].info;
    return CC_OK;
}

int cc_color_for_sample_name(const cc_graph *g, const char *name, int32_t *out_color) {
    if (!g || !name || !out_color) return fail(CC_ERR_ARG, "null argument");
    int32_t color = -1;
    int copies = 0;
    for (uint32_t c = 0; c < g->h.c; ++c) {
        if (strcasecmp(g->h.colors[c].sample_name.c_str(), name) == 0) { color = (int32_t)c; ++copies; }
    }
    if (color == -1) {       // Integer.valueOf(sampleName): optional sign, decimal digits only
        const char *p = name;
        if (*p == '+' || *p == '-') ++p;
        bool digits = *p != 0;
        for (const char *q = p; *q; ++q) digits &= (*q >= '0' && *q <= '9');
        if (digits) {                     // Integer.valueOf throws NumberFormatException outside int range: then the colour stays -1
            errno = 0;
            const long long v = strtoll(name, nullptr, 10);
            if (errno == 0 && v >= INT32_MIN && v <= INT32_MAX) { color = (int32_t)v; copies = 1; }
        }
    }
    *out_color = (copies == 1) ? color : -1;
    return CC_OK;
}

// ==================================================================== K1: record access / decode
int cc_get_records(const cc_graph *g, uint64_t first, uint64_t count, void *out_raw) {
    if (!g || (!out_raw && count)) return fail(CC_ERR_ARG, "null argument");
    if (first > g->h.num_records || count > g->h.num_records - first)
        return fail(CC_ERR_RANGE, "Record index is out of range (%llu+%llu vs 0-%lld)", (unsigned long long)first,
                    (unsigned long long)count, (long long)g->h.num_records - 1);
    const uint64_t S = g->h.record_size;
    if (g->host_image) {
        memcpy(out_raw, g->host_image + g->h.data_offset + first * S, count * S);
        return CC_OK;
    }
    DeviceGuard guard(g->device);
    CC_CUDA(cudaMemcpy(out_raw, g->dev_body + first * S, count * S, cudaMemcpyDeviceToHost));
    return CC_OK;
}

int cc_decode_records_dev(cc_graph *g, uint64_t first, uint64_t count, uint64_t *dev_words, int32_t *dev_coverage,
                          uint8_t *dev_edges, void *stream) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    if (first > g->h.num_records || count > g->h.num_records - first)
        return fail(CC_ERR_RANGE, "Record index is out of range (%llu+%llu vs 0-%lld)", (unsigned long long)first,
                    (unsigned long long)count, (long long)g->h.num_records - 1);
    DeviceGuard guard(g->device);
    return launch_decode_columns(g->dev_body + first * g->h.record_size, count, g->h.s, g->h.c, dev_words, dev_coverage, dev_edges,
                                 g->scan_ws, g->sm_count, static_cast<cudaStream_t>(stream));
}

int cc_decode_records(cc_graph *g, uint64_t first, uint64_t count, uint64_t *out_words, int32_t *out_coverage, uint8_t *out_edges) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    DeviceGuard guard(g->device);
    DevBuf w, c, e;
    if (out_words) if (int rc = w.alloc(count * g->h.s * 8)) return rc;
    if (out_coverage) if (int rc = c.alloc(count * g->h.c * 4)) return rc;
    if (out_edges) if (int rc = e.alloc(count * g->h.c)) return rc;
    if (int rc = cc_decode_records_dev(g, first, count, w.as<uint64_t>(), c.as<int32_t>(), e.as<uint8_t>(), g->stream)) return rc;
    if (out_words) CC_CUDA(cudaMemcpyAsync(out_words, w.p, count * g->h.s * 8, cudaMemcpyDeviceToHost, g->stream));
    if (out_coverage) CC_CUDA(cudaMemcpyAsync(out_coverage, c.p, count * g->h.c * 4, cudaMemcpyDeviceToHost, g->stream));
    if (out_edges) CC_CUDA(cudaMemcpyAsync(out_edges, e.p, count * g->h.c, cudaMemcpyDeviceToHost, g->stream));
    return sync_stream(g, g->stream);
}

// ==================================================================== K1+K2: novelty scan
int cc_find_novel_dev(cc_graph *g, int32_t child, const int32_t *parents, int nparents, void *dev_out_records,
                      uint64_t *dev_out_index, uint64_t cap, uint64_t *dev_out_count, void *stream) {
    if (!g || !dev_out_count || (cap && !dev_out_records)) return fail(CC_ERR_ARG, "null argument");
    if (int rc = check_colors(g, child, parents, nparents)) return rc;
    DeviceGuard guard(g->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = prepare_scan(g, g->h.num_records, parents, nparents, st)) return rc;
    ScanArgs a{};
    a.body = g->dev_body; a.n = g->h.num_records; a.index_base = g->first_index;
    a.k = g->h.k; a.s = g->h.s; a.c = g->h.c;
    a.child = child; a.nparents = nparents; a.parent_list = parents;
    a.out_records = static_cast<uint8_t *>(dev_out_records); a.out_index = dev_out_index; a.cap = cap;
    a.total_in = nullptr; a.total_out = dev_out_count;
    return launch_scan_novel(a, g->scan_ws, g->sm_count, st);
}

int cc_find_novel(cc_graph *g, int32_t child, const int32_t *parents, int nparents, void *out_records, uint64_t *out_index,
                  uint64_t cap, uint64_t *out_count) {
    if (!g || !out_count) return fail(CC_ERR_ARG, "null argument");
    if (int rc = check_colors(g, child, parents, nparents)) return rc;
    DeviceGuard guard(g->device);
    const uint64_t n = g->h.num_records, O = 8ull * g->h.s + 5;
    const bool want_index = out_index != nullptr;
    uint64_t want = std::min(cap, n);
    if (!out_records) want = 0;
    // Device staging sized for the expected (small) novel fraction; grown and re-run if it overflowed.
    uint64_t dcap = std::min<uint64_t>(want, std::max<uint64_t>(65536, n / 32));
    g->stats = cc_stats{};
    const uint64_t launches0 = g_launches.load();
    uint64_t total = 0;
    float kernel_ms = 0.f;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (dcap > g->novel_cap) {
            if (g->novel_buf) { cudaFree(g->novel_buf); g->novel_buf = nullptr; }
            if (g->novel_idx) { cudaFree(g->novel_idx); g->novel_idx = nullptr; }
            g->novel_cap = 0;
            CC_CUDA(cudaMalloc(&g->novel_buf, dcap * O + 64));
            CC_CUDA(cudaMalloc(&g->novel_idx, dcap * 8 + 64));
            g->novel_cap = dcap;
        }
        if (int rc = g->scan_ws.ensure(0, 0)) return rc;
        uint64_t *d_count = g->scan_ws.totals + 4;
        CC_CUDA(cudaEventRecord(g->ev0, g->stream));
        if (int rc = cc_find_novel_dev(g, child, parents, nparents, g->novel_buf, want_index ? static_cast<uint64_t *>(g->novel_idx) : nullptr, dcap, d_count, g->stream)) return rc;
        CC_CUDA(cudaEventRecord(g->ev1, g->stream));
        CC_CUDA(cudaMemcpyAsync(&total, d_count, 8, cudaMemcpyDeviceToHost, g->stream));
        if (int rc = sync_stream(g, g->stream)) return rc;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g->ev0, g->ev1);
        kernel_ms += ms;
        g->stats.d2h_bytes += 8;
        const uint64_t need = std::min(total, want);
        if (need <= dcap) break;
        dcap = need;
    }
    const uint64_t got = std::min(total, want);
    if (got) {
        CC_CUDA(cudaMemcpyAsync(out_records, g->novel_buf, got * O, cudaMemcpyDeviceToHost, g->stream));
        if (want_index) CC_CUDA(cudaMemcpyAsync(out_index, g->novel_idx, got * 8, cudaMemcpyDeviceToHost, g->stream));
        if (int rc = sync_stream(g, g->stream)) return rc;
        g->stats.d2h_bytes += got * O + (want_index ? got * 8 : 0);
    }
    *out_count = total;
    g->stats.kernel_ms = kernel_ms;
    g->stats.total_ms = kernel_ms;
    g->stats.launches = (uint32_t)(g_launches.load() - launches0);
    return CC_OK;
}

int cc_find_novel_host(int device, const void *host_body, uint32_t k, uint32_t s, uint32_t c, uint64_t n, int32_t child,
                       const int32_t *parents, int nparents, void *out_records, uint64_t *out_index, uint64_t cap,
                       uint64_t *out_count, cc_stats *stats) {
    if (!out_count || (!host_body && n)) return fail(CC_ERR_ARG, "null argument");
    if (int rc = check_device(device)) return rc;
    if (k == 0 || s != (k + 31) / 32) return fail(CC_ERR_ARG, "kmer_bits %u does not match kmer_size %u", s, k);
    DeviceGuard guard(device);
    // A handle over nothing, kept per device for the life of the process: only its workspace (look-back state, staging
    // scratch -- tens of MB, too costly to allocate per call), stream and header fields are used.
    static std::mutex mu[64];                  // one per device: scans on different GPUs do not serialise each other
    static cc_graph *cache[64] = {};
    if (device >= 64) return fail(CC_ERR_ARG, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(mu[device]);
    if (!cache[device]) {
        if (int rc = cc_open_device(nullptr, k, s, c, 0, 0, device, &cache[device])) return rc;
    }
    struct Borrow { cc_graph *p; cc_graph *get() const { return p; } cc_graph *operator->() const { return p; } } g{cache[device]};
    g->h.k = k; g->h.s = s; g->h.c = c;
    g->h.record_size = 8ull * s + 5ull * c;
    g->h.num_records = n;
    g->parents_cached.clear();
    if (int rc = check_colors(g.get(), child, parents, nparents)) return rc;

    const uint64_t S = g->h.record_size, O = 8ull * s + 5;
    // chunk = whole records, a multiple of 32 records so that every chunk starts 16-byte aligned on the device
    uint64_t chunk_rec = (((uint64_t)std::max(1, options().host_chunk_mb) << 20) / S) & ~31ull;
    if (chunk_rec < 32) chunk_rec = 32;
    const uint64_t nchunks = n ? (n + chunk_rec - 1) / chunk_rec : 0;
    constexpr int NB = 3;
    PoolBuf buf[NB];
    cudaStream_t copy_st = nullptr;
    cudaEvent_t copied[NB] = {}, consumed[NB] = {}, t0 = nullptr, t1 = nullptr;
    struct Cleanup {
        cudaStream_t &s; cudaEvent_t *a, *b; cudaEvent_t &t0, &t1;
        ~Cleanup() {
            for (int i = 0; i < NB; ++i) { if (a[i]) cudaEventDestroy(a[i]); if (b[i]) cudaEventDestroy(b[i]); }
            if (t0) cudaEventDestroy(t0);
            if (t1) cudaEventDestroy(t1);
            if (s) cudaStreamDestroy(s);
        }
    } cleanup{copy_st, copied, consumed, t0, t1};
    CC_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    CC_CUDA(cudaEventCreate(&t0));
    CC_CUDA(cudaEventCreate(&t1));
    for (int i = 0; i < NB; ++i) {
        CC_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        CC_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
        if (i < (int)std::min<uint64_t>(nchunks, NB)) if (int rc = buf[i].alloc(chunk_rec * S + 256, g->stream)) return rc;
    }
    const uint64_t want = out_records ? std::min(cap, n) : 0;
    // device staging for the output: expected-small; the scan is re-run with a larger one if it overflows
    uint64_t dcap = std::min<uint64_t>(want, std::max<uint64_t>(65536, n / 32));
    uint64_t total = 0;
    const uint64_t launches0 = g_launches.load();
    uint64_t h2d = 0;
    float total_ms = 0.f;
    for (int attempt = 0; attempt < 2; ++attempt) {
        PoolBuf o, ix;
        if (int rc = o.alloc(dcap * O + 64, g->stream)) return rc;
        if (out_index) if (int rc = ix.alloc(dcap * 8 + 64, g->stream)) return rc;
        if (int rc = prepare_scan(g.get(), chunk_rec, parents, nparents, g->stream)) return rc;
        uint64_t *totals = g->scan_ws.totals;     // [0],[1] ping-pong
        CC_CUDA(cudaMemsetAsync(totals, 0, 16, g->stream));
        CC_CUDA(cudaEventRecord(t0, g->stream));
        CC_CUDA(cudaStreamWaitEvent(copy_st, t0, 0));
        for (uint64_t ci = 0; ci < nchunks; ++ci) {
            const int b = (int)(ci % NB);
            const uint64_t r0 = ci * chunk_rec, nr = std::min(chunk_rec, n - r0);
            if (ci >= NB) CC_CUDA(cudaStreamWaitEvent(copy_st, consumed[b], 0));
            CC_CUDA(cudaMemcpyAsync(buf[b].p, static_cast<const uint8_t *>(host_body) + r0 * S, nr * S, cudaMemcpyHostToDevice, copy_st));
            CC_CUDA(cudaEventRecord(copied[b], copy_st));
            CC_CUDA(cudaStreamWaitEvent(g->stream, copied[b], 0));
            h2d += nr * S;
            ScanArgs a{};
            a.body = buf[b].as<uint8_t>(); a.n = nr; a.index_base = r0;
            a.k = k; a.s = s; a.c = c; a.child = child; a.nparents = nparents; a.parent_list = parents;
            a.out_records = o.as<uint8_t>(); a.out_index = out_index ? ix.as<uint64_t>() : nullptr; a.cap = dcap;
            a.total_in = totals + (ci & 1); a.total_out = totals + ((ci + 1) & 1);
            if (int rc = launch_scan_novel(a, g->scan_ws, g->sm_count, g->stream)) return rc;
            CC_CUDA(cudaEventRecord(consumed[b], g->stream));
        }
        CC_CUDA(cudaMemcpyAsync(&total, totals + (nchunks & 1), 8, cudaMemcpyDeviceToHost, g->stream));
        if (int rc = sync_stream(g.get(), g->stream)) return rc;
        const uint64_t need = std::min(total, want);
        if (need <= dcap) {
            if (need) {
                CC_CUDA(cudaMemcpyAsync(out_records, o.p, need * O, cudaMemcpyDeviceToHost, g->stream));
                if (out_index) CC_CUDA(cudaMemcpyAsync(out_index, ix.p, need * 8, cudaMemcpyDeviceToHost, g->stream));
            }
            CC_CUDA(cudaEventRecord(t1, g->stream));
            if (int rc = sync_stream(g.get(), g->stream)) return rc;
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t0, t1);
            total_ms += ms;
            if (stats) {
                stats->kernel_ms = 0.f;
                stats->total_ms = total_ms;
                stats->h2d_bytes = h2d;
                stats->d2h_bytes = 8 + need * O + (out_index ? need * 8 : 0);
                stats->launches = (uint32_t)(g_launches.load() - launches0);
            }
            break;
        }
        CC_CUDA(cudaEventRecord(t1, g->stream));
        if (int rc = sync_stream(g.get(), g->stream)) return rc;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        total_ms += ms;
        dcap = need;
    }
    *out_count = total;
    return CC_OK;
}

int cc_write_roi_file(cc_graph *g, int32_t child, const int32_t *parents, int nparents, const char *out_path, uint64_t *out_count) {
    if (!g || !out_path) return fail(CC_ERR_ARG, "null argument");
    if (int rc = check_colors(g, child, parents, nparents)) return rc;
    const uint64_t O = 8ull * g->h.s + 5;
    // first pass with the default staging gives the count; cc_find_novel re-runs itself when it overflowed
    std::vector<uint8_t> recs;
    uint64_t total = 0;
    uint64_t cap = std::max<uint64_t>(65536, g->h.num_records / 32);
    cap = std::min(cap, g->h.num_records);
    recs.resize(std::max<uint64_t>(cap, 1) * O);
    if (int rc = cc_find_novel(g, child, parents, nparents, recs.data(), nullptr, cap, &total)) return rc;
    if (total > cap) {
        cap = total;
        recs.resize(cap * O);
        if (int rc = cc_find_novel(g, child, parents, nparents, recs.data(), nullptr, cap, &total)) return rc;
    }
    const std::vector<uint8_t> hdr = make_roi_header(g->h.k, g->h.s, g->h.colors[child].sample_name);
    FILE *f = fopen(out_path, "wb");
    if (!f) return fail(CC_ERR_IO, "Cortex graph file '%s' not found: %s", out_path, strerror(errno));
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    if (total) ok = ok && fwrite(recs.data(), 1, total * O, f) == total * O;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(CC_ERR_IO, "Error while writing Cortex graph file '%s': %s", out_path, strerror(errno));
    if (out_count) *out_count = total;
    return CC_OK;
}

// ==================================================================== K3: canonicalise + pack
int cc_pack_canonical_dev(int device, const uint8_t *dev_seq, uint64_t len, uint32_t k, uint64_t *dev_words, uint8_t *dev_flags,
                          void *stream) {
    if (int rc = check_device(device)) return rc;
    if (k == 0) return fail(CC_ERR_ARG, "k must be positive");
    if (len < k) return CC_OK;
    if (!dev_seq || !dev_words || !dev_flags) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_pack_windows(dev_seq, len, k, dev_words, dev_flags, 1, len - k + 1, static_cast<cudaStream_t>(stream));
}

int cc_pack_kmers_dev(int device, const uint8_t *dev_kmers, uint64_t nq, uint32_t k, uint64_t *dev_words, uint8_t *dev_flags,
                      void *stream) {
    if (int rc = check_device(device)) return rc;
    if (k == 0) return fail(CC_ERR_ARG, "k must be positive");
    if (nq == 0) return CC_OK;
    if (!dev_kmers || !dev_words || !dev_flags) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_pack_windows(dev_kmers, nq * k, k, dev_words, dev_flags, k, nq, static_cast<cudaStream_t>(stream));
}

int cc_pack_canonical(int device, const uint8_t *seq, uint64_t len, uint32_t k, uint64_t *out_words, uint8_t *out_flags) {
    if (int rc = check_device(device)) return rc;
    if (k == 0) return fail(CC_ERR_ARG, "k must be positive");
    if (len < k) return CC_OK;
    if (!seq || !out_words || !out_flags) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    const uint64_t nw = len - k + 1, s = (k + 31) / 32;
    DevBuf dseq, dw, df;
    if (int rc = dseq.alloc(len + 64)) return rc;
    if (int rc = dw.alloc(nw * s * 8)) return rc;
    if (int rc = df.alloc(nw)) return rc;
    CC_CUDA(cudaMemcpy(dseq.p, seq, len, cudaMemcpyHostToDevice));
    if (int rc = launch_pack_windows(dseq.as<uint8_t>(), len, k, dw.as<uint64_t>(), df.as<uint8_t>(), 1, nw, nullptr)) return rc;
    CC_CUDA(cudaMemcpy(out_words, dw.p, nw * s * 8, cudaMemcpyDeviceToHost));
    CC_CUDA(cudaMemcpy(out_flags, df.p, nw, cudaMemcpyDeviceToHost));
    return CC_OK;
}

// ==================================================================== K4: lookups
int cc_build_index(cc_graph *g, int index_bits) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    DeviceGuard guard(g->device);
    if (int rc = build_index(g, index_bits)) return rc;
    if (!g->index.sorted)
        return fail(CC_ERR_UNSORTED, "Records are not sorted (record %llu sorts before its predecessor)",
                    (unsigned long long)g->index.unsorted_at);
    return CC_OK;
}

int cc_find_ascii_dev(cc_graph *g, const uint8_t *dev_kmers, uint64_t nq, int64_t *dev_index, int algo, void *stream) {
    if (!g || (nq && (!dev_kmers || !dev_index))) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    return launch_find_seq(g, dev_kmers, nq * g->h.k, g->h.k, nq, dev_index, algo, static_cast<cudaStream_t>(stream));
}

int cc_find_windows_dev(cc_graph *g, const uint8_t *dev_seq, uint64_t len, int64_t *dev_index, int algo, void *stream) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    if (len < g->h.k) return CC_OK;
    if (!dev_seq || !dev_index) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    return launch_find_seq(g, dev_seq, len, 1, len - g->h.k + 1, dev_index, algo, static_cast<cudaStream_t>(stream));
}

int cc_find_packed_dev(cc_graph *g, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, int64_t *dev_index, int algo,
                       void *stream) {
    if (!g || (nq && (!dev_words || !dev_index))) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    return launch_find_packed(g, dev_words, dev_flags, nq, dev_index, algo, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

namespace {
struct StageView {
    void *p;
    template <class T> T *as() { return static_cast<T *>(p); }
};

// Host-buffer lookups: the batch is streamed through the device in chunks on two alternating streams so that
// (with page-locked caller buffers) the copies of one chunk overlap the search of the other.
template <class Launch>
int chunked_lookup(cc_graph *g, const uint8_t *in, uint64_t in_bytes_per_q, uint64_t in_extra_bytes, const uint8_t *flags,
                   uint64_t nq, int64_t *out_index, Launch launch) {
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    g->stats = cc_stats{};
    const uint64_t launches0 = g_launches.load();
    const uint64_t chunk_q = std::max<uint64_t>(1, std::min<uint64_t>(nq, 1ull << 24));
    // the staging buffers, the second stream and its event live in the handle: a call allocates only when it needs more room
    // than any call before it (per-call cudaMalloc / cudaFree of ~2 GB cost more than the copies of a 3e7-query batch)
    if (!g->look_stream) CC_CUDA(cudaStreamCreateWithFlags(&g->look_stream, cudaStreamNonBlocking));
    if (!g->look_done) CC_CUDA(cudaEventCreateWithFlags(&g->look_done, cudaEventDisableTiming));
    cudaStream_t st[2] = {g->stream, g->look_stream};
    const int nb = nq > chunk_q ? 2 : 1;
    const size_t need_in = chunk_q * in_bytes_per_q + in_extra_bytes + 64;
    if (need_in > g->look_in_cap || chunk_q > g->look_q_cap || (nb == 2 && !g->look_in[1])) {
        CC_CUDA(cudaStreamSynchronize(st[0]));
        CC_CUDA(cudaStreamSynchronize(st[1]));
        const size_t in_cap = std::max(need_in, g->look_in_cap), q_cap = std::max<size_t>(chunk_q, g->look_q_cap);
        for (int b = 0; b < 2; ++b) {
            if (g->look_in[b]) { cudaFree(g->look_in[b]); g->look_in[b] = nullptr; }
            if (g->look_flag[b]) { cudaFree(g->look_flag[b]); g->look_flag[b] = nullptr; }
            if (g->look_out[b]) { cudaFree(g->look_out[b]); g->look_out[b] = nullptr; }
        }
        g->look_in_cap = g->look_q_cap = 0;
        for (int b = 0; b < nb; ++b) {
            CC_CUDA(cudaMalloc(&g->look_in[b], in_cap));
            CC_CUDA(cudaMalloc(&g->look_flag[b], q_cap + 64));
            CC_CUDA(cudaMalloc(&g->look_out[b], q_cap * 8 + 64));
        }
        g->look_in_cap = in_cap;
        g->look_q_cap = q_cap;
    }
    StageView din[2] = {{g->look_in[0]}, {g->look_in[1]}}, dflag[2] = {{g->look_flag[0]}, {g->look_flag[1]}}, dout[2] = {{g->look_out[0]}, {g->look_out[1]}};
    CC_CUDA(cudaEventRecord(g->ev0, g->stream));
    CC_CUDA(cudaStreamWaitEvent(st[1], g->ev0, 0));
    int b = 0;
    for (uint64_t q0 = 0; q0 < nq; q0 += chunk_q, b ^= (nb - 1)) {
        const uint64_t m = std::min(chunk_q, nq - q0);
        const uint64_t nbytes = m * in_bytes_per_q + in_extra_bytes;
        CC_CUDA(cudaMemcpyAsync(din[b].p, in + q0 * in_bytes_per_q, nbytes, cudaMemcpyHostToDevice, st[b]));
        if (flags) CC_CUDA(cudaMemcpyAsync(dflag[b].p, flags + q0, m, cudaMemcpyHostToDevice, st[b]));
        if (int rc = launch(din[b].as<uint8_t>(), flags ? dflag[b].as<uint8_t>() : nullptr, m, dout[b].as<int64_t>(), st[b])) return rc;
        CC_CUDA(cudaMemcpyAsync(out_index + q0, dout[b].p, m * 8, cudaMemcpyDeviceToHost, st[b]));
        g->stats.h2d_bytes += nbytes + (flags ? m : 0);
        g->stats.d2h_bytes += m * 8;
    }
    if (nb == 2) {
        CC_CUDA(cudaEventRecord(g->look_done, st[1]));
        CC_CUDA(cudaStreamWaitEvent(g->stream, g->look_done, 0));
    }
    CC_CUDA(cudaEventRecord(g->ev1, g->stream));
    if (int rc = sync_stream(g, g->stream)) return rc;
    cudaEventElapsedTime(&g->stats.total_ms, g->ev0, g->ev1);
    g->stats.launches = (uint32_t)(g_launches.load() - launches0);
    return CC_OK;
}
}  // namespace

extern "C" {

int cc_find_ascii(cc_graph *g, const uint8_t *kmers, uint64_t nq, int64_t *out_index, int algo) {
    if (!g || (nq && (!kmers || !out_index))) return fail(CC_ERR_ARG, "null argument");
    if (nq == 0) return CC_OK;
    const uint32_t k = g->h.k;
    return chunked_lookup(g, kmers, k, 0, nullptr, nq, out_index,
                          [&](const uint8_t *d, const uint8_t *, uint64_t m, int64_t *o, cudaStream_t st) {
                              return launch_find_seq(g, d, m * k, k, m, o, algo, st);
                          });
}

int cc_find_windows(cc_graph *g, const uint8_t *seq, uint64_t len, int64_t *out_index, int algo) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    const uint32_t k = g->h.k;
    if (len < k) return CC_OK;
    if (!seq || !out_index) return fail(CC_ERR_ARG, "null argument");
    // windows are chunked with k-1 bytes of overlap: chunk of m windows needs m + k - 1 bytes
    return chunked_lookup(g, seq, 1, k - 1, nullptr, len - k + 1, out_index,
                          [&](const uint8_t *d, const uint8_t *, uint64_t m, int64_t *o, cudaStream_t st) {
                              return launch_find_seq(g, d, m + k - 1, 1, m, o, algo, st);
                          });
}

int cc_find_packed(cc_graph *g, const uint64_t *words, const uint8_t *flags, uint64_t nq, int64_t *out_index, int algo) {
    if (!g || (nq && (!words || !out_index))) return fail(CC_ERR_ARG, "null argument");
    if (nq == 0) return CC_OK;
    return chunked_lookup(g, reinterpret_cast<const uint8_t *>(words), 8ull * g->h.s, 0, flags, nq, out_index,
                          [&](const uint8_t *d, const uint8_t *f, uint64_t m, int64_t *o, cudaStream_t st) {
                              return launch_find_packed(g, reinterpret_cast<const uint64_t *>(d), f, m, o, algo, st);
                          });
}

// The legacy per-record API in one call: TraversalEngine asks findRecord for a vertex and then for its (up to 8) neighbours
// (S/utils/traversal/TraversalEngine.java:67-252); the reference answers each from an LRU over mmap'ed pages.  Up to 64 k-mers
// take ONE kernel launch with no allocation and no staging copies (queries and answers travel through mapped pinned memory
// kept in the handle); larger batches go through cc_find_ascii and a gather of the records.
int cc_find_records(cc_graph *g, const uint8_t *kmers, uint64_t nq, int64_t *out_index, void *out_raw) {
    if (!g || (nq && (!kmers || !out_index))) return fail(CC_ERR_ARG, "null argument");
    if (nq == 0) return CC_OK;
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    const uint64_t k = g->h.k, S = g->h.record_size;
    const uint64_t in_bytes = (nq * k + 63) & ~63ull, idx_bytes = nq * 8, raw_bytes = out_raw ? nq * S : 0;
    if (nq <= 64 && in_bytes + idx_bytes + raw_bytes <= (256u << 10)) {
        if (!g->small_host) {
            const size_t cap = (256u << 10) + 4096;
            if (cudaHostAlloc(reinterpret_cast<void **>(&g->small_host), cap, cudaHostAllocMapped) != cudaSuccess ||
                cudaHostGetDevicePointer(reinterpret_cast<void **>(&g->small_dev), g->small_host, 0) != cudaSuccess) {
                cudaGetLastError();
                if (g->small_host) { cudaFreeHost(g->small_host); g->small_host = nullptr; }
                return fail(CC_ERR_CUDA, "cannot allocate mapped pinned staging for cc_find_records");
            }
            g->small_bytes = cap;
        }
        memcpy(g->small_host, kmers, nq * k);
        int64_t *d_idx = reinterpret_cast<int64_t *>(g->small_dev + in_bytes);
        uint8_t *d_raw = out_raw ? g->small_dev + in_bytes + idx_bytes : nullptr;
        if (int rc = launch_find_small(g, g->small_dev, (uint32_t)nq, d_idx, d_raw, g->stream)) return rc;
        if (int rc = sync_stream(g, g->stream)) return rc;
        memcpy(out_index, g->small_host + in_bytes, idx_bytes);
        if (out_raw) memcpy(out_raw, g->small_host + in_bytes + idx_bytes, raw_bytes);
        return CC_OK;
    }
    if (int rc = cc_find_ascii(g, kmers, nq, out_index, CC_ALGO_AUTO)) return rc;
    if (out_raw) {
        // records of the hits, one copy each (this path is for convenience, not speed)
        uint8_t *dst = static_cast<uint8_t *>(out_raw);
        for (uint64_t i = 0; i < nq; ++i) {
            if (out_index[i] < 0) { memset(dst + i * S, 0, S); continue; }
            const uint64_t local = (uint64_t)out_index[i] - g->first_index;
            if (g->host_image) memcpy(dst + i * S, g->host_image + g->h.data_offset + local * S, S);
            else CC_CUDA(cudaMemcpy(dst + i * S, g->dev_body + local * S, S, cudaMemcpyDeviceToHost));
        }
    }
    return CC_OK;
}

int cc_contains_windows(cc_graph *g, const uint8_t *seq, uint64_t len, uint8_t *out_present) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    const uint32_t k = g->h.k;
    if (len < k) return CC_OK;
    if (!seq || !out_present) return fail(CC_ERR_ARG, "null argument");
    const uint64_t nw = len - k + 1;
    std::vector<int64_t> idx(nw);
    if (int rc = cc_find_windows(g, seq, len, idx.data(), CC_ALGO_AUTO)) return rc;
    for (uint64_t i = 0; i < nw; ++i) out_present[i] = idx[i] >= 0;
    return CC_OK;
}

// ==================================================================== multi-GPU helpers
int cc_bucket_by_owner_dev(int device, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t s,
                           const uint64_t *dev_splitters, int nshards, uint64_t *dev_counts, uint64_t *dev_sorted_words,
                           uint32_t *dev_slots, void *stream) {
    if (int rc = check_device(device)) return rc;
    if (!dev_counts || (nq && (!dev_words || !dev_sorted_words || !dev_slots)) || (nshards > 1 && !dev_splitters))
        return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_bucket_by_owner(dev_words, dev_flags, nq, s, dev_splitters, nshards, dev_counts, dev_sorted_words, dev_slots,
                                  static_cast<cudaStream_t>(stream));
}

int cc_scatter_results_dev(int device, const int64_t *dev_values, const uint32_t *dev_slots, uint64_t n, int64_t *dev_out, void *stream) {
    if (int rc = check_device(device)) return rc;
    if (n && (!dev_values || !dev_slots || !dev_out)) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_scatter_results(dev_values, dev_slots, n, dev_out, static_cast<cudaStream_t>(stream));
}

// ---- routed lookups over peer memory (NVLink P2P): no collective library on the data path
int cc_route_state_bytes(uint64_t max_queries, int nshards, uint64_t *out_bytes) {
    if (!out_bytes || nshards < 1) return fail(CC_ERR_ARG, "null argument");
    *out_bytes = route_state_size(max_queries, nshards);
    return CC_OK;
}

int cc_route_queries_dev(int device, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t k,
                         const uint64_t *dev_splitters, int nshards, int my_rank, uint64_t cap, void *const *peer_inbox,
                         void *const *peer_counts_in, void *dev_route_state, uint64_t max_queries, uint64_t *dev_sent, void *stream) {
    if (int rc = check_device(device)) return rc;
    if (!peer_inbox || !peer_counts_in || !dev_sent || (nq && (!dev_words || !dev_route_state)) || (nshards > 1 && !dev_splitters))
        return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_route(dev_words, dev_flags, nq, k, dev_splitters, nshards, my_rank, cap, peer_inbox, peer_counts_in, dev_route_state,
                        max_queries, dev_sent, static_cast<cudaStream_t>(stream));
}

int cc_publish_counts_dev(int device, const uint64_t *dev_sent, int nshards, int my_rank, uint64_t cap, void *const *peer_counts_in, void *stream) {
    if (int rc = check_device(device)) return rc;
    if (!dev_sent || !peer_counts_in) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_publish_counts(dev_sent, nshards, my_rank, cap, peer_counts_in, static_cast<cudaStream_t>(stream));
}

int cc_find_routed_dev(cc_graph *g, const void *dev_inbox, const uint64_t *dev_counts_in, int world, int vsub, uint64_t cap,
                       void *dev_res, void *stream) {
    if (!g || !dev_inbox || !dev_counts_in || !dev_res) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    return launch_find_routed(g, dev_inbox, dev_counts_in, world, vsub, cap, dev_res, static_cast<cudaStream_t>(stream));
}

int cc_gather_routed_dev(int device, void *const *peer_res, const void *dev_route_state, uint64_t max_queries, uint64_t nq,
                         const uint64_t *dev_shard_first, int nshards, uint64_t cap, int64_t *dev_out, void *stream) {
    if (int rc = check_device(device)) return rc;
    if (nq && (!peer_res || !dev_route_state || !dev_shard_first || !dev_out)) return fail(CC_ERR_ARG, "null argument");
    DeviceGuard guard(device);
    return launch_gather_routed(peer_res, dev_route_state, max_queries, nq, dev_shard_first, nshards, cap, dev_out, static_cast<cudaStream_t>(stream));
}

// ==================================================================== next rows: merged view (CortexCollection / Join)
int cc_join(cc_graph *const *graphs, int ngraphs, cc_graph **out) {
    if (!graphs || !out || ngraphs < 1) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    for (int i = 0; i < ngraphs; ++i)
        if (!graphs[i]) return fail(CC_ERR_ARG, "null graph");
    const int device = graphs[0]->device;
    for (int i = 0; i < ngraphs; ++i) {
        if (graphs[i]->device != device) return fail(CC_ERR_ARG, "graphs to join must live on one device");
        if (graphs[i]->h.k != graphs[0]->h.k)      // CortexCollection.java:43-45
            return fail(CC_ERR_ARG, "Graph kmer sizes are not equal.  Expected k=%u, but found k=%u in graph %s", graphs[0]->h.k,
                        graphs[i]->h.k, graphs[i]->path.c_str());
    }
    DeviceGuard guard(device);
    for (int i = 0; i < ngraphs; ++i)
        if (int rc = ensure_index(graphs[i])) return rc;          // key columns; rejects unsorted inputs
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> g(new cc_graph(), destroy);
    g->path = "<join>";
    g->h = graphs[0]->h;
    g->h.data_offset = 0;
    if (int rc = init_handle(g.get(), device)) return rc;
    const uint32_t s = graphs[0]->h.s;
    // fold left: ((g0 U g1) U g2) ...
    void *cur_body = nullptr;          // owned intermediate (null while the running result is graphs[0] itself)
    const uint8_t *body = graphs[0]->dev_body;
    const uint64_t *keys = graphs[0]->index.keys;
    uint64_t n = graphs[0]->h.num_records;
    uint32_t c = graphs[0]->h.c;
    uint64_t *cur_keys = nullptr;
    struct Tmp { void *&b; uint64_t *&k; ~Tmp() { if (b) cudaFree(b); if (k) cudaFree(k); } } tmp{cur_body, cur_keys};
    for (int i = 1; i < ngraphs; ++i) {
        void *nb = nullptr;
        uint64_t nn = 0;
        uint64_t *nk = nullptr;        // key column of the intermediate for the next union (written by the tiled emit pass)
        if (int rc = join_pair(body, keys, n, c, graphs[i]->dev_body, graphs[i]->index.keys, graphs[i]->h.num_records, graphs[i]->h.c, s,
                               g->stream, &nb, &nn, i + 1 < ngraphs ? &nk : nullptr)) { if (nb) cudaFree(nb); if (nk) cudaFree(nk); return rc; }
        if (cur_body) cudaFree(cur_body);
        if (cur_keys) { cudaFree(cur_keys); cur_keys = nullptr; }
        cur_body = nb;
        body = static_cast<const uint8_t *>(nb);
        n = nn;
        c += graphs[i]->h.c;
        for (const ColorMeta &cm : graphs[i]->h.colors) g->h.colors.push_back(cm);
        if (i + 1 < ngraphs) {
            cur_keys = nk;
            if (!cur_keys) {           // the untiled union does not produce it: decode it from the records
                CC_CUDA(cudaMalloc(&cur_keys, std::max<uint64_t>(n * s, 2) * 8 + 64));
                if (int rc = g->scan_ws.ensure(0, 0)) return rc;
                if (int rc = launch_decode_columns(body, n, s, c, cur_keys, nullptr, nullptr, g->scan_ws, g->sm_count, g->stream)) return rc;
                CC_CUDA(cudaStreamSynchronize(g->stream));
            }
            keys = cur_keys;
        }
    }
    if (!cur_body) {                   // a single graph: copy it
        CC_CUDA(cudaMalloc(&cur_body, n * graphs[0]->h.record_size + 256));
        CC_CUDA(cudaMemcpyAsync(cur_body, body, n * graphs[0]->h.record_size, cudaMemcpyDeviceToDevice, g->stream));
        CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(cur_body) + n * graphs[0]->h.record_size, 0, 256, g->stream));
        CC_CUDA(cudaStreamSynchronize(g->stream));
    }
    g->h.c = c;
    g->h.record_size = 8ull * s + 5ull * c;
    g->h.num_records = n;
    g->dev_alloc = cur_body;
    g->dev_body = static_cast<const uint8_t *>(cur_body);
    cur_body = nullptr;
    *out = g.release();
    return CC_OK;
}

// Sort (S/commands/utils/Sort.java:19-50): records in ascending k-mer order (stable, like Arrays.sort on objects), same
// header.  Accepts the hash-ordered graphs McCortex writes; the result satisfies findRecord's sortedness requirement.
int cc_sort(cc_graph *g_in, cc_graph **out) {
    if (!g_in || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    DeviceGuard guard(g_in->device);
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> g(new cc_graph(), destroy);
    g->path = "<sort>";
    g->h = g_in->h;
    g->h.data_offset = 0;
    if (int rc = init_handle(g.get(), g_in->device)) return rc;
    const uint64_t n = g_in->h.num_records;
    const uint32_t s = g_in->h.s, S = (uint32_t)g_in->h.record_size;
    if (s > 4) return fail(CC_ERR_UNSUPPORTED, "k-mers wider than 4 words (k > 128) are not supported by sort");
    void *body = nullptr;
    CC_CUDA(cudaMalloc(&body, n * S + 256));
    g->dev_alloc = body;
    g->dev_body = static_cast<const uint8_t *>(body);
    if (n) {
        uint64_t *keys = nullptr;
        CC_CUDA(cudaMalloc(&keys, std::max<uint64_t>(n * s, 2) * 8 + 64));
        struct K { uint64_t *p; ~K() { cudaFree(p); } } kf{keys};
        if (int rc = g->scan_ws.ensure(0, 0)) return rc;
        if (int rc = launch_decode_columns(g_in->dev_body, n, s, g_in->h.c, keys, nullptr, nullptr, g->scan_ws, g->sm_count, g->stream)) return rc;
        uint32_t *perm = nullptr;
        if (int rc = sort_permutation(keys, n, s, g_in->h.k, g->stream, &perm)) return rc;
        int rc = launch_gather_records(g_in->dev_body, perm, n, S, static_cast<uint8_t *>(body), g->stream);
        cudaFreeAsync(perm, g->stream);
        if (rc) return rc;
    }
    CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(body) + n * S, 0, 256, g->stream));
    CC_CUDA(cudaStreamSynchronize(g->stream));
    *out = g.release();
    return CC_OK;
}

// ==================================================================== next rows: scan-shaped pre-filters
namespace {

// A new device-resident graph made of the records of `src` listed in `sel`, projected to the first c_out colours.
int make_selected_graph(cc_graph *src, const Header &hdr, uint32_t c_out, const uint32_t *sel, uint64_t m, const uint8_t *flags,
                        const int32_t *patch, uint32_t patch_color, cudaStream_t st, cc_graph **out) {
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> g(new cc_graph(), destroy);
    g->path = "<filter>";
    g->h = hdr;
    g->h.data_offset = 0;
    g->h.c = c_out;
    g->h.record_size = 8ull * hdr.s + 5ull * c_out;
    g->h.num_records = m;
    if (int rc = init_handle(g.get(), src->device)) return rc;
    void *body = nullptr;
    // from the stream-ordered pool: filter outputs are created and disposed per command, and a cudaMalloc / cudaFree pair of
    // half a gigabyte costs more than the whole scan
    CC_CUDA(cudaMallocAsync(&body, m * g->h.record_size + 256, g->stream));
    g->dev_alloc = body;
    g->dev_alloc_pooled = true;
    g->dev_body = static_cast<const uint8_t *>(body);
    CC_CUDA(cudaStreamSynchronize(g->stream));            // the projection below runs on the source graph's stream
    if (int rc = launch_project_records(src->dev_body, src->h.s, src->h.c, sel, m, c_out, flags, patch, patch_color,
                                        static_cast<uint8_t *>(body), src->sm_count, st)) return rc;
    CC_CUDA(cudaMemsetAsync(static_cast<uint8_t *>(body) + m * g->h.record_size, 0, 256, st));
    CC_CUDA(cudaStreamSynchronize(st));
    *out = g.release();
    return CC_OK;
}

struct StreamBuf {                      // stream-ordered scratch, freed on scope exit
    cudaStream_t st;
    std::vector<void *> p;
    explicit StreamBuf(cudaStream_t s) : st(s) {}
    ~StreamBuf() { for (void *q : p) cudaFreeAsync(q, st); }
    int alloc_bytes(void **out, uint64_t bytes) {
        void *q = nullptr;
        CC_CUDA(cudaMallocAsync(&q, bytes + 64, st));
        p.push_back(q);
        *out = q;
        return CC_OK;
    }
};

// Bit mask over the colours of g from a list that may hold -1 / out-of-range entries (a sample name that did not resolve:
// the reference's HashSet<Integer> then simply never matches, FindShared.java:45-46, CovStats.java:78-86).
std::vector<uint32_t> color_mask(const cc_graph *g, const int32_t *list, int n) {
    std::vector<uint32_t> m((g->h.c + 31) / 32 + 1, 0u);
    for (int i = 0; i < n; ++i)
        if (list[i] >= 0 && (uint32_t)list[i] < g->h.c) m[list[i] >> 5] |= 1u << (list[i] & 31);
    return m;
}

#define CC_SB_ALLOC(sb, ptr, count) (sb).alloc_bytes(reinterpret_cast<void **>(&(ptr)), (uint64_t)(count) * sizeof(*(ptr)))

}  // namespace

// FindLowCoverage (S/commands/prefilter/FindLowCoverage.java:33-66): the records whose coverage(0) is below the limit, under the
// input's header.
int cc_find_low_coverage(cc_graph *roi, int32_t min_coverage, cc_graph **out) {
    if (!roi || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    if (roi->h.c == 0) return fail(CC_ERR_ARG, "graph has no colours");
    DeviceGuard guard(roi->device);
    cudaStream_t st = roi->stream;
    const uint64_t n = roi->h.num_records;
    StreamBuf sb(st);
    int32_t *cov = nullptr; uint8_t *flags = nullptr;
    if (int rc = CC_SB_ALLOC(sb, cov, n * roi->h.c)) return rc;
    if (int rc = CC_SB_ALLOC(sb, flags, n)) return rc;
    if (int rc = roi->scan_ws.ensure(0, 0)) return rc;
    if (int rc = launch_decode_columns(roi->dev_body, n, roi->h.s, roi->h.c, nullptr, cov, nullptr, roi->scan_ws, roi->sm_count, st)) return rc;
    if (int rc = launch_lowcov_flags(cov, n, roi->h.c, min_coverage, flags, roi->sm_count, st)) return rc;
    uint32_t *sel = nullptr; uint64_t m = 0;
    if (int rc = select_flagged(flags, n, &sel, &m, st)) return rc;
    sb.p.push_back(sel);
    return make_selected_graph(roi, roi->h, roi->h.c, sel, m, nullptr, nullptr, 0, st, out);
}

// Remove (S/commands/utils/Remove.java:30-88): the merged view of the primary and the secondary graphs (CortexCollection) is
// walked; a record with coverage > 0 in any secondary colour is dropped, the others are written with the primary's colours
// under the primary's header.  K-mers that only a secondary graph holds, with coverage <= 0 there, come out as records with
// all-zero primary colours -- the reference writes them, so they are written here.
int cc_remove(cc_graph *primary, cc_graph *const *secondaries, int nsecondaries, cc_graph **out, uint64_t *out_removed) {
    if (!primary || !out || nsecondaries < 0 || (nsecondaries > 0 && !secondaries)) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    std::vector<cc_graph *> all;
    all.push_back(primary);
    for (int i = 0; i < nsecondaries; ++i) all.push_back(secondaries[i]);
    cc_graph *merged = nullptr;
    if (int rc = cc_join(all.data(), (int)all.size(), &merged)) return rc;
    std::unique_ptr<cc_graph, void (*)(cc_graph *)> mg(merged, destroy);
    DeviceGuard guard(merged->device);
    cudaStream_t st = merged->stream;
    const uint64_t n = merged->h.num_records;
    StreamBuf sb(st);
    int32_t *cov = nullptr; uint8_t *flags = nullptr;
    if (int rc = CC_SB_ALLOC(sb, cov, n * merged->h.c)) return rc;
    if (int rc = CC_SB_ALLOC(sb, flags, n)) return rc;
    if (int rc = merged->scan_ws.ensure(0, 0)) return rc;
    if (int rc = launch_decode_columns(merged->dev_body, n, merged->h.s, merged->h.c, nullptr, cov, nullptr, merged->scan_ws, merged->sm_count, st)) return rc;
    if (int rc = launch_remove_flags(cov, n, merged->h.c, primary->h.c, flags, merged->sm_count, st)) return rc;
    uint32_t *sel = nullptr; uint64_t m = 0;
    if (int rc = select_flagged(flags, n, &sel, &m, st)) return rc;
    sb.p.push_back(sel);
    if (out_removed) *out_removed = n - m;
    return make_selected_graph(merged, primary->h, primary->h.c, sel, m, nullptr, nullptr, 0, st, out);
}

// FindShared (S/commands/prefilter/FindShared.java:40-118): the ROI records whose k-mer, looked up in the pedigree graph, has
// coverage in a colour that is neither the child, a parent nor ignored.
int cc_find_shared(cc_graph *graph, cc_graph *roi, int32_t child, const int32_t *parents, int nparents, const int32_t *ignore, int nignore,
                   cc_graph **out) {
    if (!graph || !roi || !out || (nparents > 0 && !parents) || (nignore > 0 && !ignore) || nparents < 0 || nignore < 0)
        return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    if (graph->device != roi->device) return fail(CC_ERR_ARG, "graph and ROI must live on one device");
    if (graph->h.k != roi->h.k) return fail(CC_ERR_ARG, "Graph kmer sizes are not equal.  Expected k=%u, but found k=%u", graph->h.k, roi->h.k);
    DeviceGuard guard(graph->device);
    if (int rc = ensure_index(graph)) return rc;
    cudaStream_t st = graph->stream;
    const uint64_t n = roi->h.num_records;
    StreamBuf sb(st);
    uint64_t *words = nullptr; int64_t *idx = nullptr; uint8_t *flags = nullptr; uint32_t *dmask = nullptr; unsigned long long *missing = nullptr;
    if (int rc = CC_SB_ALLOC(sb, words, n * roi->h.s)) return rc;
    if (int rc = CC_SB_ALLOC(sb, idx, n)) return rc;
    if (int rc = CC_SB_ALLOC(sb, flags, n)) return rc;
    std::vector<uint32_t> mask = color_mask(graph, parents, nparents);
    for (int i = 0; i < nignore; ++i)
        if (ignore[i] >= 0 && (uint32_t)ignore[i] < graph->h.c) mask[ignore[i] >> 5] |= 1u << (ignore[i] & 31);
    if (child >= 0 && (uint32_t)child < graph->h.c) mask[child >> 5] |= 1u << (child & 31);
    // FindShared.java:64-70 dereferences the found record only inside `c != child && !parents.contains(c) && !ignore.contains(c)`:
    // with every colour excluded a ROI k-mer that is absent from the graph is not an error there, it is simply not shared
    bool any_free = false;
    for (uint32_t cc = 0; cc < graph->h.c; ++cc) any_free |= !((mask[cc >> 5] >> (cc & 31)) & 1u);
    if (int rc = CC_SB_ALLOC(sb, dmask, mask.size())) return rc;
    if (int rc = CC_SB_ALLOC(sb, missing, 1)) return rc;
    const unsigned long long none = ~0ull;
    CC_CUDA(cudaMemcpyAsync(dmask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice, st));
    CC_CUDA(cudaMemcpyAsync(missing, &none, 8, cudaMemcpyHostToDevice, st));
    if (int rc = roi->scan_ws.ensure(0, 0)) return rc;
    if (int rc = launch_decode_columns(roi->dev_body, n, roi->h.s, roi->h.c, words, nullptr, nullptr, roi->scan_ws, roi->sm_count, st)) return rc;
    if (int rc = launch_find_packed(graph, words, nullptr, n, idx, CC_ALGO_AUTO, st)) return rc;      // records hold canonical k-mers
    if (int rc = launch_shared_flags(idx, graph->dev_body, (uint32_t)graph->h.record_size, graph->h.s, graph->h.c, graph->first_index, dmask, n,
                                     flags, missing, graph->sm_count, st)) return rc;
    unsigned long long at = none;
    CC_CUDA(cudaMemcpyAsync(&at, missing, 8, cudaMemcpyDeviceToHost, st));
    CC_CUDA(cudaStreamSynchronize(st));
    if (at != none && any_free)
        return fail(CC_ERR_ARG, "ROI record %llu is not in the graph (java.lang.NullPointerException at FindShared.java:69 in the reference)", at);
    uint32_t *sel = nullptr; uint64_t m = 0;
    if (int rc = select_flagged(flags, n, &sel, &m, st)) return rc;
    sb.p.push_back(sel);
    return make_selected_graph(roi, roi->h, roi->h.c, sel, m, nullptr, nullptr, 0, st, out);
}

// RecoverExcludedKmers (S/commands/discover/recover/RecoverExcludedKmers.java:31-106).  The output header has ONE colour (the
// child's metadata, makeHeader :98-106) and CortexGraphWriter.addRecord writes header.getNumColors() colours of each record:
// the k-mer, coverage[0] and edges[0] of the pedigree record -- colour 0, whichever colour the child is.  The recovered
// coverage lands in coverage[child], so it is visible in the output only when the child is colour 0.  Reproduced as is.
int cc_recover_excluded_kmers(cc_graph *graph, cc_graph *dirty, int32_t child, cc_graph **out, uint64_t *out_recovered) {
    if (!graph || !dirty || !out) return fail(CC_ERR_ARG, "null argument");
    *out = nullptr;
    if (graph->device != dirty->device) return fail(CC_ERR_ARG, "graphs must live on one device");
    if (graph->h.k != dirty->h.k) return fail(CC_ERR_ARG, "Graph kmer sizes are not equal.  Expected k=%u, but found k=%u", graph->h.k, dirty->h.k);
    if (child < 0 || (uint32_t)child >= graph->h.c) return fail(CC_ERR_ARG, "Sample not found in pedigree graph (colour %d)", child);
    if (dirty->h.c == 0) return fail(CC_ERR_ARG, "dirty graph has no colours");
    DeviceGuard guard(graph->device);
    if (int rc = ensure_index(dirty)) return rc;
    cudaStream_t st = graph->stream;
    const uint64_t n = graph->h.num_records;
    const uint32_t c = graph->h.c, s = graph->h.s;
    StreamBuf sb(st);
    uint64_t *words = nullptr; int32_t *cov = nullptr, *patch = nullptr; int64_t *idx = nullptr;
    uint8_t *cls = nullptr, *fflags = nullptr, *flags = nullptr; unsigned long long *recovered = nullptr;
    if (int rc = CC_SB_ALLOC(sb, words, n * s)) return rc;
    if (int rc = CC_SB_ALLOC(sb, cov, n * c)) return rc;
    if (int rc = CC_SB_ALLOC(sb, patch, n)) return rc;
    if (int rc = CC_SB_ALLOC(sb, idx, n)) return rc;
    if (int rc = CC_SB_ALLOC(sb, cls, n)) return rc;
    if (int rc = CC_SB_ALLOC(sb, fflags, n)) return rc;
    if (int rc = CC_SB_ALLOC(sb, flags, n)) return rc;
    if (int rc = CC_SB_ALLOC(sb, recovered, 1)) return rc;
    CC_CUDA(cudaMemsetAsync(recovered, 0, 8, st));
    if (int rc = graph->scan_ws.ensure(0, 0)) return rc;
    if (int rc = launch_decode_columns(graph->dev_body, n, s, c, words, cov, nullptr, graph->scan_ws, graph->sm_count, st)) return rc;
    if (int rc = launch_recover_classes(cov, n, c, (uint32_t)child, cls, fflags, graph->sm_count, st)) return rc;
    if (int rc = launch_find_packed(dirty, words, fflags, n, idx, CC_ALGO_AUTO, st)) return rc;
    if (int rc = launch_recover_finalize(cls, idx, dirty->dev_body, (uint32_t)dirty->h.record_size, dirty->h.s, dirty->first_index, n, flags, patch,
                                         recovered, graph->sm_count, st)) return rc;
    uint32_t *sel = nullptr; uint64_t m = 0;
    if (int rc = select_flagged(flags, n, &sel, &m, st)) return rc;
    sb.p.push_back(sel);
    if (out_recovered) {
        unsigned long long r = 0;
        CC_CUDA(cudaMemcpyAsync(&r, recovered, 8, cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
        *out_recovered = r;
    }
    Header hdr = graph->h;
    hdr.colors.clear();
    hdr.colors.push_back(graph->h.colors.at((size_t)child));
    return make_selected_graph(graph, hdr, 1, sel, m, flags, patch, (uint32_t)child, st, out);
}

// CovStats (S/commands/utils/CovStats.java:33-72): child coverage -> sum over the records (present in the child, in a parent and in
// another sample) of numberOfParents + numberOfChildren, ascending coverage.  Counts are Java ints (they wrap).
int cc_cov_stats(cc_graph *g, int32_t child, const int32_t *parents, int nparents, int32_t *out_cov, int32_t *out_count, uint64_t cap,
                 uint64_t *out_n) {
    if (!g || !out_n || (nparents > 0 && !parents) || nparents < 0 || (cap && (!out_cov || !out_count))) return fail(CC_ERR_ARG, "null argument");
    *out_n = 0;
    // getCoverage(childColor) with childColor = -1 / out of range is an ArrayIndexOutOfBoundsException (CovStats.java:49)
    if (child < 0 || (uint32_t)child >= g->h.c) return fail(CC_ERR_ARG, "child colour %d out of range (graph has %u colours)", child, g->h.c);
    DeviceGuard guard(g->device);
    cudaStream_t st = g->stream;
    const uint64_t n = g->h.num_records;
    StreamBuf sb(st);
    uint32_t *dmask = nullptr;
    std::vector<uint32_t> mask = color_mask(g, parents, nparents);
    if (int rc = CC_SB_ALLOC(sb, dmask, mask.size())) return rc;
    CC_CUDA(cudaMemcpyAsync(dmask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice, st));
    std::vector<int32_t> hc;
    std::vector<long long> hn;
    bool fell_back = !options().covstats_fused;
    if (!fell_back)
        if (int rc = cov_stats_fused(g->dev_body, n, g->h.s, g->h.c, child, dmask, g->sm_count, st, hc, hn, &fell_back)) return rc;
    if (fell_back) {          // more huge coverages than the overflow list holds: decode the coverage matrix, sort, reduce by key
        int32_t *cov = nullptr;
        if (int rc = CC_SB_ALLOC(sb, cov, n * g->h.c)) return rc;
        if (int rc = g->scan_ws.ensure(0, 0)) return rc;
        if (int rc = launch_decode_columns(g->dev_body, n, g->h.s, g->h.c, nullptr, cov, nullptr, g->scan_ws, g->sm_count, st)) return rc;
        if (int rc = cov_stats(cov, n, g->h.c, child, dmask, g->sm_count, st, hc, hn)) return rc;
    }
    *out_n = hc.size();
    for (uint64_t i = 0; i < hc.size() && i < cap; ++i) {
        out_cov[i] = hc[i];
        out_count[i] = (int32_t)(uint32_t)(unsigned long long)hn[i];      // Java int arithmetic
    }
    return CC_OK;
}

// CortexGraphWriter.initialize :31-104 + one write of the whole body.  total_sequence is emitted the way the reference
// does after ITS round trip: it reads the field big-endian (BinaryFile.readUnsignedLong :34-38) and writes it
// little-endian (:60-63), i.e. byte-reversed with respect to the input file.
int cc_write_graph(const cc_graph *g, const char *path) {
    if (!g || !path) return fail(CC_ERR_ARG, "null argument");
    std::vector<uint8_t> hdr;
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; ++i) hdr.push_back((uint8_t)(v >> (8 * i))); };
    auto raw = [&](const void *p, size_t n) { hdr.insert(hdr.end(), (const uint8_t *)p, (const uint8_t *)p + n); };
    static const uint8_t err[16] = {0, 0xd8, 0xa3, 0x70, 0x3d, 0x0a, 0xd7, 0xa3, 0xf8, 0x3f, 0, 0, 0, 0, 0, 0};
    raw("CORTEX", 6);
    u32(6); u32(g->h.k); u32(g->h.s); u32(g->h.c);
    for (const ColorMeta &c : g->h.colors) u32(c.info.mean_read_length);
    for (const ColorMeta &c : g->h.colors) { const uint64_t v = c.info.total_sequence; for (int i = 7; i >= 0; --i) hdr.push_back((uint8_t)(v >> (8 * i))); }
    for (const ColorMeta &c : g->h.colors) { u32((uint32_t)c.sample_name.size()); raw(c.sample_name.data(), c.sample_name.size()); }
    for (size_t i = 0; i < g->h.colors.size(); ++i) raw(err, 16);
    for (const ColorMeta &c : g->h.colors) {
        hdr.push_back(c.info.tip_clipping ? 1 : 0);
        hdr.push_back(c.info.low_covg_supernodes_removed ? 1 : 0);
        hdr.push_back(c.info.low_covg_kmers_removed ? 1 : 0);
        hdr.push_back(c.info.cleaned_against_graph ? 1 : 0);
        u32(c.info.low_cov_supernodes_threshold);
        u32(c.info.low_cov_kmer_threshold);
        u32((uint32_t)c.graph_name.size());
        raw(c.graph_name.data(), c.graph_name.size());
    }
    raw("CORTEX", 6);
    FILE *f = fopen(path, "wb");
    if (!f) return fail(CC_ERR_IO, "Unable to open file '%s'", path);
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    const uint64_t bytes = g->h.num_records * g->h.record_size;
    std::vector<uint8_t> buf(std::min<uint64_t>(bytes, 256ull << 20));
    DeviceGuard guard(g->device);
    for (uint64_t off = 0; ok && off < bytes; off += buf.size()) {
        const uint64_t nb = std::min<uint64_t>(buf.size(), bytes - off);
        if (cudaMemcpy(buf.data(), g->dev_body + off, nb, cudaMemcpyDeviceToHost) != cudaSuccess) { fclose(f); return cuda_fail(cudaGetLastError(), "cudaMemcpy", __FILE__, __LINE__); }
        ok = fwrite(buf.data(), 1, nb, f) == nb;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(CC_ERR_IO, "Unable to write record to file '%s'", path);
    return CC_OK;
}

// ==================================================================== instrumentation
int cc_last_stats(const cc_graph *g, cc_stats *out) {
    if (!g || !out) return fail(CC_ERR_ARG, "null argument");
    *out = g->stats;
    return CC_OK;
}

int cc_device_body(const cc_graph *g, const void **dev_body, uint64_t *bytes) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    if (dev_body) *dev_body = g->dev_body;
    if (bytes) *bytes = g->h.num_records * g->h.record_size;
    return CC_OK;
}

int cc_device_keys(cc_graph *g, const uint64_t **dev_keys, uint64_t *n) {
    if (!g) return fail(CC_ERR_ARG, "null graph");
    DeviceGuard guard(g->device);
    if (int rc = ensure_index(g)) return rc;
    if (dev_keys) *dev_keys = g->index.keys;
    if (n) *n = g->h.num_records;
    return CC_OK;
}

}  // extern "C"
