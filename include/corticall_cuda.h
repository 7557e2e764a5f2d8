/*
 * corticall_cuda.h -- C ABI of libcorticall_cuda, the B200 (sm_100a) implementation of Corticall's
 * data-parallel k-mer hot path.  This is the drop-in boundary: a JNI shim (csrc/jni_shim.cpp, see
 * INTEGRATION.md) binds exactly these entry points behind the reference's Java classes.
 *
 * Reference interfaces replaced (S/ = public/java/src/uk/ac/ox/well/cortexjdk/ in mcveanlab/Corticall):
 *   S/utils/io/graph/DeBruijnGraph.java:16-53           the graph interface CortexGraph implements
 *   S/utils/io/graph/cortex/CortexGraph.java:40-48      constructors              -> cc_open / cc_open_memory
 *   S/utils/io/graph/cortex/CortexGraph.java:66-168     loadCortexGraph (header)  -> cc_header / cc_color_*
 *   S/utils/io/graph/cortex/CortexGraph.java:183-237    getRecord / getNextRecord -> cc_get_records / cc_decode_records
 *   S/utils/io/graph/cortex/CortexGraph.java:272-321    findRecord (all overloads)-> cc_find_ascii / cc_find_windows / cc_find_packed
 *   S/utils/io/graph/cortex/CortexGraph.java:335-366    getColorForSampleName(s)  -> cc_color_for_sample_name
 *   S/commands/discover/roi/FindROIs.java:31-105        the novel-k-mer step      -> cc_find_novel / cc_write_roi_file
 *   S/utils/io/graph/cortex/CortexGraphWriter.java:31-138  output layout of that step
 *   S/utils/sequence/SequenceUtils.java:206-225, S/utils/io/graph/cortex/CortexRecord.java:313-334,
 *   S/utils/kmer/CanonicalKmer.java:13-37, S/utils/kmer/CortexBinaryKmer.java:15-17
 *                                                        canonicalise + 2-bit pack -> cc_pack_canonical
 *   S/commands/discover/call/Call.java:2348-2381,2425-2451  loadRois / loadChildWalk / getRegions
 *                                                        -> cc_find_windows (child walk), cc_contains_windows (ROI membership)
 *
 * Conventions
 *   - every function returns a cc_status (0 = OK); cc_last_error() gives a thread-local message that
 *     carries the text of the CortexJDKException the reference would throw;
 *   - plain pointers and sizes only; all buffers are caller-owned; "_dev" variants take DEVICE pointers
 *     and a cudaStream_t (as void*) and are asynchronous on that stream; the others take HOST pointers
 *     and are synchronous (host<->device copies are inside the call);
 *   - k-mer words: `s = ceil(k/32)` uint64 per k-mer, NATIVE byte order, word 0 most significant, bases
 *     right-aligned, A=0 C=1 G=2 T=3 -- i.e. exactly the on-disk words read little-endian.  The Java
 *     long[] of CortexRecord is Long.reverseBytes() of each word (CortexGraph.java:208-209);
 *   - coverage is returned as int32 (Java int: values >= 2^31 wrap negative, BinaryUtils.java:6-17);
 *   - record index results are int64, -1 = the reference's `null`;
 *   - one caller thread per cc_graph, and all "_dev" calls on one handle go to ONE stream (the handle's scan workspace --
 *     ticket counter, look-back epochs -- is not shared between streams); there is NO CPU fallback: without a CUDA device every compute
 *     entry point fails with CC_ERR_CUDA.
 */
#ifndef CORTICALL_CUDA_H
#define CORTICALL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CC_API __attribute__((visibility("default")))
#else
#define CC_API
#endif

typedef enum {
    CC_OK = 0,
    CC_ERR_NOT_CORTEX = 1,   /* CortexGraph.java:74-76   "does not appear to be a Cortex graph"      */
    CC_ERR_BAD_VERSION = 2,  /* CortexGraph.java:82-84   "is not a version 6 Cortex graph"           */
    CC_ERR_BAD_TRAILER = 3,  /* CortexGraph.java:140-142 "didn't see a proper header terminator"     */
    CC_ERR_IO = 4,           /* CortexGraph.java:163-167 file not found / parse error                */
    CC_ERR_UNSORTED = 5,     /* CortexGraph.java:295-301 "Records are not sorted"                    */
    CC_ERR_RANGE = 6,        /* CortexGraph.java:173-175 record index out of range                   */
    CC_ERR_CUDA = 7,         /* no device / CUDA runtime failure                                     */
    CC_ERR_NCCL = 8,         /* reserved for the multi-GPU exchange                                  */
    CC_ERR_ARG = 9,          /* bad argument (colour out of range = Java ArrayIndexOutOfBounds)      */
    CC_ERR_UNSUPPORTED = 10  /* shape outside what the kernels cover                 */
} cc_status;

typedef struct cc_graph cc_graph;   /* opaque: host mapping, device buffers, index, stream */

/* Per-colour header block (ctx_spec.md tables 1-3; CortexColor.java). */
typedef struct {
    uint32_t mean_read_length;
    uint64_t total_sequence;             /* little-endian per the spec (the Java reader mis-parses it) */
    uint8_t  tip_clipping, low_covg_supernodes_removed, low_covg_kmers_removed, cleaned_against_graph;
    uint32_t low_cov_supernodes_threshold, low_cov_kmer_threshold;
} cc_color_info;

/* Scan / lookup statistics of the last call on a handle (device time from CUDA events). */
typedef struct {
    float    kernel_ms;        /* device time of the kernels of the last call                 */
    float    total_ms;         /* device time of the whole call incl. copies (host variants)  */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t launches;         /* kernels launched by the last call                           */
} cc_stats;

/* ---------------------------------------------------------------- library */
CC_API const char *cc_last_error(void);                 /* thread-local, never NULL */
CC_API const char *cc_version(void);
CC_API int cc_device_count(int *out);                   /* CC_ERR_CUDA when no driver / device */

/* ---------------------------------------------------------------- lifecycle (CortexGraph ctors, :40-48,:66-168) */
/* Parse header, map the file, upload the record body to `device`. */
CC_API int cc_open(const char *path, int device, cc_graph **out);
/* Same from an in-memory image of a whole .ctx file (header + body). */
CC_API int cc_open_memory(const void *file_image, uint64_t size, int device, cc_graph **out);
/* Wrap a DEVICE-resident record array (on-disk record layout, n records of 8s+5c bytes, caller keeps it
 * alive; must lie in an allocation readable up to the next 16-byte boundary).  Used for shards of a
 * k-mer-range-partitioned graph and by the benchmark.  `first_index` is added to every record index the
 * handle reports (the shard's offset in the global array). */
CC_API int cc_open_device(const void *dev_body, uint32_t k, uint32_t s, uint32_t c, uint64_t n,
                          uint64_t first_index, int device, cc_graph **out);
/* Frees device memory and the mapping.  NOT CortexGraph.close(): the Java close() only closes the
 * RandomAccessFile and the graph stays usable (CortexGraph.java:253-255,264-270). */
CC_API void cc_dispose(cc_graph *g);

/* ---------------------------------------------------------------- header / colours */
CC_API int cc_header(const cc_graph *g, uint32_t *version, uint32_t *kmer_size, uint32_t *kmer_bits,
                     uint32_t *num_colors, uint64_t *num_records, uint64_t *data_offset, uint64_t *record_size);
CC_API int cc_color_name(const cc_graph *g, uint32_t color, char *buf, size_t cap);          /* getSampleName */
CC_API int cc_color_graph_name(const cc_graph *g, uint32_t color, char *buf, size_t cap);    /* getCleanedAgainstGraphName */
CC_API int cc_color_info_get(const cc_graph *g, uint32_t color, cc_color_info *out);
/* CortexGraph.getColorForSampleName :335-354: case-insensitive name, else integer literal; -1 unless exactly one. */
CC_API int cc_color_for_sample_name(const cc_graph *g, const char *name, int32_t *out_color);

/* ---------------------------------------------------------------- K1: record access / streaming decode */
/* Raw on-disk bytes of records [first, first+count) -> count*S bytes (getRecord(i) for the Java side). */
CC_API int cc_get_records(const cc_graph *g, uint64_t first, uint64_t count, void *out_raw);
/* Device decode of records [first, first+count) into columns (any of the outputs may be NULL):
 * words[count*s] (native), coverage[count*c] (int32), edges[count*c]. */
CC_API int cc_decode_records(cc_graph *g, uint64_t first, uint64_t count,
                             uint64_t *out_words, int32_t *out_coverage, uint8_t *out_edges);
CC_API int cc_decode_records_dev(cc_graph *g, uint64_t first, uint64_t count,
                                 uint64_t *dev_words, int32_t *dev_coverage, uint8_t *dev_edges, void *stream);

/* ---------------------------------------------------------------- K1+K2: the novel-k-mer step (FindROIs) */
/* FindROIs.isNovel :72-82 over every record, in file order: coverage[child] > 0 (signed) and
 * coverage[p] == 0 for every listed parent.  Output records have the 1-colour layout
 * CortexGraphWriter.addRecord :106-138 emits: s words verbatim, child coverage u32 LE, child edge byte
 * (8s+5 bytes each).  *out_count receives the TOTAL number of novel records even when it exceeds cap
 * (only the first cap are stored).  out_index (optional) receives each one's input record index. */
CC_API int cc_find_novel(cc_graph *g, int32_t child, const int32_t *parents, int nparents,
                         void *out_records, uint64_t *out_index, uint64_t cap, uint64_t *out_count);
CC_API int cc_find_novel_dev(cc_graph *g, int32_t child, const int32_t *parents, int nparents,
                             void *dev_out_records, uint64_t *dev_out_index, uint64_t cap,
                             uint64_t *dev_out_count, void *stream);
/* Same scan over a HOST-resident record array streamed through the device in chunks (copies overlap the
 * kernel); nothing stays resident.  This is the end-to-end form of FindROIs for a graph that lives on
 * disk / in page cache.  host_body should be page-locked for full PCIe speed. */
CC_API int cc_find_novel_host(int device, const void *host_body, uint32_t k, uint32_t s, uint32_t c, uint64_t n,
                              int32_t child, const int32_t *parents, int nparents,
                              void *out_records, uint64_t *out_index, uint64_t cap, uint64_t *out_count,
                              cc_stats *stats);
/* FindROIs.execute :31-70 end to end: scan + write `out_path` (header of FindROIs.makeCortexHeader :85-105). */
CC_API int cc_write_roi_file(cc_graph *g, int32_t child, const int32_t *parents, int nparents,
                             const char *out_path, uint64_t *out_count);

/* ---------------------------------------------------------------- K3: canonicalise + 2-bit pack */
/* Every k-window of seq (nw = len-k+1, 0 if len<k): canonical orientation by the reference's ASCII rule
 * (SequenceUtils.java:206-225), packed like CortexRecord.encodeBinaryKmer :313-334.
 * flags[i]: bit0 = reverse complement chosen, bit1 = window holds a byte outside ACGTacgt (Java throws;
 * words are 0), bit2 = window holds lowercase (packs, but can never equal a record under findRecord). */
CC_API int cc_pack_canonical(int device, const uint8_t *seq, uint64_t len, uint32_t k,
                             uint64_t *out_words, uint8_t *out_flags);
CC_API int cc_pack_canonical_dev(int device, const uint8_t *dev_seq, uint64_t len, uint32_t k,
                                 uint64_t *dev_words, uint8_t *dev_flags, void *stream);
/* nq independent k-byte k-mers, row-major (no shared windows). */
CC_API int cc_pack_kmers_dev(int device, const uint8_t *dev_kmers, uint64_t nq, uint32_t k,
                             uint64_t *dev_words, uint8_t *dev_flags, void *stream);

/* ---------------------------------------------------------------- K4: batched lookups (findRecord) */
/* algo: 0 = auto (the line index: one 64-byte bucket line per lookup), 1 = plain binary search over the key column,
 *       2 = sorted-merge: a batch in ascending key order (the records of another graph, or any batch after the radix sort this mode
 *           runs when it finds the batch unsorted) is cut into tiles whose window of the key column is found by one binary search per
 *           tile and searched through shared memory; queries and results stream through once.  Synchronises the stream once. */
enum { CC_ALGO_AUTO = 0, CC_ALGO_BSEARCH = 1, CC_ALGO_MERGE = 2 };
/* Build (or rebuild) the lookup index: key column + bucket lines (an order-preserving table of 64-byte lines over the
 * array's own key range); validates ascending order (CC_ERR_UNSORTED).  Called lazily by the first
 * lookup.  index_bits = log2 of the number of key-range bins (0 picks a default, at most 13). */
CC_API int cc_build_index(cc_graph *g, int index_bits);
/* nq k-byte ASCII queries, row-major: canonicalise, search; out_index[i] = record index or -1. */
CC_API int cc_find_ascii(cc_graph *g, const uint8_t *kmers, uint64_t nq, int64_t *out_index, int algo);
CC_API int cc_find_ascii_dev(cc_graph *g, const uint8_t *dev_kmers, uint64_t nq, int64_t *dev_index, int algo, void *stream);
/* Every k-window of seq (Call.loadChildWalk :2358-2381): out_index[i] for window i. */
CC_API int cc_find_windows(cc_graph *g, const uint8_t *seq, uint64_t len, int64_t *out_index, int algo);
CC_API int cc_find_windows_dev(cc_graph *g, const uint8_t *dev_seq, uint64_t len, int64_t *dev_index, int algo, void *stream);
/* Canonical packed queries (nq*s native words).  flags (optional, from cc_pack_canonical): queries with
 * bit1 or bit2 set miss. */
CC_API int cc_find_packed(cc_graph *g, const uint64_t *words, const uint8_t *flags, uint64_t nq, int64_t *out_index, int algo);
CC_API int cc_find_packed_dev(cc_graph *g, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq,
                              int64_t *dev_index, int algo, void *stream);
/* The legacy per-record findRecord for a handful of k-mers in one call -- a vertex and its neighbours, as TraversalEngine asks
 * for them (S/utils/traversal/TraversalEngine.java:67-252): out_index[i] = record index or -1; out_raw (optional) receives the
 * record_size bytes of every hit's record (zeros for a miss), so the caller needs no second call to decode it.  Up to 64
 * k-mers cost one kernel launch, no allocation and no staging copies. */
CC_API int cc_find_records(cc_graph *g, const uint8_t *kmers, uint64_t nq, int64_t *out_index, void *out_raw);
/* ROI membership (`rois.contains(new CanonicalKmer(window))`, Call.java:191-197,2425-2451): 1/0 per window. */
CC_API int cc_contains_windows(cc_graph *g, const uint8_t *seq, uint64_t len, uint8_t *out_present);

/* ---------------------------------------------------------------- multi-GPU helpers (k-mer-range shards) */
/* Owner shard of each canonical packed query given nshards-1 splitter keys (first key of shards 1..):
 * counts per owner, and queries/slots grouped by owner (stable).  Used before the all-to-all. */
CC_API int cc_bucket_by_owner_dev(int device, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t s,
                                  const uint64_t *dev_splitters, int nshards,
                                  uint64_t *dev_counts /* nshards */, uint64_t *dev_sorted_words, uint32_t *dev_slots,
                                  void *stream);
/* dev_out[slots[i]] = values[i] (the return leg after the reverse all-to-all). */
CC_API int cc_scatter_results_dev(int device, const int64_t *dev_values, const uint32_t *dev_slots, uint64_t n,
                                  int64_t *dev_out, void *stream);

/* Routed lookups over PEER MEMORY (NVLink P2P; buffers are symmetric allocations mapped into every rank, e.g. with
 * torch.distributed._symmetric_memory or cudaIpc): each leg is one kernel fused with its transfer, no collective
 * library on the data path.  Keys travel in a compact wire format (kw = ceil(2k/32) 32-bit words, most significant
 * first: 12 bytes at k = 47); results are 4-byte indices local to the owner's shard (0xffffffff = miss) that stay on the
 * owner and are pulled (P2P reads) and rebased by the origin.
 * Virtual shards: every rank's shard may be cut into `vsub` contiguous sub-ranges; the route and gather legs then see
 * nshards = world * vsub owners (owner v lives on rank v / vsub) and the search leg walks its vsub sub-ranges one after
 * another, so that the slice of the key column in use fits L2 (the partitioned form of the lookup for large batches;
 * world = 1 gives the single-GPU partitioned lookup).  vsub = 1 is the plain sharded lookup.
 * Layout, identical on every rank (cap = capacity of one (source, owner) segment):
 *   inbox     uint32 [vsub][world][cap][kw] on the OWNER : keys routed to it, segment = (sub-range, source rank)
 *   counts_in uint64 [vsub][world]          on the OWNER : keys in each segment
 *   res       uint32 [vsub][world][cap]     on the OWNER : results, same segments and positions as the inbox
 *   route_state (cc_route_state_bytes(max_queries, nshards) bytes), sent uint64 [nshards]   local to the origin
 *   shard_first uint64 [nshards]            first global record index of the RANK SHARD owner v belongs to
 *   splitters   uint64 [nshards - 1][s]     first key of owners 1..nshards-1
 * Order of use per batch on one stream: cc_route_queries_dev -> cross-rank barrier -> cc_find_routed_dev -> barrier ->
 * cc_gather_routed_dev (writes every out[i], -1 for misses and flagged queries).  peer_inbox / peer_counts_in are HOST
 * arrays of nshards DEVICE pointers (entry v = the [world][cap][kw] / [world] block of owner v on its rank); peer_res is a
 * HOST array of nshards DEVICE pointers (entry v = the [cap] result segment (owner v, source = this rank) on owner v's rank).  With cap >= the batch size no segment can
 * overflow; with a smaller cap the caller must check sent[v] <= cap after the batch (keys beyond cap are dropped and
 * their queries report -1).  Segment space is reserved in multiples of 4 keys per tile and owner (runs leave as 16-byte aligned bulk
 * copies); the pad slots hold zero keys, are counted in sent[] / counts_in and are never read back. */
CC_API int cc_route_state_bytes(uint64_t max_queries, int nshards, uint64_t *out_bytes);
CC_API int cc_route_queries_dev(int device, const uint64_t *dev_words, const uint8_t *dev_flags, uint64_t nq, uint32_t k,
                                const uint64_t *dev_splitters, int nshards, int my_rank, uint64_t cap,
                                void *const *peer_inbox, void *const *peer_counts_in,
                                void *dev_route_state, uint64_t max_queries, uint64_t *dev_sent, void *stream);
/* Staged form of the route leg (the copy engines move the keys instead of the SMs): cc_route_queries_dev is pointed at LOCAL
 * staging (peer_inbox[v] = stage + (v - my_rank) * cap * kw * 4, peer_counts_in[v] = local_counts + (v - my_rank)), the caller
 * copies stage segment v into owner v's inbox segment (cudaMemcpyAsync over NVLink) and this call publishes the counts:
 * counts_in[my_rank] on owner v := min(dev_sent[v], cap). */
CC_API int cc_publish_counts_dev(int device, const uint64_t *dev_sent, int nshards, int my_rank, uint64_t cap,
                                 void *const *peer_counts_in, void *stream);
CC_API int cc_find_routed_dev(cc_graph *g, const void *dev_inbox, const uint64_t *dev_counts_in, int world, int vsub,
                              uint64_t cap, void *dev_res, void *stream);
CC_API int cc_gather_routed_dev(int device, void *const *peer_res, const void *dev_route_state, uint64_t max_queries, uint64_t nq,
                                const uint64_t *dev_shard_first, int nshards, uint64_t cap, int64_t *dev_out, void *stream);

/* ---------------------------------------------------------------- one graph over several GPUs of this process */
/* The reference constructs ONE CortexGraph per file in one JVM (S/utils/arguments/ArgumentHandler.java:271-274); a
 * cc_sharded is that graph with its sorted record array cut into k-mer-range shards (contiguous record slices), one per
 * entry of `devices` (SURVEY.md Appendix C cc_open_sharded).  One host thread drives all devices: buffers are plain
 * cudaMalloc memory opened to the other devices with cudaDeviceEnablePeerAccess, the legs of a lookup batch (route ->
 * search -> gather, the kernels of the routed lookups above) are ordered across devices with CUDA events, and no
 * collective library is involved.  A device id may be listed more than once (several shards on one GPU; used by the tests).
 * Results are identical to the single-GPU entry points on the same file: global record indices / -1, and the novel
 * records of all shards concatenated in shard order = file order (FindROIs.java:52-64). */
typedef struct cc_sharded cc_sharded;
typedef struct {
    float    route_ms, search_ms, gather_ms;   /* device time of the legs of the first chunk of the last call, max over devices */
    float    chunk_ms;                         /* route start -> gather end of that chunk, max over devices                    */
    uint64_t h2d_bytes, d2h_bytes;             /* host-buffer entry points                                                    */
    uint32_t launches;                         /* kernels launched by the last call                                           */
    uint32_t overflow_retries;                 /* chunks re-run in pieces because a (source, owner) segment overflowed        */
} cc_sharded_stats;
CC_API int cc_open_sharded(const char *path, const int *devices, int ndev, cc_sharded **out);
CC_API int cc_open_sharded_memory(const void *file_image, uint64_t size, const int *devices, int ndev, cc_sharded **out);
/* Placement of the record array over the devices.  CC_PLACE_RANGE (what cc_open_sharded does): contiguous k-mer ranges, one per
 * device -- for graphs larger than one GPU (BASELINE configs[3]); lookups are routed to the owning device.  CC_PLACE_REPLICATE:
 * every device holds the whole array and its own index; a batch is split evenly and every device answers its share from its own
 * copy -- no exchange, lookups scale with the device count; the scan still gives every device 1/ndev of the records.
 * CC_PLACE_AUTO: replicas when a lookup-ready copy takes at most a quarter of every device's free memory, ranges otherwise.
 * Results are identical under every placement (the reference has one CortexGraph per file: ArgumentHandler.java:271-274). */
enum { CC_PLACE_RANGE = 0, CC_PLACE_REPLICATE = 1, CC_PLACE_AUTO = 2 };
CC_API int cc_open_sharded_placed(const char *path, const int *devices, int ndev, int placement, cc_sharded **out);
CC_API int cc_open_sharded_memory_placed(const void *file_image, uint64_t size, const int *devices, int ndev, int placement, cc_sharded **out);
CC_API int cc_sharded_placement(const cc_sharded *sh, int *placement);   /* CC_PLACE_RANGE or CC_PLACE_REPLICATE: what the handle uses */
/* Wrap slices that are already resident: dev_bodies[r] = counts[r] records (on-disk layout) on devices[r], ascending across r. */
CC_API int cc_open_sharded_device(const void *const *dev_bodies, const uint64_t *counts, uint32_t k, uint32_t s, uint32_t c,
                                  const int *devices, int ndev, cc_sharded **out);
CC_API void cc_dispose_sharded(cc_sharded *sh);
CC_API int cc_sharded_info(const cc_sharded *sh, int *ndev, uint64_t *num_records, uint32_t *kmer_size, uint32_t *num_colors);
/* The handle of one shard (owned by sh; header / colour queries, cc_decode_records ...), its device and first record index. */
CC_API int cc_sharded_shard(const cc_sharded *sh, int rank, cc_graph **shard, int *device, uint64_t *first_index);
CC_API int cc_sharded_last_stats(const cc_sharded *sh, cc_sharded_stats *out);
/* findRecord batches from HOST buffers: the batch is split evenly over the devices, copied, (packed,) routed, searched,
 * gathered and copied back; out_index[i] = global record index or -1, exactly as cc_find_packed / cc_find_ascii /
 * cc_find_windows answer on one device. */
CC_API int cc_find_packed_sharded(cc_sharded *sh, const uint64_t *words, const uint8_t *flags, uint64_t nq, int64_t *out_index);
CC_API int cc_find_ascii_sharded(cc_sharded *sh, const uint8_t *kmers, uint64_t nq, int64_t *out_index);
CC_API int cc_find_windows_sharded(cc_sharded *sh, const uint8_t *seq, uint64_t len, int64_t *out_index);
/* The same with the queries already on the devices: dev_words[r] / dev_flags[r] / dev_out[r] live on devices[r], nq[r] queries
 * each (dev_flags or any of its entries may be NULL).  Synchronous; cc_sharded_last_stats has the device times. */
CC_API int cc_find_packed_sharded_dev(cc_sharded *sh, const uint64_t *const *dev_words, const uint8_t *const *dev_flags,
                                      const uint64_t *nq, int64_t *const *dev_out);
/* FindROIs over all shards (arguments as cc_find_novel): every device scans its slice, the per-shard counts become
 * exclusive offsets and the novel records land in out_records at those offsets -- one globally sorted list. */
CC_API int cc_find_novel_sharded(cc_sharded *sh, int32_t child, const int32_t *parents, int nparents,
                                 void *out_records, uint64_t *out_index, uint64_t cap, uint64_t *out_count);
CC_API int cc_write_roi_file_sharded(cc_sharded *sh, int32_t child, const int32_t *parents, int nparents,
                                     const char *out_path, uint64_t *out_count);

/* ---------------------------------------------------------------- next rows: merged view of several graphs */
/* CortexCollection / Join (S/utils/io/graph/cortex/CortexCollection.java:34-62,245-293, S/commands/utils/Join.java:23-57):
 * the sorted union of the graphs' k-mers; colours concatenated in argument order; a k-mer absent from a graph has
 * coverage 0 and no edges in that graph's colours.  All graphs on one device, same k.  Returns a new device-resident
 * graph (dispose with cc_dispose). */
CC_API int cc_join(cc_graph *const *graphs, int ngraphs, cc_graph **out);
/* Remove (S/commands/utils/Remove.java:30-88): the records of the merged view of `primary` and the secondary graphs that
 * have no coverage > 0 in any secondary colour, written with the primary's colours under the primary's header (a new
 * device-resident graph).  *out_removed (optional) = merged records dropped. */
CC_API int cc_remove(cc_graph *primary, cc_graph *const *secondaries, int nsecondaries, cc_graph **out, uint64_t *out_removed);
/* Sort (S/commands/utils/Sort.java:19-50): a new graph with the records in ascending k-mer order (stable), same header. */
CC_API int cc_sort(cc_graph *g, cc_graph **out);
/* CortexGraphWriter (S/utils/io/graph/cortex/CortexGraphWriter.java:31-138): header from the colours, then every record. */
CC_API int cc_write_graph(const cc_graph *g, const char *path);

/* ---------------------------------------------------------------- next rows: scan-shaped pre-filters / recovery */
/* Each returns a new device-resident graph (cc_dispose it; cc_write_graph writes what the reference command writes).
 * FindLowCoverage (S/commands/prefilter/FindLowCoverage.java:33-66): the records with coverage(0) < min_coverage, input header. */
CC_API int cc_find_low_coverage(cc_graph *roi, int32_t min_coverage, cc_graph **out);
/* FindShared (S/commands/prefilter/FindShared.java:40-118): ROI records whose k-mer has coverage > 0 in a colour of `graph` that is
 * neither child, parent nor ignored.  Colour lists may hold -1 (unresolved sample names never match).  A ROI k-mer absent
 * from `graph` is an error (NullPointerException in the reference). */
CC_API int cc_find_shared(cc_graph *graph, cc_graph *roi, int32_t child, const int32_t *parents, int nparents,
                          const int32_t *ignore, int nignore, cc_graph **out);
/* RecoverExcludedKmers (S/commands/discover/recover/RecoverExcludedKmers.java:31-106): records of `graph` with child coverage > 0,
 * plus those with coverage elsewhere whose k-mer the `dirty` graph holds with coverage(0) > 0 (their child coverage becomes the
 * dirty one).  One-colour output: k-mer, coverage[0], edges[0] of the pedigree record under the child's colour header, as the
 * reference's writer emits it. */
CC_API int cc_recover_excluded_kmers(cc_graph *graph, cc_graph *dirty, int32_t child, cc_graph **out, uint64_t *out_recovered);
/* CovStats (S/commands/utils/CovStats.java:33-72): rows (child coverage, count) in ascending coverage; *out_n = number of rows
 * (only the first cap are stored).  Counts wrap like Java ints. */
CC_API int cc_cov_stats(cc_graph *g, int32_t child, const int32_t *parents, int nparents,
                        int32_t *out_cov, int32_t *out_count, uint64_t cap, uint64_t *out_n);

/* ---------------------------------------------------------------- instrumentation */
CC_API int cc_last_stats(const cc_graph *g, cc_stats *out);
/* Total kernels this library has launched in this process (bench.py's gpu_launches). */
CC_API uint64_t cc_launch_count(void);
/* Device pointer / size of the resident body and key column (benchmarks, tests). */
CC_API int cc_device_body(const cc_graph *g, const void **dev_body, uint64_t *bytes);
CC_API int cc_device_keys(cc_graph *g, const uint64_t **dev_keys, uint64_t *n);
/* Process-wide tuning knobs (benchmarks and sweeps; the defaults are the measured optima):
 *   scan:    "scan_stages", "scan_tile_bytes", "scan_ctas_per_sm", "scan_chunk_tiles", "scan_stage_buf_bytes", "scan_fast", "host_chunk_mb"
 *   index:   "index_bits" (log2 of the number of bins, 0 = auto), "index_fill_pct" (average fill of a bucket line, default 50)
 *   lookups: "lookup_l2_hints" (line loads: -1 auto, 0 plain, 1 evict-first in L2, 2 evict-normal, 3 evict-last), "find_bins_smem" (1 = bin table staged in shared memory),
 *            "rows_fused" (ASCII lists: 1 = pack + search in one kernel), "rows_warp" (pack rows: 2 / 3 = warp-autonomous kernel with that many buffers per warp, 0 = CTA tiles), "rows_rpt2_max_k"
 *   prefilters: "covstats_fused" (1 = CovStats as one histogram pass, 0 = coverage matrix + sort + reduce by key),
 *            "join_tiled" (Join / Remove: 1 = tiled union through shared memory, 0 = one global merge-path search per thread),
 *            "join_tile_kb" (shared-memory budget per CTA of the tiled union)
 *   routed:  "route_stage_depth" (2..4 staging areas per route CTA), "route_blocks_per_sm", "routed_search_blocks_per_sm", "gather_blocks_per_sm"
 * Unknown names fail with CC_ERR_ARG. */
CC_API int cc_set_option(const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif
