/*
 * ctx_oracle.h -- CPU restatement of Corticall's k-mer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (libcorticall_cuda, corticall_b200/) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker or as the timed CPU baseline.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against the reference's
 * own golden vectors: the 66-record table of CortexGraphTest.java:71-136, the find hits/miss of
 * :310-331, the encode/decode round trip of :267-280, SequenceUtilsTest.java:19-72 and the
 * TempGraphAssembler record strings of TraversalEngineTest.java:48-95.  The novelty predicate and the
 * pre-filter commands (FindLowCoverage, FindShared, RecoverExcludedKmers, CovStats) have NO tests in the
 * reference: for them parity is unpinned by reference vectors and rests on the cited source lines, a
 * hand-computed case and the agreement of this C restatement with the numpy one.  The reference itself (Java)
 * cannot be compiled or run in this image (no JDK, jars not vendored), so there is no oracle/_ref.
 *
 * Path prefix used in citations:  S/ = public/java/src/uk/ac/ox/well/cortexjdk/
 */
#ifndef CTX_ORACLE_H
#define CTX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_NAME 256

/* status codes shared with the tests */
enum { ORC_OK = 0, ORC_NOT_CORTEX = 1, ORC_BAD_VERSION = 2, ORC_BAD_TRAILER = 3, ORC_IO = 4,
       ORC_UNSORTED = 5, ORC_RANGE = 6 };

typedef struct {
    uint32_t version, kmer_size, kmer_bits, num_colors;
    uint64_t data_offset;   /* file offset of record 0        (CortexGraph.java:145) */
    uint64_t record_size;   /* 8*kmer_bits + 5*num_colors     (CortexGraph.java:148) */
    uint64_t num_records;   /* floor((size-offset)/recsize)   (CortexGraph.java:149) */
} orc_header;

/* A graph view over an in-memory copy of a whole .ctx file. */
typedef struct {
    const uint8_t *file;    /* whole file bytes */
    uint64_t file_size;
    orc_header h;
    /* Emulation of the reference's LRU for the N<=2 quirk: which record indices have
     * been materialised.  Only consulted when num_records <= 2 (for N>=3 the search result does not
     * depend on the cache).  Bit i set = record i is in the cache. */
    uint32_t cached_small;
} orc_graph;

/* S/utils/io/graph/cortex/CortexGraph.java:66-168 */
int orc_open(orc_graph *g, const uint8_t *file, uint64_t file_size);
/* colour name i (NUL-truncated like fixStringsWithEarlyTerminators :50-64); returns length or -1 */
int orc_color_name(const orc_graph *g, uint32_t color, char *buf, size_t cap);
/* CortexGraph.getColorForSampleName :335-354 */
int orc_color_for_sample_name(const orc_graph *g, const char *name);

/* CortexGraph.getNextRecord :189-237 -- record i as the Java object holds it:
 * binary_kmer[s] are the BYTE-SWAPPED on-disk words (big-endian getLong of LE bytes),
 * coverages are Java ints (BinaryUtils.toUnsignedInt wraps), edges raw bytes.
 * Returns 0, or ORC_RANGE if i >= num_records (Java returns null). */
int orc_get_record(orc_graph *g, uint64_t i, int64_t *binary_kmer, int32_t *coverages, uint8_t *edges);

/* CortexRecord.decodeBinaryKmer :291-307 / encodeBinaryKmer :313-334 (Java long[] convention). */
void orc_decode_binary_kmer(const int64_t *binary_kmer, uint32_t kmer_size, uint32_t kmer_bits, uint8_t *out);
int  orc_encode_binary_kmer(const uint8_t *kmer, uint32_t kmer_size, int64_t *out); /* -1: non-ACGTacgt (Java throws) */
uint32_t orc_kmer_bits(uint32_t kmer_size);           /* CortexRecord.getKmerBits :309-311 */
/* CortexRecord.getEdgesAsBytes :117-140 -> 8 chars */
void orc_edges_to_string(uint8_t edge, char out[9]);

/* S/utils/sequence/SequenceUtils.java:61-86, :127-135, :206-225 */
uint8_t orc_complement(uint8_t b);
void orc_reverse_complement(const uint8_t *seq, size_t n, uint8_t *out);
/* writes canonical orientation to out, returns 1 if the reverse complement was chosen */
int  orc_lowest_orientation(const uint8_t *seq, size_t n, uint8_t *out);
/* S/utils/kmer/CortexByteKmer.java:41-49 (signed bytes, over length n) */
int  orc_byte_kmer_compare(const uint8_t *a, const uint8_t *b, size_t n);

/* CortexGraph.findRecord(byte[]) :272-317.  Returns record index, -1 for null,
 * -2 if the reference would throw "Records are not sorted". */
int64_t orc_find_record(orc_graph *g, const uint8_t *kmer);

/* commands/discover/roi/FindROIs.java:72-82 */
int orc_is_novel(const int32_t *coverages, const int32_t *parents, int nparents, int32_t child);
/* FindROIs.execute :52-67 + CortexGraphWriter.addRecord :106-138: appends the (8s+5)-byte output
 * records of all novel k-mers, in input order, to out (capacity cap records).  If out_index is
 * non-NULL it receives the input record index of each.  faithful!=0 also performs, per record, the
 * k-step k-mer string decode the reference does for its LRU key (CortexGraph.java:225).
 * Returns the number of novel records (may exceed cap; only cap are stored). */
uint64_t orc_find_rois(orc_graph *g, int32_t child, const int32_t *parents, int nparents,
                       uint8_t *out, uint64_t *out_index, uint64_t cap, int faithful);
/* Same scan over a headerless record array (bench: bodies generated in memory). */
uint64_t orc_find_rois_body(const uint8_t *body, uint64_t n, uint32_t kmer_size, uint32_t kmer_bits, uint32_t num_colors,
                            int32_t child, const int32_t *parents, int nparents,
                            uint8_t *out, uint64_t *out_index, uint64_t cap, int faithful);

/* FindROIs.makeCortexHeader :85-105 + CortexGraphWriter.initialize :31-104.
 * Writes the 1-colour ROI header; returns its length (76 + strlen(name)). */
size_t orc_write_roi_header(uint32_t kmer_size, uint32_t kmer_bits, const char *sample_name, uint8_t *out, size_t cap);

/* Batch drivers used for timing and bulk parity (loops of the functions above). */
/* Call.loadChildWalk :2358-2381 -- every window of seq looked up; out[i] = index / -1 / -2 */
void orc_find_windows(orc_graph *g, const uint8_t *seq, uint64_t len, int64_t *out);
/* nq independent k-byte queries, row-major */
void orc_find_batch(orc_graph *g, const uint8_t *kmers, uint64_t nq, int64_t *out);
/* canonicalise + 2-bit pack every window: words[(len-k+1)*s] NATIVE order (word0 most significant,
 * == byteswap of the Java long), flags bit0 flipped, bit1 not packable (non-ACGT; words zeroed). */
void orc_pack_windows(const uint8_t *seq, uint64_t len, uint32_t kmer_size, uint64_t *words, uint8_t *flags);

/* ---- scan-shaped pre-filters, restated record by record through orc_get_record / orc_find_record.
 * Each fills one decision per record of the iterated graph; the tests assemble the output files from them and compare
 * with the numpy restatement (oracle_np.py), which was written independently.  parents / ignore may hold -1 (a sample
 * name that did not resolve: the reference's HashSet<Integer> then never matches). */
/* S/commands/prefilter/FindLowCoverage.java:47-58: written[i] = !(coverage(0) >= min_coverage) */
void orc_find_low_coverage(orc_graph *roi, int32_t min_coverage, uint8_t *written);
/* S/commands/prefilter/FindShared.java:60-109: written[i] = ROI record i is shared.  Returns -1 when GRAPH.findRecord
 * gives null for a ROI k-mer (NullPointerException at :67 in the reference), else 0. */
int orc_find_shared(orc_graph *graph, orc_graph *roi, int32_t child, const int32_t *parents, int nparents,
                    const int32_t *ignore, int nignore, uint8_t *written);
/* S/commands/discover/recover/RecoverExcludedKmers.java:49-92: written[i] = 1 (as is) / 2 (recovered) / 0, cov0[i] =
 * coverages[0] of the record that CortexGraphWriter.addRecord receives (after coverages[child] = dirty coverage).
 * Returns the number of recovered records. */
uint64_t orc_recover_excluded_kmers(orc_graph *graph, orc_graph *dirty, int32_t child, uint8_t *written, int32_t *cov0);
/* S/commands/utils/CovStats.java:46-66: per record, key[i] = coverage(child) if the record counts (else 0) and
 * weight[i] = numberOfParents + numberOfChildren. */
void orc_cov_stats_pairs(orc_graph *graph, int32_t child, const int32_t *parents, int nparents, int32_t *key, int32_t *weight);

/* S/commands/utils/Remove.java:30-88 over CortexCollection.next (:245-293), record by record: graphs[0] is the primary; kept records
 * (primary colours, on-disk layout) go to out (at most cap); returns the number kept, *removed the number dropped. */
uint64_t orc_remove(orc_graph **graphs, int ngraphs, uint8_t *out, uint64_t cap, uint64_t *removed);

#ifdef __cplusplus
}
#endif
#endif
