/*
 * ctx_oracle.c -- CPU restatement of Corticall's k-mer hot path.  TEST INFRASTRUCTURE ONLY
 * (see ctx_oracle.h for the rules and the parity-pinning status).
 *
 * Every function names the reference lines it restates.  The restatement keeps the reference's
 * CONTROL FLOW (three-point binary search, per-base shifts, signed-byte compares, Java int wrap)
 * because that is what decides bit-exactness; it does not simulate JVM allocation or the LRU map
 * except where the LRU changes results (N <= 2).
 *
 * S/ = public/java/src/uk/ac/ox/well/cortexjdk/
 */
#include "ctx_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <strings.h>

/* ---------------------------------------------------------------- little helpers */

/* S/utils/io/utils/BinaryFile.java:18-32 and BinaryUtils.java:6-17: four LE bytes -> Java int
 * (values >= 2^31 wrap negative). */
static int32_t le_u32_as_java_int(const uint8_t *b) {
    uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    return (int32_t)v;
}

/* ByteBuffer.wrap(b).getLong() with the default BIG_ENDIAN order (CortexGraph.java:208-209). */
static int64_t be_bytes_as_java_long(const uint8_t *b) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v = (v << 8) | b[i];
    return (int64_t)v;
}

/* CortexRecord.reverse :370-377 -- byte swap of a long. */
static uint64_t swap64(uint64_t x) {
    uint64_t r = 0;
    for (int i = 0; i < 8; i++) { r = (r << 8) | (x & 0xff); x >>= 8; }
    return r;
}

static int magic_ok(const uint8_t *p) {          /* equalsIgnoreCase("CORTEX") :74,:140 */
    return strncasecmp((const char *)p, "CORTEX", 6) == 0;
}

uint32_t orc_kmer_bits(uint32_t kmer_size) {     /* CortexRecord.getKmerBits :309-311 */
    return (kmer_size + 31) / 32;
}

/* ---------------------------------------------------------------- header */

/* Walks the header exactly in the order CortexGraph.loadCortexGraph :70-142 reads it.
 * If names != NULL it receives the file offsets of each colour's name-length field. */
static int walk_header(const uint8_t *f, uint64_t n, orc_header *h, uint64_t *name_pos, uint32_t max_names) {
    uint64_t p = 0;
    if (n < 22 || !magic_ok(f)) return ORC_NOT_CORTEX;
    p = 6;
    h->version = (uint32_t)le_u32_as_java_int(f + p); p += 4;
    if (h->version != 6) return ORC_BAD_VERSION;
    h->kmer_size  = (uint32_t)le_u32_as_java_int(f + p); p += 4;
    h->kmer_bits  = (uint32_t)le_u32_as_java_int(f + p); p += 4;
    h->num_colors = (uint32_t)le_u32_as_java_int(f + p); p += 4;
    uint64_t c = h->num_colors;
    p += 4 * c;                                   /* mean read lengths :94-96 */
    p += 8 * c;                                   /* total sequence    :98-100 */
    for (uint64_t i = 0; i < c; i++) {            /* sample names      :102-111 */
        if (p + 4 > n) return ORC_IO;
        if (name_pos && i < max_names) name_pos[i] = p;
        uint32_t L = (uint32_t)le_u32_as_java_int(f + p); p += 4 + (uint64_t)L;
    }
    p += 16 * c;                                  /* error rates       :114-117 */
    for (uint64_t i = 0; i < c; i++) {            /* cleaning blocks   :119-134 */
        if (p + 16 > n) return ORC_IO;
        p += 12;
        uint32_t G = (uint32_t)le_u32_as_java_int(f + p); p += 4 + (uint64_t)G;
    }
    if (p + 6 > n) return ORC_IO;
    if (!magic_ok(f + p)) return ORC_BAD_TRAILER; /* :136-142 */
    p += 6;
    h->data_offset = p;                           /* :145 */
    h->record_size = 8ull * h->kmer_bits + 5ull * h->num_colors;   /* :148 */
    h->num_records = h->record_size ? (n - p) / h->record_size : 0; /* :149 (floors) */
    return ORC_OK;
}

int orc_open(orc_graph *g, const uint8_t *file, uint64_t file_size) {
    memset(g, 0, sizeof *g);
    g->file = file;
    g->file_size = file_size;
    int rc = walk_header(file, file_size, &g->h, NULL, 0);
    if (rc != ORC_OK) return rc;
    /* loadCortexGraph ends with position(0) (:162) which materialises record 0 into the LRU. */
    g->cached_small = (g->h.num_records > 0) ? 1u : 0u;
    return ORC_OK;
}

int orc_color_name(const orc_graph *g, uint32_t color, char *buf, size_t cap) {
    if (color >= g->h.num_colors || color >= 4096) return -1;
    uint64_t *pos = (uint64_t *)malloc(sizeof(uint64_t) * g->h.num_colors);
    orc_header tmp;
    walk_header(g->file, g->file_size, &tmp, pos, g->h.num_colors);
    uint64_t p = pos[color];
    free(pos);
    uint32_t L = (uint32_t)le_u32_as_java_int(g->file + p);
    const uint8_t *s = g->file + p + 4;
    /* fixStringsWithEarlyTerminators :50-64: cut at the FIRST NUL */
    uint32_t len = L;
    for (uint32_t i = 0; i < L; i++) if (s[i] == 0) { len = i; break; }
    if (len + 1 > cap) return -1;
    memcpy(buf, s, len);
    buf[len] = 0;
    return (int)len;
}

int orc_color_for_sample_name(const orc_graph *g, const char *name) {   /* :335-354 */
    int color = -1, copies = 0;
    char buf[ORC_MAX_NAME];
    for (uint32_t c = 0; c < g->h.num_colors; c++) {
        if (orc_color_name(g, c, buf, sizeof buf) >= 0 && strcasecmp(buf, name) == 0) { color = (int)c; copies++; }
    }
    if (color == -1) {                            /* Integer.valueOf(sampleName) */
        char *end = NULL;
        const char *p = name;
        if (*p == '+' || *p == '-') p++;
        if (*p) {
            int all_digits = 1;
            for (const char *q = p; *q; q++) if (*q < '0' || *q > '9') all_digits = 0;
            if (all_digits) { long v = strtol(name, &end, 10); color = (int)v; copies = 1; }
        }
    }
    return copies == 1 ? color : -1;
}

/* ---------------------------------------------------------------- records */

int orc_get_record(orc_graph *g, uint64_t i, int64_t *binary_kmer, int32_t *coverages, uint8_t *edges) {
    if (i >= g->h.num_records) return ORC_RANGE;              /* :190,:236 -> null */
    const uint8_t *p = g->file + g->h.data_offset + i * g->h.record_size;   /* :197 */
    for (uint32_t w = 0; w < g->h.kmer_bits; w++) { binary_kmer[w] = be_bytes_as_java_long(p); p += 8; }  /* :202-210 */
    for (uint32_t c = 0; c < g->h.num_colors; c++) { coverages[c] = le_u32_as_java_int(p); p += 4; }      /* :212-218 */
    memcpy(edges, p, g->h.num_colors);                                                                     /* :220-221 */
    if (i < 32) g->cached_small |= (1u << i);                                                              /* :224-225 */
    return ORC_OK;
}

void orc_decode_binary_kmer(const int64_t *kmer, uint32_t kmer_size, uint32_t kmer_bits, uint8_t *out) {
    static const uint8_t ch[4] = { 'A', 'C', 'G', 'T' };
    uint64_t w[8];
    uint64_t *bk = kmer_bits <= 8 ? w : (uint64_t *)malloc(8 * (size_t)kmer_bits);
    for (uint32_t i = 0; i < kmer_bits; i++) bk[i] = swap64((uint64_t)kmer[i]);      /* :296-298 */
    for (int64_t i = (int64_t)kmer_size - 1; i >= 0; i--) {                            /* :300-304 */
        out[i] = ch[bk[kmer_bits - 1] & 3];
        for (uint32_t j = kmer_bits - 1; j > 0; j--) {                                 /* shiftBinaryKmerByOneBase :362-368 */
            bk[j] >>= 2;
            bk[j] |= bk[j - 1] << 62;
        }
        bk[0] >>= 2;
    }
    if (bk != w) free(bk);
}

static int nuc_code(uint8_t b) {                  /* charToBinaryNucleotide :347-360 */
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

int orc_encode_binary_kmer(const uint8_t *kmer, uint32_t kmer_size, int64_t *out) {   /* :313-334 */
    int32_t nb = (int32_t)orc_kmer_bits(kmer_size);
    int32_t len = (int32_t)kmer_size;
    for (int32_t b = 0; b < nb; b++) {
        uint64_t acc = 0;
        for (int32_t i = len - 32 * (b + 1); i < len - 32 * b; i++) {
            if (i >= 0) {
                int code = nuc_code(kmer[i]);
                if (code < 0) return -1;
                acc |= (uint64_t)code;
            }
            if (i < len - 32 * b - 1) acc <<= 2;
        }
        out[nb - b - 1] = (int64_t)swap64(acc);
    }
    return 0;
}

void orc_edges_to_string(uint8_t edge, char out[9]) {          /* CortexRecord.getEdgesAsBytes :117-140 */
    static const char str[8] = { 'a', 'c', 'g', 't', 'A', 'C', 'G', 'T' };
    int left = ((int8_t)edge) >> 4;                              /* Java byte is signed; only low 4 bits used */
    int right = edge & 0xf;
    for (int i = 0; i < 4; i++) {
        out[i] = (left & (1 << (3 - i))) ? str[i] : '.';
        out[i + 4] = (right & (1 << i)) ? str[i + 4] : '.';
    }
    out[8] = 0;
}

/* ---------------------------------------------------------------- sequence utils */

uint8_t orc_complement(uint8_t b) {               /* SequenceUtils.complement :61-86 */
    switch (b) {
        case 'A': return 'T'; case 'a': return 't';
        case 'C': return 'G'; case 'c': return 'g';
        case 'G': return 'C'; case 'g': return 'c';
        case 'T': return 'A'; case 't': return 'a';
        case 'N': return 'N'; case 'n': return 'n';
        case '.': return '.';
        default:  return b;
    }
}

void orc_reverse_complement(const uint8_t *seq, size_t n, uint8_t *out) {   /* :127-135 */
    for (size_t i = 0; i < n; i++) out[n - 1 - i] = orc_complement(seq[i]);
}

int orc_lowest_orientation(const uint8_t *seq, size_t n, uint8_t *out) {    /* :206-225 */
    for (size_t i = 0; i < n; i++) {
        int8_t rc = (int8_t)orc_complement(seq[n - 1 - i]);
        int8_t fw = (int8_t)seq[i];                /* Java bytes compare signed */
        if (fw < rc) { memmove(out, seq, n); return 0; }
        if (fw > rc) {
            if (out == seq) {                      /* in-place request: go through a temporary */
                uint8_t *tmp = (uint8_t *)malloc(n);
                orc_reverse_complement(seq, n, tmp);
                memcpy(out, tmp, n);
                free(tmp);
            } else {
                orc_reverse_complement(seq, n, out);
            }
            return 1;
        }
    }
    memmove(out, seq, n);
    return 0;
}

int orc_byte_kmer_compare(const uint8_t *a, const uint8_t *b, size_t n) {   /* CortexByteKmer.compareTo :41-49 */
    for (size_t i = 0; i < n; i++) {
        if ((int8_t)a[i] < (int8_t)b[i]) return -1;
        if ((int8_t)a[i] > (int8_t)b[i]) return 1;
    }
    return 0;
}

/* ---------------------------------------------------------------- findRecord */

static void record_kmer_bytes(orc_graph *g, uint64_t i, uint8_t *out) {
    int64_t bk[64];
    int32_t cov_stack[64]; uint8_t edge_stack[64];
    uint32_t c = g->h.num_colors, s = g->h.kmer_bits;
    int64_t *bkp = s <= 64 ? bk : (int64_t *)malloc(8 * (size_t)s);
    int32_t *cov = c <= 64 ? cov_stack : (int32_t *)malloc(4 * (size_t)c);
    uint8_t *ed = c <= 64 ? edge_stack : (uint8_t *)malloc(c);
    orc_get_record(g, i, bkp, cov, ed);            /* getRecord -> position -> getNextRecord :183-187 */
    orc_decode_binary_kmer(bkp, g->h.kmer_size, s, out);   /* getKmerAsByteKmer, CortexRecord.java:113 */
    if (bkp != bk) free(bkp);
    if (cov != cov_stack) free(cov);
    if (ed != edge_stack) free(ed);
}

int64_t orc_find_record(orc_graph *g, const uint8_t *query) {   /* CortexGraph.java:272-317 */
    uint32_t k = g->h.kmer_size;
    uint8_t qs[256], a[256], m[256], z[256];
    uint8_t *heap = NULL;
    uint8_t *q = qs, *ka = a, *km = m, *kz = z;
    if (k > 256) { heap = (uint8_t *)malloc(4 * (size_t)k); q = heap; ka = heap + k; km = heap + 2 * k; kz = heap + 3 * k; }
    int64_t result = -1;

    orc_lowest_orientation(query, k, q);                         /* :273 */

    int64_t n = (int64_t)g->h.num_records;
    if (n <= 2) {
        /* :274-276 -- the LRU can answer before the search; for N<=2 it is the ONLY way to hit
         * because the loop below never runs. */
        for (int64_t i = 0; i < n; i++) {
            if (g->cached_small & (1u << i)) {
                uint32_t keep = g->cached_small;
                record_kmer_bytes(g, (uint64_t)i, ka);
                g->cached_small = keep;
                if (memcmp(ka, q, k) == 0) { result = i; goto done; }
            }
        }
    }

    {
        int64_t start = 0, stop = n - 1;
        int64_t mid = start + (stop - start) / 2;                /* Java division truncates toward 0 :278-280 */
        while (start != mid && mid != stop) {                    /* :282 */
            record_kmer_bytes(g, (uint64_t)start, ka);           /* :285-293 */
            record_kmer_bytes(g, (uint64_t)mid, km);
            record_kmer_bytes(g, (uint64_t)stop, kz);
            if (orc_byte_kmer_compare(ka, kz, k) > 0) { result = -2; goto done; }   /* :295-297 */
            if (orc_byte_kmer_compare(ka, km, k) > 0) { result = -2; goto done; }   /* :299-301 */
            if (orc_byte_kmer_compare(q, kz, k) > 0 || orc_byte_kmer_compare(q, ka, k) < 0) { result = -1; goto done; }  /* :303 */
            else if (memcmp(ka, q, k) == 0) { result = start; goto done; }          /* :304 */
            else if (memcmp(km, q, k) == 0) { result = mid; goto done; }            /* :305 */
            else if (memcmp(kz, q, k) == 0) { result = stop; goto done; }           /* :306 */
            else if (orc_byte_kmer_compare(q, ka, k) > 0 && orc_byte_kmer_compare(q, km, k) < 0) {   /* :307-309 */
                stop = mid;
                mid = start + (stop - start) / 2;
            } else if (orc_byte_kmer_compare(q, km, k) > 0 && orc_byte_kmer_compare(q, kz, k) < 0) { /* :310-312 */
                start = mid;
                mid = start + (stop - start) / 2;
            }
        }
    }
done:
    if (heap) free(heap);
    return result;
}

/* ---------------------------------------------------------------- FindROIs */

int orc_is_novel(const int32_t *cov, const int32_t *parents, int nparents, int32_t child) {   /* FindROIs.java:72-82 */
    int parents_lack = 1;
    for (int i = 0; i < nparents; i++) parents_lack &= (cov[parents[i]] == 0);
    int child_has = cov[child] > 0;                 /* signed compare on the wrapped int */
    return child_has && parents_lack;
}

uint64_t orc_find_rois_body(const uint8_t *body, uint64_t n, uint32_t kmer_size, uint32_t s, uint32_t c,
                            int32_t child, const int32_t *parents, int nparents,
                            uint8_t *out, uint64_t *out_index, uint64_t cap, int faithful) {
    uint64_t S = 8ull * s + 5ull * c, O = 8ull * s + 5;
    uint64_t novel = 0;
    int64_t *bk = (int64_t *)malloc(8 * (size_t)(s ? s : 1));
    int32_t *cov = (int32_t *)malloc(4 * (size_t)(c ? c : 1));
    uint8_t *ed = (uint8_t *)malloc(c ? c : 1);
    uint8_t *kstr = (uint8_t *)malloc(kmer_size ? kmer_size : 1);
    volatile uint8_t sink = 0;
    for (uint64_t i = 0; i < n; i++) {              /* for (CortexRecord cr : GRAPH) FindROIs.java:52 */
        const uint8_t *p = body + i * S;
        for (uint32_t w = 0; w < s; w++) { bk[w] = be_bytes_as_java_long(p); p += 8; }
        for (uint32_t j = 0; j < c; j++) { cov[j] = le_u32_as_java_int(p); p += 4; }
        memcpy(ed, p, c);
        if (faithful) {                             /* cache.put(cr.getKmerAsByteKmer(), cr) CortexGraph.java:225 */
            orc_decode_binary_kmer(bk, kmer_size, s, kstr);
            sink ^= kstr[0];
        }
        if (orc_is_novel(cov, parents, nparents, child)) {       /* :53 */
            if (novel < cap) {
                /* CortexGraphWriter.addRecord :106-138: longs big-endian (= original disk bytes),
                 * child coverage LE, child edge byte. */
                uint8_t *o = out + novel * O;
                for (uint32_t w = 0; w < s; w++) {
                    uint64_t v = (uint64_t)bk[w];
                    for (int b = 7; b >= 0; b--) { o[b] = (uint8_t)(v & 0xff); v >>= 8; }
                    o += 8;
                }
                uint32_t cv = (uint32_t)cov[child];
                o[0] = (uint8_t)cv; o[1] = (uint8_t)(cv >> 8); o[2] = (uint8_t)(cv >> 16); o[3] = (uint8_t)(cv >> 24);
                o[4] = ed[child];
                if (out_index) out_index[novel] = i;
            }
            novel++;
        }
    }
    (void)sink;
    free(bk); free(cov); free(ed); free(kstr);
    return novel;
}

uint64_t orc_find_rois(orc_graph *g, int32_t child, const int32_t *parents, int nparents,
                       uint8_t *out, uint64_t *out_index, uint64_t cap, int faithful) {
    return orc_find_rois_body(g->file + g->h.data_offset, g->h.num_records, g->h.kmer_size, g->h.kmer_bits,
                              g->h.num_colors, child, parents, nparents, out, out_index, cap, faithful);
}

static void put_u32(uint8_t **p, uint32_t v) { (*p)[0] = (uint8_t)v; (*p)[1] = (uint8_t)(v >> 8); (*p)[2] = (uint8_t)(v >> 16); (*p)[3] = (uint8_t)(v >> 24); *p += 4; }

size_t orc_write_roi_header(uint32_t kmer_size, uint32_t kmer_bits, const char *name, uint8_t *out, size_t cap) {
    /* CortexGraphWriter.initialize :45-94 with the header FindROIs.makeCortexHeader :85-105 builds:
     * 1 colour, meanReadLength 0, totalSequence 0, flags 0, thresholds 0, cleaned-against name "". */
    static const uint8_t err[16] = { 0, 0xd8, 0xa3, 0x70, 0x3d, 0x0a, 0xd7, 0xa3, 0xf8, 0x3f, 0, 0, 0, 0, 0, 0 };  /* :76 */
    size_t L = strlen(name), need = 76 + L;
    if (cap < need) return 0;
    uint8_t *p = out;
    memcpy(p, "CORTEX", 6); p += 6;
    put_u32(&p, 6); put_u32(&p, kmer_size); put_u32(&p, kmer_bits); put_u32(&p, 1);
    put_u32(&p, 0);                                  /* mean read length */
    put_u32(&p, 0); put_u32(&p, 0);                  /* total sequence (u64) */
    put_u32(&p, (uint32_t)L); memcpy(p, name, L); p += L;
    memcpy(p, err, 16); p += 16;
    p[0] = p[1] = p[2] = p[3] = 0; p += 4;           /* four booleans */
    put_u32(&p, 0); put_u32(&p, 0);                  /* thresholds */
    put_u32(&p, 0);                                  /* cleaned-against name length 0 */
    memcpy(p, "CORTEX", 6); p += 6;
    return (size_t)(p - out);
}

/* ---------------------------------------------------------------- batch drivers */

void orc_find_windows(orc_graph *g, const uint8_t *seq, uint64_t len, int64_t *out) {   /* Call.java:2358-2381 */
    uint32_t k = g->h.kmer_size;
    if (len < k) return;
    for (uint64_t i = 0; i + k <= len; i++) out[i] = orc_find_record(g, seq + i);
}

void orc_find_batch(orc_graph *g, const uint8_t *kmers, uint64_t nq, int64_t *out) {
    uint32_t k = g->h.kmer_size;
    for (uint64_t i = 0; i < nq; i++) out[i] = orc_find_record(g, kmers + i * k);
}

void orc_pack_windows(const uint8_t *seq, uint64_t len, uint32_t k, uint64_t *words, uint8_t *flags) {
    uint32_t s = orc_kmer_bits(k);
    if (len < k) return;
    uint8_t *canon = (uint8_t *)malloc(k);
    int64_t *jl = (int64_t *)malloc(8 * (size_t)s);
    for (uint64_t i = 0; i + k <= len; i++) {
        int flipped = orc_lowest_orientation(seq + i, k, canon);     /* CortexBinaryKmer(byte[]) ctor, CortexBinaryKmer.java:15-17 */
        int lower = 0;
        for (uint32_t j = 0; j < k; j++) lower |= (canon[j] >= 'a' && canon[j] <= 'z');
        if (orc_encode_binary_kmer(canon, k, jl) != 0) {
            for (uint32_t w = 0; w < s; w++) words[i * s + w] = 0;
            flags[i] = 2;
        } else {
            for (uint32_t w = 0; w < s; w++) words[i * s + w] = swap64((uint64_t)jl[w]);
            flags[i] = (uint8_t)((flipped ? 1 : 0) | (lower ? 4 : 0));
        }
    }
    free(canon); free(jl);
}


/* ------------------------------------------------------------------ scan-shaped pre-filters */
#define ORC_MAX_WORDS 64          /* k <= 2048 */
static int in_list(const int32_t *list, int n, int32_t v) {
    for (int i = 0; i < n; ++i) if (list[i] == v) return 1;
    return 0;
}

void orc_find_low_coverage(orc_graph *roi, int32_t min_coverage, uint8_t *written) {   /* FindLowCoverage.java:47-58 */
    int64_t bk[ORC_MAX_WORDS]; int32_t *cov = malloc(4 * (roi->h.num_colors + 1)); uint8_t *ed = malloc(roi->h.num_colors + 1);
    for (uint64_t i = 0; i < roi->h.num_records; ++i) {                 /* for (CortexRecord cr : ROI) */
        orc_get_record(roi, i, bk, cov, ed);
        if (cov[0] >= min_coverage) written[i] = 0;                      /* numKept++ */
        else written[i] = 1;                                             /* cgw.addRecord(cr) */
    }
    free(cov); free(ed);
}

int orc_find_shared(orc_graph *graph, orc_graph *roi, int32_t child, const int32_t *parents, int nparents,
                    const int32_t *ignore, int nignore, uint8_t *written) {            /* FindShared.java:60-109 */
    int64_t bk[ORC_MAX_WORDS], gk[ORC_MAX_WORDS];
    int32_t *rcov = malloc(4 * (roi->h.num_colors + 1)), *gcov = malloc(4 * (graph->h.num_colors + 1));
    uint8_t *red = malloc(roi->h.num_colors + 1), *ged = malloc(graph->h.num_colors + 1);
    uint8_t *kmer = malloc(roi->h.kmer_size + 1);
    int rc = 0;
    for (uint64_t i = 0; i < roi->h.num_records && rc == 0; ++i) {
        orc_get_record(roi, i, bk, rcov, red);
        orc_decode_binary_kmer(bk, roi->h.kmer_size, roi->h.kmer_bits, kmer);           /* rr.getCanonicalKmer() */
        const int64_t at = orc_find_record(graph, kmer);                                /* GRAPH.findRecord(...) :65 */
        if (at >= 0) orc_get_record(graph, (uint64_t)at, gk, gcov, ged);
        int shared = 0;
        for (uint32_t c = 0; c < graph->h.num_colors; ++c) {                            /* :67-73 */
            if ((int32_t)c != child && !in_list(parents, nparents, (int32_t)c) && !in_list(ignore, nignore, (int32_t)c)) {
                if (at < 0) { rc = -1; break; }                                         /* cr == null: cr.getCoverage(c) is the NPE at :67; */
                if (gcov[c] > 0) {                                                      /* the && short-circuits, so only a free colour gets here */
                    shared = 1;
                    break;
                }
            }
        }
        if (rc) break;
        written[i] = (uint8_t)shared;                                                   /* second loop :95-104 writes the shared ones */
    }
    free(rcov); free(gcov); free(red); free(ged); free(kmer);
    return rc;
}

uint64_t orc_recover_excluded_kmers(orc_graph *graph, orc_graph *dirty, int32_t child, uint8_t *written, int32_t *cov0) {
    int64_t bk[ORC_MAX_WORDS], dk[ORC_MAX_WORDS];                                       /* RecoverExcludedKmers.java:49-92 */
    int32_t *cov = malloc(4 * (graph->h.num_colors + 1)), *dcov = malloc(4 * (dirty->h.num_colors + 1));
    uint8_t *ed = malloc(graph->h.num_colors + 1), *ded = malloc(dirty->h.num_colors + 1);
    uint8_t *kmer = malloc(graph->h.kmer_size + 1);
    uint64_t recovered = 0;
    for (uint64_t i = 0; i < graph->h.num_records; ++i) {
        orc_get_record(graph, i, bk, cov, ed);
        written[i] = 0;
        cov0[i] = cov[0];
        if (cov[child] > 0) { written[i] = 1; continue; }                              /* :50-52 */
        int others = 0;
        for (uint32_t c = 0; c < graph->h.num_colors; ++c)                              /* :55-59 */
            if ((int32_t)c != child && cov[c] > 0) ++others;
        if (others > 0) {
            orc_decode_binary_kmer(bk, graph->h.kmer_size, graph->h.kmer_bits, kmer);
            const int64_t at = orc_find_record(dirty, kmer);                            /* DIRTY.findRecord(cr.getCanonicalKmer()) :63 */
            if (at >= 0) {
                orc_get_record(dirty, (uint64_t)at, dk, dcov, ded);
                if (dcov[0] > 0) {                                                      /* :65 */
                    cov[child] = dcov[0];                                               /* coverages[childColor] = dr.getCoverage(0) :74 */
                    written[i] = 2;
                    cov0[i] = cov[0];                                                   /* the writer emits colour 0 only (1-colour header) */
                    ++recovered;
                }
            }
        }
    }
    free(cov); free(dcov); free(ed); free(ded); free(kmer);
    return recovered;
}

void orc_cov_stats_pairs(orc_graph *graph, int32_t child, const int32_t *parents, int nparents, int32_t *key, int32_t *weight) {
    int64_t bk[ORC_MAX_WORDS];                                                          /* CovStats.java:46-66 */
    int32_t *cov = malloc(4 * (graph->h.num_colors + 1));
    uint8_t *ed = malloc(graph->h.num_colors + 1);
    for (uint64_t i = 0; i < graph->h.num_records; ++i) {
        orc_get_record(graph, i, bk, cov, ed);
        int is_in_child = cov[child] > 0;                                               /* :47 */
        int np = 0, nc = 0;
        for (uint32_t c = 0; c < graph->h.num_colors; ++c) {                            /* :51-57 */
            if (cov[c] > 0) {
                if (child == (int32_t)c) is_in_child = 1;
                else if (in_list(parents, nparents, (int32_t)c)) ++np;
                else ++nc;
            }
        }
        const int counts = is_in_child && np > 0 && nc > 0;                            /* :61 */
        key[i] = counts ? cov[child] : 0;
        weight[i] = counts ? np + nc : 0;
    }
    free(cov); free(ed);
}

/* ---------------------------------------------------------------- Remove over a CortexCollection
 * S/commands/utils/Remove.java:30-88 driving S/utils/io/graph/cortex/CortexCollection.java:245-293: next() picks the graph whose
 * pending record has the smallest k-mer STRING, merges every graph whose pending record has that same string into one record
 * (colours concatenated in graph order, absent graphs zero), and advances those graphs.  Remove drops the merged record when a
 * colour >= PGRAPH.getNumColors() has coverage > 0 and otherwise writes the primary's colours.  graphs[0] is the primary.
 * out: kept records in the primary's on-disk layout (s LE words, cp LE uint32, cp edge bytes), at most cap of them.
 * Returns the number kept; *removed gets the number dropped. */
uint64_t orc_remove(orc_graph **graphs, int ngraphs, uint8_t *out, uint64_t cap, uint64_t *removed) {
    const uint32_t k = graphs[0]->h.kmer_size, s = graphs[0]->h.kmer_bits, cp = graphs[0]->h.num_colors;
    uint32_t ctot = 0;
    for (int i = 0; i < ngraphs; ++i) ctot += graphs[i]->h.num_colors;
    uint64_t *pos = calloc((size_t)ngraphs, sizeof(uint64_t));                      /* nextRecs[i] = graph i's pending record */
    uint8_t *kstr = malloc((size_t)ngraphs * (k + 1));
    int64_t bk[ORC_MAX_WORDS], merged_bk[ORC_MAX_WORDS];
    int32_t *cov = malloc(4 * (size_t)(ctot + 1)), *gcov = malloc(4 * (size_t)(ctot + 1));
    uint8_t *edges = malloc(ctot + 1), *ged = malloc(ctot + 1);
    for (int i = 0; i < ngraphs; ++i)
        if (pos[i] < graphs[i]->h.num_records) record_kmer_bytes(graphs[i], pos[i], kstr + (size_t)i * (k + 1));
    uint64_t kept = 0, dropped = 0;
    const size_t out_size = 8 * (size_t)s + 5 * (size_t)cp;
    for (;;) {
        int lowc = -1;                                                              /* CortexCollection.next :245-254 */
        for (int i = 0; i < ngraphs; ++i) {
            if (pos[i] >= graphs[i]->h.num_records) continue;                       /* nextRecs[i] == null */
            if (lowc == -1 || memcmp(kstr + (size_t)i * (k + 1), kstr + (size_t)lowc * (k + 1), k) < 0) lowc = i;   /* String.compareTo on ACGT */
        }
        if (lowc < 0) break;                                                        /* hasNext() == false */
        uint8_t comp[ORC_MAX_WORDS * 32 + 1];
        memcpy(comp, kstr + (size_t)lowc * (k + 1), k);
        memset(cov, 0, 4 * (size_t)ctot);
        memset(edges, 0, ctot);
        uint32_t c0 = 0;
        for (int i = 0; i < ngraphs; ++i) {                                         /* :262-283 */
            const uint32_t ci = graphs[i]->h.num_colors;
            if (pos[i] < graphs[i]->h.num_records && memcmp(kstr + (size_t)i * (k + 1), comp, k) == 0) {
                orc_get_record(graphs[i], pos[i], bk, gcov, ged);
                memcpy(merged_bk, bk, 8 * (size_t)s);                               /* binaryKmer = cr.getBinaryKmer() */
                for (uint32_t c = 0; c < ci; ++c) { cov[c0 + c] = gcov[c]; edges[c0 + c] = ged[c]; }
                pos[i]++;                                                           /* nextRecs[i] = graphList.get(i).next() */
                if (pos[i] < graphs[i]->h.num_records) record_kmer_bytes(graphs[i], pos[i], kstr + (size_t)i * (k + 1));
            }
            c0 += ci;
        }
        int found = 0;                                                              /* Remove.java:47-55 */
        for (uint32_t c = cp; c < ctot; ++c) if (cov[c] > 0) { found = 1; break; }
        if (!found) {                                                               /* :57-73 + CortexGraphWriter.addRecord */
            if (kept < cap) {
                uint8_t *p = out + kept * out_size;
                for (uint32_t w = 0; w < s; ++w) {                                  /* the writer emits each long big-endian: the on-disk word */
                    const uint64_t v = (uint64_t)merged_bk[w];
                    for (int b = 0; b < 8; ++b) p[8 * w + b] = (uint8_t)(v >> (56 - 8 * b));
                }
                for (uint32_t c = 0; c < cp; ++c) {
                    const uint32_t v = (uint32_t)cov[c];
                    p[8 * s + 4 * c] = (uint8_t)v; p[8 * s + 4 * c + 1] = (uint8_t)(v >> 8);
                    p[8 * s + 4 * c + 2] = (uint8_t)(v >> 16); p[8 * s + 4 * c + 3] = (uint8_t)(v >> 24);
                }
                memcpy(p + 8 * s + 4 * cp, edges, cp);
            }
            kept++;
        } else dropped++;
    }
    free(pos); free(kstr); free(cov); free(gcov); free(edges); free(ged);
    if (removed) *removed = dropped;
    return kept;
}
