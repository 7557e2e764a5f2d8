/* abi_example.c -- the drop-in boundary used from plain C, without Python or Java: open a .ctx graph, run the novel-k-mer step
 * (FindROIs.java:31-105) and look every novel k-mer up again (CortexGraph.findRecord, CortexGraph.java:272-317).
 *   gcc -std=c11 -Iinclude tools/abi_example.c -Lcorticall_b200 -lcorticall_cuda -Wl,-rpath,$PWD/corticall_b200 -o abi_example
 *   ./abi_example tests/golden/two_short_contigs.ctx 0 1        # child colour, then the parent colours
 *   CC_DEVICES=0,1,2,3 ./abi_example graph.ctx 0 1               # also: the same graph sharded over these devices (cc_open_sharded);
 *                                                                # its novel records and lookups must equal the single-GPU answers
 * Exit status: 0 ok, 2 usage, otherwise the cc_status of the failing call (7 = no CUDA device: there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "corticall_cuda.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != CC_OK) {                                                      \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, cc_last_error()); \
            return rc_;                                                          \
        }                                                                        \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s graph.ctx child_colour [parent_colour ...]\n", argv[0]);
        return 2;
    }
    cc_graph *g = NULL;
    CHECK(cc_open(argv[1], 0, &g));
    uint32_t version, k, s, c;
    uint64_t n, data_offset, record_size;
    CHECK(cc_header(g, &version, &k, &s, &c, &n, &data_offset, &record_size));
    printf("version %u k %u words %u colours %u records %llu\n", version, k, s, c, (unsigned long long)n);

    const int32_t child = atoi(argv[2]);
    int32_t parents[64];
    int np = 0;
    for (int i = 3; i < argc && np < 64; ++i) parents[np++] = atoi(argv[i]);
    const uint64_t out_size = 8ull * s + 5;                     /* one-colour record of the ROI graph */
    uint8_t *records = malloc((n ? n : 1) * out_size);
    uint64_t *index = malloc((n ? n : 1) * sizeof(uint64_t));
    uint64_t novel = 0;
    CHECK(cc_find_novel(g, child, parents, np, records, index, n, &novel));
    printf("novel %llu\n", (unsigned long long)novel);

    /* every novel k-mer, as packed canonical words, must be found at the index the scan reported */
    uint64_t *words = malloc((novel ? novel : 1) * s * sizeof(uint64_t));
    int64_t *found = malloc((novel ? novel : 1) * sizeof(int64_t));
    for (uint64_t i = 0; i < novel; ++i) memcpy(words + i * s, records + i * out_size, 8ull * s);
    CHECK(cc_find_packed(g, words, NULL, novel, found, CC_ALGO_AUTO));
    uint64_t ok = 0;
    for (uint64_t i = 0; i < novel; ++i) ok += (found[i] == (int64_t)index[i]);
    printf("found %llu of %llu at the reported index\n", (unsigned long long)ok, (unsigned long long)novel);
    /* the same graph over several devices of this process: one handle, same calls, same answers */
    int sharded_ok = 1;
    const char *devs = getenv("CC_DEVICES");
    if (devs && *devs) {
        int ids[64], nd = 0;
        for (const char *p = devs; *p && nd < 64;) {            /* "0,1,2" */
            ids[nd++] = atoi(p);
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
        cc_sharded *sh = NULL;
        CHECK(cc_open_sharded(argv[1], ids, nd, &sh));
        uint8_t *records2 = malloc((n ? n : 1) * out_size);
        uint64_t *index2 = malloc((n ? n : 1) * sizeof(uint64_t));
        int64_t *found2 = malloc((novel ? novel : 1) * sizeof(int64_t));
        uint64_t novel2 = 0;
        CHECK(cc_find_novel_sharded(sh, child, parents, np, records2, index2, n, &novel2));
        CHECK(cc_find_packed_sharded(sh, words, NULL, novel, found2));
        sharded_ok = novel2 == novel && memcmp(records, records2, novel * out_size) == 0 && memcmp(index, index2, novel * 8) == 0 &&
                     memcmp(found, found2, novel * 8) == 0;
        cc_sharded_stats st;
        CHECK(cc_sharded_last_stats(sh, &st));
        printf("sharded over %d devices: novel %llu, lookups %s the single-GPU answers (%u kernel launches)\n", nd, (unsigned long long)novel2,
               sharded_ok ? "equal" : "DIFFER FROM", st.launches);
        free(records2); free(index2); free(found2);
        cc_dispose_sharded(sh);
    }
    free(records); free(index); free(words); free(found);
    cc_dispose(g);
    return ok == novel && sharded_ok ? 0 : 1;
}
