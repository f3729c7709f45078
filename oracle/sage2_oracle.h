/* sage2_oracle.h -- CPU restatement (plain C) of SAGE2 steps 1-3 (reference main.cpp:37-132).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this; the product (libsage2gpu) never links, imports or executes anything in oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the UNMODIFIED reference compiled here
 * (oracle/_ref/SAGE2, recipe in oracle/Makefile): tests/golden/ holds the md5 of the reference's
 * own `.reads` and `.graph3` for seeded inputs plus its log counters, and
 * tests/test_oracle_golden.py checks this file against them.
 */
#ifndef SAGE2_ORACLE_H
#define SAGE2_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t id;      /* extension read id (1-based), 0 = none        economyGraph.h:24-30 */
    uint32_t type;    /* 0 = same strand, 1 = reverse strand                                */
    uint32_t length;  /* overhang (22-bit field in the reference)                           */
} sgo_ext;

typedef struct {
    uint64_t from, to;   /* from < to, 1-based read ids                                     */
    uint32_t type;       /* edge type 0..3 (economyGraph.cpp:607-626)                       */
    uint32_t delta;      /* overhang of `to` beyond `from`  (Edge::lengthOfEdge)             */
    uint32_t delta_twin; /* overhang of the twin edge       (overlapGraph.cpp:147)           */
} sgo_edge;

typedef struct {
    /* step 1 (readLoader.cpp) */
    uint64_t total_reads, good_reads, unique_reads, total_bp, avg_len;
    uint16_t *length;      /* [U+1]  1-based                                               */
    uint16_t *frequency;   /* [U+1]                                                        */
    uint64_t *byte_off;    /* [U+2]  offsets into fwd/rc byte buffers                      */
    uint8_t  *fwd, *rc;    /* packed 2-bit, reference byte layout (utils.cpp:96-119)       */
    /* step 2 (hashTable.cpp) */
    uint64_t hash_len, distinct_keys, keys_over_threshold;
    /* step 3 phase A/B (economyGraph.cpp:37-490) */
    sgo_ext *right_ext, *left_ext;   /* [U+1] after phase A                                */
    uint8_t *explored_a;             /* [U+1] state after phase A (0,5,6)                  */
    uint8_t *explored_b;             /* [U+1] state after phase B (0,4,5,6)                */
    uint64_t compare_calls;          /* V of SURVEY 8(d): gated partner comparisons, phase A */
    uint64_t contained_ext, contained_size, left_to_explore;
    /* phase C (economyGraph.cpp:495-707) */
    uint64_t edges_inserted_c, transitive_removed;
    /* consumer view (overlapGraph.cpp:84-115,338-369) */
    uint64_t n_edges;
    sgo_edge *edges;                 /* in .graph3 order                                   */
} sgo_result;

/* bases: concatenated ASCII reads; offsets[n_reads+1].  Returns 0 on success. */
int  sgo_run(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int min_overlap,
             int n_threads, sgo_result *out);
void sgo_free(sgo_result *r);
/* Text writers in the reference's -s formats (readLoader.cpp:270-287, overlapGraph.cpp:338-369). */
int  sgo_write_reads(const sgo_result *r, const char *path);
int  sgo_write_graph3(const sgo_result *r, const char *path);

/* Step-6 mapping of reads to ids (ReadLoader::getIdOfRead, readLoader.cpp:319-353, gated by isGoodRead as in
 * matePair.cpp:176-179): ids[r] = +/- id or 0, good[r] (may be NULL) = isGoodRead.  Pinned against the reference's
 * own getIdOfRead (oracle/ref_mapids.cpp -> tests/golden/mapids.json). */
int  sgo_map_reads(const sgo_result *r, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                   int min_overlap, int64_t *ids, uint8_t *good);

/* Small known-answer entry points for unit tests (utils.cpp restatements). */
void     sgo_chars_to_bytes(const uint8_t *s, int len, uint8_t *out);
uint64_t sgo_get64(const uint8_t *read, int start, int length);
int      sgo_string_compare(const uint8_t *r1, int l1, const uint8_t *r2, int l2);

#ifdef __cplusplus
}
#endif
#endif
