/* sage2_oracle.c -- CPU restatement (plain C) of SAGE2 steps 1-3.
 *
 * TEST INFRASTRUCTURE ONLY (see sage2_oracle.h).  Every function names the reference lines it
 * restates; paths are relative to /root/reference.  The restatement keeps the reference's byte
 * layout and byte-wise arithmetic on purpose, so that it shares no bit tricks with the CUDA path
 * it checks (which works on word-big-endian 64-bit records).
 *
 * Deliberate, documented differences from the reference (none observable in its outputs):
 *  - the hash table's slot function / prime sizes (hashTable.cpp:233-254,303-314) are not
 *    restated: only "key -> list of (readId,type) in insertion order, lists that reach 100
 *    entries are invisible" is observable, so a power-of-two linear-probe
 *    table that stores the key is used;
 *  - per-call malloc/free of keys is gone;
 *  - the phase-A race between exploredReads[read2]=6 (any thread, economyGraph.cpp:735) and
 *    exploredReads[i]=5 (owner, :444) is resolved the way a 1-thread run of the reference
 *    resolves it: the write that happens later in ascending-i order wins.
 */
#define _GNU_SOURCE
#include "sage2_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------- */
/* utils.cpp                                                                                   */
/* ------------------------------------------------------------------------------------------- */

static int code_of(uint8_t c)
{
    /* utils.cpp:109-112 */
    if (c == 'A' || c == 'a') return 0;
    if (c == 'C' || c == 'c') return 1;
    if (c == 'G' || c == 'g') return 2;
    if (c == 'T' || c == 't') return 3;
    return 0;
}

/* utils.cpp:96-119 charsToBytes: base i -> byte i/4, bit offset 6-2(i%4); last byte left aligned */
void sgo_chars_to_bytes(const uint8_t *s, int len, uint8_t *out)
{
    int nbytes = (len + 3) / 4;
    int shift = 2 * (4 * nbytes - len);
    int j = 0;
    memset(out, 0, (size_t)nbytes);
    for (int i = 0; i < len; i++) {
        j = i / 4;
        out[j] = (uint8_t)((out[j] << 2) | code_of(s[i]));
    }
    if (len > 0) out[j] = (uint8_t)(out[j] << shift);
}

/* utils.cpp:124-137 bytesToChars */
static void bytes_to_chars(const uint8_t *b, int len, char *out)
{
    static const char T[4] = { 'A', 'C', 'G', 'T' };
    for (int i = 0; i < len; i++) out[i] = T[(b[i >> 2] >> (8 - 2 * (i % 4 + 1))) & 3];
}

/* utils.cpp:73-91 reverseComplement (on characters already upper-cased ACGT) */
static void reverse_complement(const char *s, int len, char *out)
{
    int j = 0;
    for (int i = len - 1; i >= 0; i--, j++) {
        char c = s[i];
        out[j] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
}

/* utils.cpp:144-166 isGoodRead: len > minOvlp, only ACGT (lower case folded to upper) */
static int is_good_read(char *s, int len, int min_ovlp)
{
    if (len <= min_ovlp) return 0;
    int i;
    for (i = 0; i < len; i++) {
        char c = s[i];
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') continue;
        if (c == 'a') s[i] = 'A';
        else if (c == 'c') s[i] = 'C';
        else if (c == 'g') s[i] = 'G';
        else if (c == 't') s[i] = 'T';
        else break;
    }
    return i == len;
}

/* utils.cpp:189-207 get64BitInt: bases [start,start+length) (length<=32) as a big-endian integer.
 * Buffers handed to this function carry one spare zero byte (reference quirk: it reads
 * read[(start+length)>>2] even when that is one past the end; the value is shifted out). */
uint64_t sgo_get64(const uint8_t *read, int start, int length)
{
    uint64_t number = 0;
    int byte, f1 = (start & 3) << 1, f2 = ((start + length) & 3) << 1;
    if ((start >> 2) == ((start + length) >> 2))
        return (uint64_t)((read[start >> 2] & (0xFF >> f1)) >> (8 - f2));
    for (byte = start >> 2; byte < ((start + length) >> 2); byte++) {
        if (byte == (start >> 2)) number = (uint64_t)(read[byte] & (0xFF >> f1));
        else number = (number << 8) | read[byte];
    }
    number = (number << f2) | (uint64_t)(read[byte] >> (8 - f2));
    return number;
}

/* utils.cpp:171-187 get64Bit2Int: v[1] = last min(len,32) bases, v[0] = leading len-32 bases */
static void get64x2(const uint8_t *read, int start, int length, uint64_t v[2])
{
    v[0] = 0; v[1] = 0;
    if (length <= 32) v[1] = sgo_get64(read, start, length);
    else {
        v[0] = sgo_get64(read, start, length - 32);
        v[1] = sgo_get64(read, start + length - 32, 32);
    }
}

/* utils.cpp:224-242 stringCompareInBytes */
int sgo_string_compare(const uint8_t *r1, int l1, const uint8_t *r2, int l2)
{
    int n1 = (l1 + 3) / 4, n2 = (l2 + 3) / 4;
    for (int i = 0; i < n1 && i < n2; i++) {
        if (r1[i] < r2[i]) return -1;
        if (r1[i] > r2[i]) return 1;
    }
    if (l1 < l2) return -1;
    if (l1 > l2) return 1;
    return 0;
}

/* utils.cpp:212-219 reverseEdgeType */
static uint32_t reverse_edge_type(uint32_t t) { return t == 0 ? 3u : t == 3 ? 0u : t; }

/* ------------------------------------------------------------------------------------------- */
/* step 1: readLoader.cpp                                                                      */
/* ------------------------------------------------------------------------------------------- */

typedef struct {
    const uint8_t *fwd, *rc;   /* packed, each with one spare byte */
    uint16_t len, freq;
} oread;

typedef struct { uint8_t *bytes; uint16_t len; } raw_read;

static int raw_cmp(const void *a, const void *b)
{
    const raw_read *x = (const raw_read *)a, *y = (const raw_read *)b;
    return sgo_string_compare(x->bytes, x->len, y->bytes, y->len);   /* readLoader.cpp:11-18 */
}

typedef struct {
    uint64_t U;
    oread *reads;          /* 1-based */
    int k, h;
} octx;

/* ------------------------------------------------------------------------------------------- */
/* step 2: hashTable.cpp                                                                       */
/* ------------------------------------------------------------------------------------------- */

typedef struct { uint64_t id; uint8_t type; } helem;      /* hashTable.h:13-18 */
typedef struct {
    uint64_t v0, v1;
    helem *e;              /* e[0..count) */
    uint32_t count;
    uint8_t used, masked;
} hslot;
typedef struct { hslot *s; uint64_t mask; uint64_t distinct, over; } htable;

static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static hslot *ht_find(const htable *t, const uint64_t v[2], int create)
{
    uint64_t p = mix64(v[1] ^ mix64(v[0] + 0x9e3779b97f4a7c15ULL)) & t->mask;
    for (;;) {
        hslot *s = &t->s[p];
        if (!s->used) {
            if (!create) return NULL;
            s->used = 1; s->v0 = v[0]; s->v1 = v[1];
            return s;
        }
        if (s->v0 == v[0] && s->v1 == v[1]) return s;
        p = (p + 1) & t->mask;
    }
}

/* hashTable.cpp:133-188 hashTableInsert: append while count <= hashThreshold(100) (:178) */
static void ht_insert(htable *t, const uint64_t v[2], uint64_t id, uint8_t type)
{
    hslot *s = ht_find(t, v, 1);
    if (s->count == 0) { t->distinct++; s->e = (helem *)malloc(2 * sizeof(helem)); }
    if (s->count <= 100) {
        s->e = (helem *)realloc(s->e, (s->count + 2) * sizeof(helem));
        s->e[s->count].id = id; s->e[s->count].type = type;
        s->count++;
    }
}

/* hashTable.cpp:70-128 hashPrefixesAndSuffix */
static void build_table(const octx *c, htable *t)
{
    uint64_t size = 16;
    while (size < 8 * c->U) size <<= 1;
    t->s = (hslot *)calloc(size, sizeof(hslot));
    t->mask = size - 1; t->distinct = 0; t->over = 0;
    int h = c->h;
    for (uint64_t i = 1; i <= c->U; i++) {                       /* :94-109, serial, id ascending */
        const oread *r = &c->reads[i];
        uint64_t v[2];
        get64x2(r->fwd, 0, h, v);          ht_insert(t, v, i, 0);
        get64x2(r->fwd, r->len - h, h, v); ht_insert(t, v, i, 1);
        get64x2(r->rc, 0, h, v);           ht_insert(t, v, i, 2);
        get64x2(r->rc, r->len - h, h, v);  ht_insert(t, v, i, 3);
    }
    for (uint64_t p = 0; p < size; p++)                          /* :111-123 mask count >= 100 */
        if (t->s[p].used && t->s[p].count >= 100) { t->s[p].masked = 1; t->over++; }
}

/* hashTable.cpp:193-231 hashTableSearch: masked buckets are skipped => invisible */
static const hslot *ht_search(const htable *t, const uint64_t v[2])
{
    const hslot *s = ht_find(t, v, 0);
    if (s == NULL || s->masked) return NULL;
    return s;
}

static void free_table(htable *t)
{
    for (uint64_t p = 0; p <= t->mask; p++) free(t->s[p].e);
    free(t->s);
}

/* ------------------------------------------------------------------------------------------- */
/* step 3: economyGraph.cpp                                                                    */
/* ------------------------------------------------------------------------------------------- */

/* economyGraph.cpp:712-758 compareStringInBytes / :763-808 compareStringInBytesPrevious.
 * Returns 1 = X's remainder matches (overlap), 2 = Y's remainder fits inside X and matches
 * (containment), 0 = mismatch.  ...InBytes maps {1->1, 2->mark read2 contained and return 0};
 * ...Previous maps {1->1, 2->1}. */
static int compare_core(const uint8_t *x, const uint8_t *y, int start, int len1, int len2, int h)
{
    int start1 = start + h, start2 = h, length;
    uint64_t a[2], b[2];
    if (len2 - start2 <= len1 - start1) {
        while (start2 < len2) {
            length = len2 - start2 < 64 ? len2 - start2 : 64;
            get64x2(x, start1, length, a);
            get64x2(y, start2, length, b);
            if (a[0] != b[0] || a[1] != b[1]) return 0;
            start1 += length; start2 += length;
        }
        return 2;
    }
    while (start1 < len1) {
        length = len1 - start1 < 64 ? len1 - start1 : 64;
        get64x2(x, start1, length, a);
        get64x2(y, start2, length, b);
        if (a[0] != b[0] || a[1] != b[1]) return 0;
        start1 += length; start2 += length;
    }
    return 1;
}

typedef struct { uint64_t id; uint32_t type, mark, length; } eedge;    /* economyGraph.h:14-22 */
typedef struct { eedge *e; uint32_t n, cap; } elist;

/* economyGraph.cpp:813-849 insertEdgeEconomy */
static int insert_edge_economy(const octx *c, elist *g, uint64_t u, uint64_t v, uint32_t delta,
                               uint32_t type)
{
    if (u == v) return 0;
    uint32_t lu = c->reads[u].len, lv = c->reads[v].len;
    uint32_t delta2 = lu - (lv - delta);                       /* :822 (32-bit wrap as in C++) */
    eedge uv = { v, type, 0, delta & 0xFFFFF };                /* 20-bit length field           */
    eedge vu = { u, reverse_edge_type(type), 0, delta2 & 0xFFFFF };
    elist *a = &g[u], *b = &g[v];
    if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 2; a->e = (eedge *)realloc(a->e, a->cap * sizeof(eedge)); }
    a->e[a->n++] = uv;
    if (b->n == b->cap) { b->cap = b->cap ? 2 * b->cap : 2; b->e = (eedge *)realloc(b->e, b->cap * sizeof(eedge)); }
    b->e[b->n++] = vu;
    return 1;
}

/* economyGraph.cpp:853-871 compareLengthBased (descending length, id, type) */
static int cmp_length_based(const void *pa, const void *pb)
{
    const eedge *a = (const eedge *)pa, *b = (const eedge *)pb;
    if (a->length != b->length) return a->length > b->length ? -1 : 1;
    if (a->id != b->id) return a->id > b->id ? -1 : 1;
    if (a->type != b->type) return a->type > b->type ? -1 : 1;
    return 0;
}

/* economyGraph.cpp:875-893 compareIdBased (ascending id, type, length) */
static int cmp_id_based(const void *pa, const void *pb)
{
    const eedge *a = (const eedge *)pa, *b = (const eedge *)pb;
    if (a->id != b->id) return a->id < b->id ? -1 : 1;
    if (a->type != b->type) return a->type < b->type ? -1 : 1;
    if (a->length != b->length) return a->length < b->length ? -1 : 1;
    return 0;
}

/* Phase A, economyGraph.cpp:64-452, one read.  `cont_max[x]` records the largest i whose scan
 * marked x contained (the time-ordered last writer of exploredReads[x]=6 in a 1-thread run). */
static void phase_a_read(const octx *c, const htable *t, uint64_t i, sgo_ext *R, sgo_ext *L,
                         uint8_t *flag5, uint64_t *cont_max, uint64_t *compare_calls)
{
    const oread *r1 = &c->reads[i];
    const int h = c->h, k = c->k, len1 = r1->len;
    uint64_t prevIDRight = 0, prevIDLeft = 0, connections = 0, calls = 0;
    int markAmbigRight = 0, markAmbigLeft = 0, itsAmbigRight = 0, itsAmbigLeft = 0;
    int prevTypeRight = 0, prevLengthRight = 0, prevTypeLeft = 0, prevLengthLeft = 0, markFirstRight = 0;
    sgo_ext right = { 0, 0, 0 }, left = { 0, 0, 0 };

    for (int j = 0; j <= len1 - h; j++) {                                         /* :77 */
        uint64_t key[2];
        get64x2(r1->fwd, j, h, key);
        const hslot *s = ht_search(t, key);
        if (s == NULL) continue;
        markAmbigRight = 0; markAmbigLeft = 0; markFirstRight = 0;                /* :86-88 */
        for (uint32_t q = 0; q < s->count; q++) {
            uint64_t read2 = s->e[q].id;
            int type = s->e[q].type;
            const oread *r2 = &c->reads[read2];
            const int len2 = r2->len;
            if (read2 == i) continue;
            if (type == 0 || type == 2) {                                         /* :94, :187 */
                if (!(j <= (int)(uint16_t)(len1 - k))) continue;
                const uint8_t *y = type == 0 ? r2->fwd : r2->rc;
                int t01 = type == 0 ? 0 : 1;
                calls++;
                int res = compare_core(r1->fwd, y, j, len1, len2, h);
                if (res == 2) {                                                    /* :735 */
                    uint64_t old = __atomic_load_n(&cont_max[read2], __ATOMIC_RELAXED);
                    while (old < i && !__atomic_compare_exchange_n(&cont_max[read2], &old, i, 0,
                                                                   __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
                }
                if (res != 1) continue;
                connections++;
                if (right.id == 0) {                                               /* :97-107 */
                    right.id = read2; right.type = (uint32_t)t01;
                    right.length = (uint32_t)(len2 - (len1 - j));
                    prevIDRight = read2; prevTypeRight = t01; prevLengthRight = j;
                    markAmbigRight = 1; markFirstRight = 1;
                } else {                                                           /* :108-184 */
                    const oread *pr = &c->reads[prevIDRight];
                    const uint8_t *p = prevTypeRight == 0 ? pr->fwd : pr->rc;
                    if (compare_core(p, y, j - prevLengthRight, pr->len, len2, h)) {
                        if (markAmbigRight == 1) {
                            if (len2 > pr->len) {
                                if (markFirstRight == 1) {
                                    right.id = read2; right.type = (uint32_t)t01;
                                    right.length = (uint32_t)(len2 - (len1 - j));
                                }
                                prevIDRight = read2; prevTypeRight = t01; prevLengthRight = j;
                            }
                        } else {
                            prevIDRight = read2; prevTypeRight = t01; prevLengthRight = j;
                            markAmbigRight = 1;
                        }
                    } else itsAmbigRight = 1;
                }
            } else {                                                               /* :279, :359 */
                if (!(j >= (int)(uint16_t)(k - h))) continue;
                const uint8_t *y = type == 1 ? r2->rc : r2->fwd;
                int t01 = type == 1 ? 0 : 1;
                int p1 = len1 - j - h;
                calls++;
                int res = compare_core(r1->rc, y, p1, len1, len2, h);
                if (res == 2) {
                    uint64_t old = __atomic_load_n(&cont_max[read2], __ATOMIC_RELAXED);
                    while (old < i && !__atomic_compare_exchange_n(&cont_max[read2], &old, i, 0,
                                                                   __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
                }
                if (res != 1) continue;
                connections++;
                if (left.id == 0) {                                                /* :282-291 */
                    left.id = read2; left.type = (uint32_t)t01;
                    left.length = (uint32_t)(len2 - j - h);
                    prevIDLeft = read2; prevTypeLeft = t01; prevLengthLeft = p1;
                    markAmbigLeft = 1;
                } else {                                                           /* :292-356 */
                    const oread *pr = &c->reads[prevIDLeft];
                    const uint8_t *p = prevTypeLeft == 0 ? pr->rc : pr->fwd;
                    if (compare_core(y, p, prevLengthLeft - p1, len2, pr->len, h)) {
                        if (markAmbigLeft == 1) {
                            if (len2 > pr->len) {
                                left.id = read2; left.type = (uint32_t)t01;
                                left.length = (uint32_t)(len2 - j - h);
                                prevIDLeft = read2; prevTypeLeft = t01; prevLengthLeft = p1;
                            }
                        } else {
                            left.id = read2; left.type = (uint32_t)t01;
                            left.length = (uint32_t)(len2 - j - h);
                            prevIDLeft = read2; prevTypeLeft = t01; prevLengthLeft = p1;
                            markAmbigLeft = 1;
                        }
                    } else itsAmbigLeft = 1;
                }
            }
        }
    }
    if (connections > 300) flag5[i] = 1;                                           /* :443 */
    if (itsAmbigRight == 1 || itsAmbigLeft == 1) { left.length = 0; right.length = 0; } /* :446 */
    R[i] = right; L[i] = left;
    __atomic_fetch_add(compare_calls, calls, __ATOMIC_RELAXED);
}

/* insertAllEdgesOfRead, economyGraph.cpp:580-638 */
static uint64_t insert_all_edges_of_read(const octx *c, const htable *t, elist *g,
                                         uint8_t *explored, uint64_t read1)
{
    uint64_t inserted = 0;
    if (explored[read1] != 0) return 0;
    explored[read1] = 1;
    const oread *r1 = &c->reads[read1];
    const int h = c->h, k = c->k, len1 = r1->len;
    for (int j = 0; j <= (int)(uint16_t)(len1 - h); j++) {
        uint64_t key[2];
        get64x2(r1->fwd, j, h, key);
        const hslot *s = ht_search(t, key);
        if (s == NULL) continue;
        for (uint32_t q = 0; q < s->count; q++) {
            uint64_t read2 = s->e[q].id;
            int type = s->e[q].type;
            const oread *r2 = &c->reads[read2];
            const int len2 = r2->len;
            int32_t ovlp = -1; int etype = -1;
            if (explored[read2]) continue;                                         /* :605 */
            if (type == 0 && j <= (int)(uint16_t)(len1 - k) && compare_core(r1->fwd, r2->fwd, j, len1, len2, h)) {
                ovlp = len2 - (len1 - j); etype = 3;
            } else if (type == 1 && j >= (int)(uint16_t)(k - h) && compare_core(r1->rc, r2->rc, len1 - j - h, len1, len2, h)) {
                ovlp = len2 - j - h; etype = 0;
            } else if (type == 2 && j <= (int)(uint16_t)(len1 - k) && compare_core(r1->fwd, r2->rc, j, len1, len2, h)) {
                ovlp = len2 - (len1 - j); etype = 2;
            } else if (type == 3 && j >= (int)(uint16_t)(k - h) && compare_core(r1->rc, r2->fwd, len1 - j - h, len1, len2, h)) {
                ovlp = len2 - j - h; etype = 1;
            }
            if (ovlp != -1)                                                        /* :627 */
                inserted += (uint64_t)insert_edge_economy(c, g, read1, read2, (uint32_t)ovlp, (uint32_t)etype);
        }
    }
    if (g[read1].n > 1) qsort(g[read1].e, g[read1].n, sizeof(eedge), cmp_length_based);   /* :634 */
    return inserted * 2;
}

/* markTransitiveEdge, economyGraph.cpp:643-679 */
static void mark_transitive_edge(elist *g, uint8_t *marked, uint8_t *explored, uint64_t from)
{
    elist *lf = &g[from];
    for (uint32_t i = 0; i < lf->n; i++) marked[lf->e[i].id] = 1;
    for (uint32_t i = 0; i < lf->n; i++) {
        uint64_t a = lf->e[i].id;
        if (marked[a] != 1) continue;
        elist *la = &g[a];
        for (uint32_t j = 0; j < la->n; j++) {
            uint64_t b = la->e[j].id;
            if (marked[b] != 1) continue;
            uint32_t t1 = lf->e[i].type, t2 = la->e[j].type;
            if ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) marked[b] = 2;
            else if ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3)) marked[b] = 2;
        }
    }
    for (uint32_t i = 0; i < lf->n; i++) if (marked[lf->e[i].id] == 2) lf->e[i].mark = 1;
    for (uint32_t i = 0; i < lf->n; i++) marked[lf->e[i].id] = 0;
    marked[from] = 0;
    explored[from] = 2;
}

/* removeTransitiveEdges, economyGraph.cpp:681-707 */
static uint64_t remove_transitive_edges(elist *g, uint64_t read)
{
    elist *l = &g[read];
    uint32_t n = 0;
    for (uint32_t i = 0; i < l->n; i++) if (!l->e[i].mark) l->e[n++] = l->e[i];
    uint64_t removed = l->n - n;
    l->n = n;
    return removed;
}

/* ------------------------------------------------------------------------------------------- */
/* driver                                                                                      */
/* ------------------------------------------------------------------------------------------- */

int sgo_run(const uint8_t *bases, const int64_t *offsets, int64_t n_reads, int min_overlap,
            int n_threads, sgo_result *out)
{
    memset(out, 0, sizeof(*out));
    const int k = min_overlap;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif

    /* ---- step 1a: readDatasetInBytes + insertReadIntoList (readLoader.cpp:133-213) ---------- */
    raw_read *raw = (raw_read *)malloc((size_t)(n_reads + 1) * sizeof(raw_read));
    uint64_t good = 0, total_bp = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        int len = (int)(offsets[r + 1] - offsets[r]);
        if (len <= k || len > 65535) continue;                   /* :149 smallRead */
        char *s = (char *)malloc((size_t)len * 2);
        char *rcs = s + len;
        memcpy(s, bases + offsets[r], (size_t)len);
        if (!is_good_read(s, len, k)) { free(s); continue; }
        reverse_complement(s, len, rcs);
        const char *keep = memcmp(s, rcs, (size_t)len) < 0 ? s : rcs;   /* :195 read.compare(rc)<0 */
        uint8_t *b = (uint8_t *)calloc((size_t)(len + 3) / 4 + 1, 1);
        sgo_chars_to_bytes((const uint8_t *)keep, len, b);
        raw[good].bytes = b; raw[good].len = (uint16_t)len;
        good++; total_bp += (uint64_t)len;
        free(s);
    }
    out->total_reads = (uint64_t)n_reads;
    out->good_reads = good;
    out->total_bp = total_bp;
    out->avg_len = good ? total_bp / good : 0;                   /* :161 integer division */

    /* ---- step 1b: organizeReads (readLoader.cpp:215-260) ----------------------------------- */
    qsort(raw, good, sizeof(raw_read), raw_cmp);                 /* :221 */
    uint64_t U = 0;
    oread *reads = (oread *)calloc(good + 2, sizeof(oread));
    for (uint64_t i = 0; i < good; i++) {                        /* :225-235 */
        int same = 0;
        if (U > 0) same = sgo_string_compare(reads[U].fwd, reads[U].len, raw[i].bytes, raw[i].len) == 0;
        if (!same) { U++; reads[U].fwd = raw[i].bytes; reads[U].len = raw[i].len; reads[U].freq = 0; }
        else free(raw[i].bytes);
        reads[U].freq++;                                         /* uint16_t, wraps like the reference */
    }
    free(raw);
    for (uint64_t i = 1; i <= U; i++) {                          /* :249-255 packed revcomp */
        int len = reads[i].len;
        char *s = (char *)malloc((size_t)len * 2);
        bytes_to_chars(reads[i].fwd, len, s);
        reverse_complement(s, len, s + len);
        uint8_t *b = (uint8_t *)calloc((size_t)(len + 3) / 4 + 1, 1);
        sgo_chars_to_bytes((const uint8_t *)(s + len), len, b);
        reads[i].rc = b;
        free(s);
    }
    out->unique_reads = U;
    out->length = (uint16_t *)calloc(U + 1, sizeof(uint16_t));
    out->frequency = (uint16_t *)calloc(U + 1, sizeof(uint16_t));
    out->byte_off = (uint64_t *)calloc(U + 2, sizeof(uint64_t));
    for (uint64_t i = 1; i <= U; i++) {
        out->length[i] = reads[i].len; out->frequency[i] = reads[i].freq;
        out->byte_off[i + 1] = out->byte_off[i] + (uint64_t)(reads[i].len + 3) / 4;
    }
    out->fwd = (uint8_t *)malloc(out->byte_off[U + 1] + 1);
    out->rc = (uint8_t *)malloc(out->byte_off[U + 1] + 1);
    for (uint64_t i = 1; i <= U; i++) {
        memcpy(out->fwd + out->byte_off[i], reads[i].fwd, (size_t)(reads[i].len + 3) / 4);
        memcpy(out->rc + out->byte_off[i], reads[i].rc, (size_t)(reads[i].len + 3) / 4);
    }

    octx c = { U, reads, k, k > 64 ? 64 : k };                   /* hashTable.cpp:78-81 */
    out->hash_len = (uint64_t)c.h;

    /* ---- step 2 --------------------------------------------------------------------------- */
    htable t;
    build_table(&c, &t);
    out->distinct_keys = t.distinct;
    out->keys_over_threshold = t.over;

    /* ---- step 3 phase A (economyGraph.cpp:37-452) ------------------------------------------ */
    sgo_ext *R = (sgo_ext *)calloc(U + 1, sizeof(sgo_ext));
    sgo_ext *L = (sgo_ext *)calloc(U + 1, sizeof(sgo_ext));
    uint8_t *flag5 = (uint8_t *)calloc(U + 1, 1);
    uint64_t *cont_max = (uint64_t *)calloc(U + 1, sizeof(uint64_t));
    uint8_t *explored = (uint8_t *)calloc(U + 1, 1);
    uint64_t calls = 0;
    #pragma omp parallel for schedule(dynamic, 256)
    for (uint64_t i = 1; i <= U; i++)
        phase_a_read(&c, &t, i, R, L, flag5, cont_max, &calls);
    for (uint64_t i = 1; i <= U; i++) {
        /* 1-thread time order: 6 is written during iteration cont_max[i], 5 at the end of i */
        if (cont_max[i] && flag5[i]) explored[i] = (i >= cont_max[i]) ? 5 : 6;
        else if (cont_max[i]) explored[i] = 6;
        else if (flag5[i]) explored[i] = 5;
    }
    out->compare_calls = calls;
    out->right_ext = R; out->left_ext = L;
    out->explored_a = (uint8_t *)malloc(U + 1);
    memcpy(out->explored_a, explored, U + 1);
    free(flag5); free(cont_max);

    /* ---- phase B (economyGraph.cpp:455-480) ------------------------------------------------ */
    elist *g = (elist *)calloc(U + 1, sizeof(elist));
    uint64_t contained = 0, contained_size = 0;
    for (uint64_t i = 1; i <= U; i++) {
        if (explored[i] != 6) {
            if ((L[i].length != 0 && (R[L[i].id].id == i || L[L[i].id].id == i)) &&
                (R[i].length != 0 && (R[R[i].id].id == i || L[R[i].id].id == i))) {
                if (explored[L[i].id] != 4)
                    insert_edge_economy(&c, g, i, L[i].id, L[i].length, L[i].type == 0 ? 0 : 1);
                if (explored[R[i].id] != 4)
                    insert_edge_economy(&c, g, i, R[i].id, R[i].length, R[i].type == 0 ? 3 : 2);
                contained++;
                explored[i] = 4;
            }
        } else contained_size++;
    }
    out->contained_ext = contained; out->contained_size = contained_size;
    out->left_to_explore = U - contained - contained_size;
    out->explored_b = (uint8_t *)malloc(U + 1);
    memcpy(out->explored_b, explored, U + 1);

    /* ---- phase C (economyGraph.cpp:495-574) ------------------------------------------------ */
    uint8_t *marked = (uint8_t *)calloc(U + 1, 1);
    uint64_t *queue = (uint64_t *)malloc((U + 1) * sizeof(uint64_t));
    uint64_t inserted = 0, removed = 0;
    for (uint64_t i = 1; i <= U; i++) {
        if (explored[i] != 0) continue;
        uint64_t start = 0, end = 0;
        queue[end++] = i;
        while (start < end) {
            uint64_t read1 = queue[start++];
            if (explored[read1] == 0) inserted += insert_all_edges_of_read(&c, &t, g, explored, read1);
            if (g[read1].e == NULL) continue;                    /* :525 economyGraphList[read1]!=NULL */
            if (explored[read1] == 1) {
                for (uint32_t x = 0; x < g[read1].n; x++) {
                    uint64_t read2 = g[read1].e[x].id;
                    if (explored[read2] == 0) {
                        queue[end++] = read2;
                        inserted += insert_all_edges_of_read(&c, &t, g, explored, read2);
                    }
                }
                mark_transitive_edge(g, marked, explored, read1);
            }
            if (explored[read1] == 2) {
                for (uint32_t x = 0; x < g[read1].n; x++) {
                    uint64_t read2 = g[read1].e[x].id;
                    if (explored[read2] == 1) {
                        for (uint32_t y = 0; y < g[read2].n; y++) {
                            uint64_t read3 = g[read2].e[y].id;
                            if (explored[read3] == 0) {
                                queue[end++] = read3;
                                inserted += insert_all_edges_of_read(&c, &t, g, explored, read3);
                            }
                        }
                        mark_transitive_edge(g, marked, explored, read2);
                    }
                }
                removed += remove_transitive_edges(g, read1);
            }
        }
    }
    out->edges_inserted_c = inserted; out->transitive_removed = removed;
    free(marked); free(queue); free(explored);

    /* ---- sortEconomyGraph (:896-913) + convertGraph (overlapGraph.cpp:84-115) -------------- */
    uint64_t n_edges = 0, cap = U + 16;
    sgo_edge *edges = (sgo_edge *)malloc(cap * sizeof(sgo_edge));
    for (uint64_t i = 1; i <= U; i++) {
        if (g[i].e == NULL) continue;
        if (g[i].n > 1) qsort(g[i].e, g[i].n, sizeof(eedge), cmp_id_based);
        for (uint32_t j = 0; j < g[i].n; j++) {
            const eedge *e = &g[i].e[j];
            if (j > 0 && g[i].e[j - 1].id == e->id && g[i].e[j - 1].type == e->type) continue;   /* :101 */
            if (i < e->id) {
                if (n_edges == cap) { cap *= 2; edges = (sgo_edge *)realloc(edges, cap * sizeof(sgo_edge)); }
                uint32_t ul = reads[i].len, vl = reads[e->id].len;
                edges[n_edges].from = i; edges[n_edges].to = e->id; edges[n_edges].type = e->type;
                edges[n_edges].delta = e->length;
                edges[n_edges].delta_twin = ul - (vl - e->length);           /* overlapGraph.cpp:147 */
                n_edges++;
            }
        }
    }
    out->n_edges = n_edges; out->edges = edges;

    for (uint64_t i = 1; i <= U; i++) { free(g[i].e); free((void *)reads[i].fwd); free((void *)reads[i].rc); }
    free(g); free(reads);
    free_table(&t);
    return 0;
}

void sgo_free(sgo_result *r)
{
    free(r->length); free(r->frequency); free(r->byte_off); free(r->fwd); free(r->rc);
    free(r->right_ext); free(r->left_ext); free(r->explored_a); free(r->explored_b); free(r->edges);
    memset(r, 0, sizeof(*r));
}

/* readLoader.cpp:29-36,270-287: U, then "freq\tlen\tfwd\trc" per read */
int sgo_write_reads(const sgo_result *r, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    fprintf(f, "%llu\n", (unsigned long long)r->unique_reads);
    char *buf = (char *)malloc(2 * 65536 + 8);
    for (uint64_t i = 1; i <= r->unique_reads; i++) {
        int len = r->length[i];
        bytes_to_chars(r->fwd + r->byte_off[i], len, buf);
        buf[len] = '\t';
        bytes_to_chars(r->rc + r->byte_off[i], len, buf + len + 1);
        buf[2 * len + 1] = '\n';
        fprintf(f, "%u\t%u\t", (unsigned)r->frequency[i], (unsigned)len);
        fwrite(buf, 1, (size_t)(2 * len + 2), f);
    }
    free(buf);
    return fclose(f);
}

/* overlapGraph.cpp:12-20,338-369: genomeSize(0), numberOfReads, averageReadLength, then per edge
 * "from\tto\ttype\t1\tdelta\t0\t0\n\n" followed by its twin. */
int sgo_write_graph3(const sgo_result *r, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    fprintf(f, "0\n%llu\n%llu\n", (unsigned long long)r->good_reads, (unsigned long long)r->avg_len);
    for (uint64_t e = 0; e < r->n_edges; e++) {
        const sgo_edge *x = &r->edges[e];
        fprintf(f, "%llu\t%llu\t%u\t1\t%u\t0\t0\n\n", (unsigned long long)x->from,
                (unsigned long long)x->to, x->type, x->delta);
        fprintf(f, "%llu\t%llu\t%u\t1\t%u\t0\t0\n\n", (unsigned long long)x->to,
                (unsigned long long)x->from, reverse_edge_type(x->type), x->delta_twin);
    }
    return fclose(f);
}

/* ------------------------------------------------------------------------------------------- */
/* step 6 mapping of mates to read ids: ReadLoader::getIdOfRead (readLoader.cpp:319-353) behind  */
/* the gate of MatePair::processMatePairs (matePair.cpp:176-179).                                */
/* ids[r] = +id when the read itself is the stored orientation (read < revcomp, strictly),       */
/*          -id when its reverse complement is (ties included), 0 when it is not in the list;    */
/* good[r] = isGoodRead (a bad read is never looked up: ids[r] = 0).                             */
/* ------------------------------------------------------------------------------------------- */
int sgo_map_reads(const sgo_result *r, const uint8_t *bases, const int64_t *offsets, int64_t n_reads,
                  int min_overlap, int64_t *ids, uint8_t *good)
{
    for (int64_t q = 0; q < n_reads; q++) {
        const int len = (int)(offsets[q + 1] - offsets[q]);
        ids[q] = 0;
        if (good) good[q] = 0;
        char *s = (char *)malloc((size_t)2 * (size_t)(len > 0 ? len : 1) + 2);
        uint8_t *b = (uint8_t *)calloc((size_t)(len + 3) / 4 + 2, 1);
        if (!s || !b) { free(s); free(b); return 1; }
        memcpy(s, bases + offsets[q], (size_t)len);
        if (is_good_read(s, len, min_overlap)) {
            if (good) good[q] = 1;
            char *rcs = s + len;
            reverse_complement(s, len, rcs);
            int flag;
            if (strncmp(s, rcs, (size_t)len) < 0) { sgo_chars_to_bytes((const uint8_t *)s, len, b); flag = 1; }      /* :325-329 */
            else { sgo_chars_to_bytes((const uint8_t *)rcs, len, b); flag = -1; }                                      /* :330-334 */
            int64_t lb = 1, ub = (int64_t)r->unique_reads;
            while (lb <= ub) {                                                                                        /* :335-348 */
                const int64_t mid = (ub + lb) >> 1;
                const int c = sgo_string_compare(b, len, r->fwd + r->byte_off[mid], r->length[mid]);
                if (c == 0) { ids[q] = mid * flag; break; }
                if (c > 0) lb = mid + 1; else ub = mid - 1;
            }
        }
        free(s); free(b);
    }
    return 0;
}
