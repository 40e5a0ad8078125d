/*
 * TEST INFRASTRUCTURE -- C restatement of the BULK counting loop of te_counter (oracle).
 *
 * Only tests/, __graft_entry__ and bench.py's parity / cpu legs may load this library, and only as
 * the checker.  It exists so that a full-size result (hundreds of millions of records) can be checked
 * bit for bit in seconds; oracle/te_oracle.py is the readable restatement it is itself checked
 * against (tests/test_oracle_c.py), and that one is pinned to the unmodified reference by
 * tests/golden.
 *
 * It follows the reference literally and NOT the closed forms of the CUDA kernels:
 *   bucket hash            miniglbase/genelist.py:367-380  feature n is listed in every bucket of
 *                          range((L//bs)*bs, ((R+bs)//bs)*bs, bs)
 *   candidate gather       te_count.py:106-116 / :222-231  the union of exactly two buckets,
 *                          (loc1-1)//bs and (loc2+1)//bs
 *   point tests            te_count.py:118-126 / :233-241
 *   type rule and tally    te_count.py:128-149 / :243-261
 *   filter, mate merge     te_count.py:76-102 / :203-218
 * Units are independent, so they are split over threads (pthreads); counters are per thread.
 *
 * build: oracle/Makefile  ->  oracle/libteoracle.so
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T_GENE 1
#define T_TE 2
#define T_SNRNA 3
#define T_ENH 4
#define F_REJECT 7u          /* unmapped | dup | qcfail */
#define F_NAME_MISMATCH 16u

typedef struct {
    int n_chrom, n_ensg, bs;
    int64_t n_feat;
    const int32_t *L, *R, *ensg;
    const uint8_t* type;
    /* per chromosome: bucket numbers 0 .. n_buck[c]-1, CSR lists of feature ids; has[c] = chromosome is a key */
    int64_t* buck_off;       /* offsets into list_off, per chromosome */
    int64_t* n_buck;
    int64_t* list_off;       /* per bucket: offset into ids */
    int32_t* ids;
    uint8_t* has;
} teo_index;

static int64_t floordiv64(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

void* teo_index_build(int64_t n_feat, const int32_t* chrom_id, const int32_t* L, const int32_t* R, const int32_t* ensg,
                      const uint8_t* type, int n_chrom, int n_ensg, int bs) {
    teo_index* ix = (teo_index*)calloc(1, sizeof(teo_index));
    ix->n_chrom = n_chrom; ix->n_ensg = n_ensg; ix->bs = bs; ix->n_feat = n_feat;
    ix->L = L; ix->R = R; ix->ensg = ensg; ix->type = type;
    ix->n_buck = (int64_t*)calloc((size_t)n_chrom + 1, 8);
    ix->buck_off = (int64_t*)calloc((size_t)n_chrom + 2, 8);
    ix->has = (uint8_t*)calloc((size_t)n_chrom + 1, 1);
    for (int64_t n = 0; n < n_feat; ++n) {
        const int c = chrom_id[n];
        const int64_t lb = floordiv64(L[n], bs), rb = floordiv64((int64_t)R[n] + bs, bs);   /* range(lb*bs, rb*bs, bs) */
        ix->has[c] = 1;                          /* genelist.py:367-368: the key exists for every feature row */
        if (rb > lb && rb > ix->n_buck[c]) ix->n_buck[c] = rb;
    }
    for (int c = 0; c < n_chrom; ++c) ix->buck_off[c + 1] = ix->buck_off[c] + ix->n_buck[c];
    const int64_t nb = ix->buck_off[n_chrom];
    ix->list_off = (int64_t*)calloc((size_t)nb + 2, 8);
    for (int64_t n = 0; n < n_feat; ++n) {
        const int c = chrom_id[n];
        const int64_t lb = floordiv64(L[n], bs), rb = floordiv64((int64_t)R[n] + bs, bs);
        for (int64_t b = lb; b < rb; ++b) if (b >= 0) ix->list_off[ix->buck_off[c] + b + 1]++;
    }
    for (int64_t b = 0; b < nb; ++b) ix->list_off[b + 1] += ix->list_off[b];
    ix->ids = (int32_t*)malloc((size_t)(ix->list_off[nb] + 1) * 4);
    int64_t* fill = (int64_t*)malloc((size_t)(nb + 1) * 8);
    memcpy(fill, ix->list_off, (size_t)(nb + 1) * 8);
    for (int64_t n = 0; n < n_feat; ++n) {       /* append in linearData order, as the reference does */
        const int c = chrom_id[n];
        const int64_t lb = floordiv64(L[n], bs), rb = floordiv64((int64_t)R[n] + bs, bs);
        for (int64_t b = lb; b < rb; ++b) if (b >= 0) ix->ids[fill[ix->buck_off[c] + b]++] = (int32_t)n;
    }
    free(fill);
    return ix;
}

void teo_index_free(void* p) {
    teo_index* ix = (teo_index*)p;
    if (!ix) return;
    free(ix->n_buck); free(ix->buck_off); free(ix->list_off); free(ix->ids); free(ix->has); free(ix);
}

typedef struct {
    const teo_index* ix;
    int paired, qual;
    int64_t u0, u1;                          /* units [u0, u1) */
    const int32_t *start, *end;
    const uint16_t* chrom;
    const uint8_t *mapq, *flag;
    int64_t* counts;                         /* n_ensg, this thread's */
    int64_t stats[6];                        /* assigned, lowq, badchrom, qcfail, crash_enhancer, crash_name */
    uint32_t* estamp;                        /* n_ensg: distinct ensg of one unit */
} teo_job;

static void* teo_worker(void* arg) {
    teo_job* j = (teo_job*)arg;
    const teo_index* ix = j->ix;
    const int bs = ix->bs;
    uint32_t tick = 0;
    int32_t hits[4096];
    for (int64_t u = j->u0; u < j->u1; ++u) {
        int64_t c, loc1, loc2;
        if (j->paired) {
            const int64_t r1 = 2 * u, r2 = 2 * u + 1;
            if (j->flag[r1] & F_REJECT) { j->stats[3]++; continue; }            /* :81 */
            if (j->flag[r2] & F_REJECT) { j->stats[3]++; continue; }            /* :84 */
            if ((int)j->mapq[r1] < j->qual) { j->stats[1]++; continue; }        /* :88 */
            if (j->flag[r1] & F_NAME_MISMATCH) { j->stats[5]++; continue; }     /* :92-94 */
            c = j->chrom[r1]; loc1 = j->start[r1]; loc2 = j->start[r2];         /* :96-98 */
        } else {
            if (j->flag[u] & F_REJECT) { j->stats[3]++; continue; }             /* :204 */
            if ((int)j->mapq[u] < j->qual) { j->stats[1]++; continue; }         /* :208 */
            c = j->chrom[u]; loc1 = j->start[u]; loc2 = j->end[u];              /* :212-214 */
        }
        if (c >= ix->n_chrom || !ix->has[c]) { j->stats[2]++; continue; }       /* :100 / :216 */
        const int64_t b1 = floordiv64(loc1 - 1, bs), b2 = floordiv64(loc2 + 1, bs);   /* :106-107 */
        if (++tick == 0) { memset(j->estamp, 0, (size_t)ix->n_ensg * 4); tick = 1; }
        int nh = 0, overflow = 0;
        unsigned types = 0;
        /* loc_ids = set union of the two buckets' lists; the lists hold feature indices in ascending order
         * (appended in linearData order), so the union is a merge walk */
        const int64_t *pa = 0, *pb = 0;
        int64_t ta = 0, ea = 0, tb = 0, eb = 0;
        (void)pa; (void)pb;
        if (b1 >= 0 && b1 < ix->n_buck[c]) { ta = ix->list_off[ix->buck_off[c] + b1]; ea = ix->list_off[ix->buck_off[c] + b1 + 1]; }
        if (b2 != b1 && b2 >= 0 && b2 < ix->n_buck[c]) { tb = ix->list_off[ix->buck_off[c] + b2]; eb = ix->list_off[ix->buck_off[c] + b2 + 1]; }
        while (ta < ea || tb < eb) {
            int32_t i;
            if (tb >= eb || (ta < ea && ix->ids[ta] <= ix->ids[tb])) {
                i = ix->ids[ta++];
                if (tb < eb && ix->ids[tb] == i) ++tb;
            } else i = ix->ids[tb++];
            const int a = loc1 >= ix->L[i] && loc1 + 1 <= ix->R[i];            /* :122 */
            const int bq = loc2 - 1 >= ix->L[i] && loc2 <= ix->R[i];           /* :125 */
            if (a || bq) {
                types |= 1u << ix->type[i];
                if (nh < 4096) hits[nh++] = i; else overflow = 1;
            }
        }
        if (!nh) continue;                                                      /* :128 */
        j->stats[0]++;                                                          /* :149 */
        if (overflow) abort();                                                  /* never with sane indices; keeps the checker honest */
        if (!(types & ((1u << T_GENE) | (1u << T_TE) | (1u << T_SNRNA)))) {
            if (types & (1u << T_ENH)) j->stats[4]++;                          /* :145-147 NameError */
            continue;
        }
        for (int h = 0; h < nh; ++h) {                                          /* one per distinct ensg */
            const int32_t e = ix->ensg[hits[h]];
            if (j->estamp[e] == tick) continue;
            j->estamp[e] = tick;
            j->counts[e]++;
        }
    }
    return 0;
}

/* counts[n_ensg]; stats[7] = units, assigned, lowq, badchrom, qcfail, crash_enhancer, crash_name (tecount.h order) */
int teo_bulk_count(void* index, int paired, int qual, int64_t n_rec, const int32_t* start, const int32_t* end,
                   const uint16_t* chrom, const uint8_t* mapq, const uint8_t* flag, int64_t* counts, int64_t* stats, int n_threads) {
    const teo_index* ix = (const teo_index*)index;
    const int64_t n_units = paired ? n_rec / 2 : n_rec;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    teo_job* jobs = (teo_job*)calloc((size_t)n_threads, sizeof(teo_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        teo_job* j = &jobs[t];
        j->ix = ix; j->paired = paired; j->qual = qual;
        j->u0 = n_units * t / n_threads; j->u1 = n_units * (t + 1) / n_threads;
        j->start = start; j->end = end; j->chrom = chrom; j->mapq = mapq; j->flag = flag;
        j->counts = (int64_t*)calloc((size_t)ix->n_ensg + 1, 8);
        j->estamp = (uint32_t*)calloc((size_t)ix->n_ensg + 1, 4);
        pthread_create(&th[t], 0, teo_worker, j);
    }
    memset(counts, 0, (size_t)ix->n_ensg * 8);
    memset(stats, 0, 7 * 8);
    stats[0] = n_units;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], 0);
        for (int e = 0; e < ix->n_ensg; ++e) counts[e] += jobs[t].counts[e];
        for (int k = 0; k < 6; ++k) stats[1 + k] += jobs[t].stats[k];
        free(jobs[t].counts); free(jobs[t].estamp);
    }
    free(jobs); free(th);
    return 0;
}
