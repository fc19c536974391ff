/*
 * oracle.c -- CPU restatement of inquiSTR's `call` hot path (see oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no runnable reference here).
 * All `call.rs:N` / `repeats.rs:N` citations are relative to
 * /root/reference/src/ (wdecoster/inquiSTR v0.13.0).
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* BAM CIGAR op codes, SAM spec 4.2: MIDNSHP=X -> 0..8 */
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };

/* ---- call.rs:377-413 call_from_cigar -------------------------------------- */
int64_t orc_call_from_cigar(int32_t ref_start, const uint32_t *cigar, uint64_t n_cigar,
                            uint32_t minlen, uint32_t start_ext, uint32_t end_ext,
                            int accidental_2d, int *clip)
{
    int64_t call = 0;
    /* call.rs:380 -- 1-based cursor held in a u32 (release build: wrapping add) */
    uint32_t refpos = (uint32_t)((int64_t)ref_start + 1);
    int clipped = 0;
    for (uint64_t i = 0; i < n_cigar; ++i) {
        uint32_t len = cigar[i] >> 4;
        switch (cigar[i] & 0xF) {
        case OP_M: case OP_EQ: case OP_X:            /* call.rs:384-386 */
            refpos += len;
            break;
        case OP_D:                                   /* call.rs:387-392: test before advancing */
            if (len > minlen && start_ext < refpos && refpos < end_ext) call -= (int64_t)len;
            refpos += len;
            break;
        case OP_S:                                   /* call.rs:393-398 */
            if (!accidental_2d && len > minlen && start_ext < refpos && refpos < end_ext) {
                call += (int64_t)len;
                clipped = 1;
            }
            break;
        case OP_I:                                   /* call.rs:399-403 */
            if (len > minlen && start_ext < refpos && refpos < end_ext) call += (int64_t)len;
            break;
        case OP_N:                                   /* call.rs:404 */
            refpos += len;
            break;
        default:                                     /* call.rs:405: H, P ignored */
            break;
        }
    }
    if (clip) *clip = clipped;                       /* call.rs:408-412 */
    return call;
}

/* ---- call.rs:497-522 median_str_length ------------------------------------ */
static int cmp_i64_asc(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}
static int cmp_i64_desc(const void *a, const void *b) { return cmp_i64_asc(b, a); }

double orc_median_str_length(const int64_t *values, const uint8_t *clip, size_t n,
                             size_t support, int *panicked)
{
    if (panicked) *panicked = 0;
    if (n < support) return NAN;                     /* call.rs:498-500 */
    int64_t *spanning = (int64_t *)malloc(sizeof(int64_t) * (n + support + 1));
    int64_t *clipped = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    size_t ns = 0, nc = 0;
    for (size_t i = 0; i < n; ++i) {                 /* call.rs:503-508 */
        if (clip[i]) clipped[nc++] = values[i];
        else spanning[ns++] = values[i];
    }
    if (ns <= support) {                             /* call.rs:509-513 */
        qsort(clipped, nc, sizeof(int64_t), cmp_i64_desc);
        size_t take = support - ns;                  /* <= nc because n >= support */
        for (size_t i = 0; i < take; ++i) spanning[ns++] = clipped[i];
    }
    qsort(spanning, ns, sizeof(int64_t), cmp_i64_asc); /* call.rs:514 */
    double out;
    if (ns == 0) {
        /* call.rs:516: `len/2 - 1` underflows on an empty vector (only reachable
         * with support == 0) -> panic in the reference */
        if (panicked) *panicked = 1;
        out = NAN;
    } else if ((ns % 2) == 0) {                      /* call.rs:515-518 */
        out = (double)(spanning[ns / 2 - 1] + spanning[ns / 2]) / 2.0;
    } else {                                         /* call.rs:519-521 */
        out = (double)spanning[ns / 2];
    }
    free(spanning);
    free(clipped);
    return out;
}

/* ---- call.rs:461-477 cigar_to_rlen ---------------------------------------- */
int64_t orc_cigar_to_rlen(const char *s)
{
    int64_t rlen = 0, num = 0;
    for (; *s; ++s) {
        if (*s >= '0' && *s <= '9') {
            num = num * 10 + (*s - '0');
        } else {
            switch (*s) {
            case 'M': case '=': case 'X': case 'D': case 'N': rlen += num; break;
            default: break;
            }
            num = 0;
        }
    }
    return rlen;
}

/* ---- call.rs:415-459 is_accidental_2d ------------------------------------- */
int orc_is_accidental_2d(int is_reverse, const char *sa, int64_t ref_start, int64_t ref_end)
{
    if (!sa) return 0;                               /* call.rs:425-427 */
    char read_strand = is_reverse ? '-' : '+';       /* call.rs:422 */
    /* call.rs:434: split on ';', drop empty entries */
    const char *entry = NULL;
    size_t entry_len = 0;
    int n_entries = 0;
    const char *p = sa;
    while (1) {
        const char *q = strchr(p, ';');
        size_t len = q ? (size_t)(q - p) : strlen(p);
        if (len > 0) {
            if (n_entries == 0) { entry = p; entry_len = len; }
            ++n_entries;
        }
        if (!q) break;
        p = q + 1;
    }
    if (n_entries > 1) return 0;                     /* call.rs:436-438 */
    if (n_entries == 0) return 0;                    /* reference would index-panic; not produced by aligners */
    /* call.rs:439: rname,POS,strand,CIGAR,mapQ,NM */
    char buf[4096];
    if (entry_len >= sizeof(buf)) entry_len = sizeof(buf) - 1;
    memcpy(buf, entry, entry_len);
    buf[entry_len] = 0;
    char *field[6] = {0};
    int nf = 0;
    char *tok = buf;
    while (nf < 6) {
        field[nf++] = tok;
        char *c = strchr(tok, ',');
        if (!c) break;
        *c = 0;
        tok = c + 1;
    }
    if (nf < 4) return 0;
    if (read_strand == field[2][0]) return 0;        /* call.rs:441-443 */
    int64_t sa_start = strtoll(field[1], NULL, 10);  /* call.rs:450: 1-based POS used as is */
    int64_t sa_end = sa_start + orc_cigar_to_rlen(field[3]); /* call.rs:451 */
    int64_t lo = ref_start > sa_start ? ref_start : sa_start;
    int64_t hi = ref_end < sa_end ? ref_end : sa_end;
    return lo < hi;                                  /* call.rs:454 */
}

/* ---- read index standing in for htslib's fetch (call.rs:288,338) ----------- */
struct orc_index {
    int32_t n_contigs;
    uint64_t *contig_off;  /* n_contigs+1 into order[] */
    uint64_t *order;       /* read ids, per contig sorted by (ref_start, id) */
    int32_t *pmax_end;     /* running max of ref_end along order[], per contig */
};

typedef struct { int32_t start; uint64_t id; } start_id;
static int cmp_start_id(const void *a, const void *b)
{
    const start_id *x = (const start_id *)a, *y = (const start_id *)b;
    if (x->start != y->start) return (x->start > y->start) - (x->start < y->start);
    return (x->id > y->id) - (x->id < y->id);
}

orc_index *orc_index_build(const orc_reads *rd, int32_t n_contigs)
{
    orc_index *ix = (orc_index *)calloc(1, sizeof(*ix));
    ix->n_contigs = n_contigs;
    ix->contig_off = (uint64_t *)calloc((size_t)n_contigs + 1, sizeof(uint64_t));
    uint64_t n = rd->n_reads, kept = 0;
    for (uint64_t r = 0; r < n; ++r) {
        int32_t c = rd->contig[r];
        if (c >= 0 && c < n_contigs) { ix->contig_off[c + 1]++; kept++; }
    }
    for (int32_t c = 0; c < n_contigs; ++c) ix->contig_off[c + 1] += ix->contig_off[c];
    ix->order = (uint64_t *)malloc(sizeof(uint64_t) * (kept + 1));
    ix->pmax_end = (int32_t *)malloc(sizeof(int32_t) * (kept + 1));
    uint64_t *cursor = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_contigs + 1));
    memcpy(cursor, ix->contig_off, sizeof(uint64_t) * ((size_t)n_contigs + 1));
    for (uint64_t r = 0; r < n; ++r) {
        int32_t c = rd->contig[r];
        if (c >= 0 && c < n_contigs) ix->order[cursor[c]++] = r;
    }
    free(cursor);
    for (int32_t c = 0; c < n_contigs; ++c) {
        uint64_t a = ix->contig_off[c], b = ix->contig_off[c + 1];
        int sorted = 1;
        for (uint64_t i = a + 1; i < b && sorted; ++i)
            if (rd->ref_start[ix->order[i - 1]] > rd->ref_start[ix->order[i]]) sorted = 0;
        if (!sorted) {
            start_id *tmp = (start_id *)malloc(sizeof(start_id) * (b - a));
            for (uint64_t i = a; i < b; ++i) {
                tmp[i - a].start = rd->ref_start[ix->order[i]];
                tmp[i - a].id = ix->order[i];
            }
            qsort(tmp, b - a, sizeof(start_id), cmp_start_id);
            for (uint64_t i = a; i < b; ++i) ix->order[i] = tmp[i - a].id;
            free(tmp);
        }
        int32_t m = INT32_MIN;
        for (uint64_t i = a; i < b; ++i) {
            int32_t e = rd->ref_end[ix->order[i]];
            if (e > m) m = e;
            ix->pmax_end[i] = m;
        }
    }
    return ix;
}

void orc_index_free(orc_index *ix)
{
    if (!ix) return;
    free(ix->contig_off);
    free(ix->order);
    free(ix->pmax_end);
    free(ix);
}

typedef struct { int64_t v; uint8_t clip; } callrec;
static int cmp_callrec(const void *a, const void *b)
{
    /* call.rs:308-312 sorts on the value only (unstable); ties between a Span
     * and a Clip of equal value are resolved here as Span first. */
    const callrec *x = (const callrec *)a, *y = (const callrec *)b;
    if (x->v != y->v) return (x->v > y->v) - (x->v < y->v);
    return (int)x->clip - (int)y->clip;
}

typedef struct { callrec *p; size_t n, cap; } callvec;
static void cv_push(callvec *v, int64_t val, int clip)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->p = (callrec *)realloc(v->p, v->cap * sizeof(callrec));
    }
    v->p[v->n].v = val;
    v->p[v->n].clip = (uint8_t)clip;
    v->n++;
}

static double median_of(const callrec *c, size_t n, size_t support, int *panicked)
{
    int64_t *v = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    uint8_t *k = (uint8_t *)malloc(n + 1);
    for (size_t i = 0; i < n; ++i) { v[i] = c[i].v; k[i] = c[i].clip; }
    double r = orc_median_str_length(v, k, n, support, panicked);
    free(v);
    free(k);
    return r;
}

/* ---- call.rs:279-327 / 329-374 -------------------------------------------- */
int orc_genotype_locus(const orc_reads *rd, const orc_index *ix, int32_t tid,
                       uint32_t start, uint32_t end, uint32_t minlen, size_t support,
                       int unphased, double *phase1, double *phase2, uint64_t *op_visits)
{
    *phase1 = NAN;
    *phase2 = NAN;
    if (start < 10) return ORC_PANIC_START_LT_10;    /* call.rs:285/335 `start - 10` on u32 */
    if (tid < 0 || tid >= ix->n_contigs) return ORC_PANIC_BAD_INTERVAL;
    uint32_t start_ext = start - 10;                 /* call.rs:285 */
    uint32_t end_ext = end + 10;                     /* call.rs:286 */

    /* htslib region query: pos < end_ext && endpos > start_ext on this tid */
    uint64_t a = ix->contig_off[tid], b = ix->contig_off[tid + 1];
    uint64_t lo = a, hi = b;
    { /* first i in [a,b) with pmax_end > start_ext */
        uint64_t l = a, h = b;
        while (l < h) {
            uint64_t m = l + (h - l) / 2;
            if ((int64_t)ix->pmax_end[m] > (int64_t)start_ext) h = m; else l = m + 1;
        }
        lo = l;
    }
    { /* first i in [a,b) with ref_start >= end_ext */
        uint64_t l = a, h = b;
        while (l < h) {
            uint64_t m = l + (h - l) / 2;
            if ((int64_t)rd->ref_start[ix->order[m]] >= (int64_t)end_ext) h = m; else l = m + 1;
        }
        hi = l;
    }

    callvec bucket[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; /* phased: HP 0,1,2; unphased uses [0] */
    int rc = ORC_OK;
    uint64_t visits = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        uint64_t r = ix->order[i];
        if (!((int64_t)rd->ref_end[r] > (int64_t)start_ext)) continue; /* not yielded by fetch */
        uint32_t rs = (uint32_t)rd->ref_start[r];    /* `as u32` call.rs:297/351 */
        uint32_t re = (uint32_t)rd->ref_end[r];
        uint8_t mq = rd->mapq[r];
        uint8_t hp = rd->hp[r];
        if (unphased) {
            /* call.rs:297-302 */
            if (start_ext < rs || re < end_ext || mq <= 10) continue;
        } else {
            /* call.rs:349-355: a || (b && c) || d */
            if (hp == 0xFF || (start_ext < rs && re < end_ext) || mq <= 10) continue;
        }
        int clip = 0;
        uint64_t c0 = rd->cigar_off[r], c1 = rd->cigar_off[r + 1];
        /* call.rs:303/357 -> 394: the walk evaluates is_accidental_2d on the first S op, which panics
           on a non-string / malformed SA tag (call.rs:431,439-450) -- only for reads that got this far */
        if (rd->flags[r] & 2) { rc = ORC_PANIC_BAD_SA; break; }
        int64_t v = orc_call_from_cigar(rd->ref_start[r], rd->cigar + c0, c1 - c0, minlen,
                                        start_ext, end_ext, rd->flags[r] & 1, &clip);
        visits += c1 - c0;
        if (unphased) {
            cv_push(&bucket[0], v, clip);            /* call.rs:303-304 */
        } else {
            if (hp > 2) { rc = ORC_PANIC_BAD_HP; break; } /* call.rs:358 */
            cv_push(&bucket[hp], v, clip);
        }
    }
    if (op_visits) __atomic_fetch_add(op_visits, visits, __ATOMIC_RELAXED);
    if (rc == ORC_OK) {
        int p1 = 0, p2 = 0;
        if (unphased) {
            /* call.rs:308-321 */
            qsort(bucket[0].p, bucket[0].n, sizeof(callrec), cmp_callrec);
            size_t half = bucket[0].n / 2;
            *phase1 = median_of(bucket[0].p, half, support, &p1);
            *phase2 = median_of(bucket[0].p + half, bucket[0].n - half, support, &p2);
        } else {
            /* call.rs:365-369 */
            *phase1 = median_of(bucket[1].p, bucket[1].n, support, &p1);
            *phase2 = median_of(bucket[2].p, bucket[2].n, support, &p2);
        }
        if (p1 || p2) rc = ORC_PANIC_MEDIAN_EMPTY;
    }
    free(bucket[0].p);
    free(bucket[1].p);
    free(bucket[2].p);
    return rc;
}

/* ---- call.rs:103-158 fan-out over loci ------------------------------------ */
typedef struct {
    const orc_reads *rd;
    const orc_index *ix;
    uint64_t n_loci;
    const int32_t *lc;
    const uint32_t *ls, *le;
    uint32_t minlen;
    size_t support;
    int unphased;
    double *p1, *p2;
    uint64_t *op_visits;
    uint64_t next;          /* shared work counter */
    int rc;
    uint64_t rc_locus;
    pthread_mutex_t mu;
} fanout;

static void *fanout_worker(void *arg)
{
    fanout *f = (fanout *)arg;
    const uint64_t grain = 64;
    for (;;) {
        uint64_t i0 = __atomic_fetch_add(&f->next, grain, __ATOMIC_RELAXED);
        if (i0 >= f->n_loci) break;
        uint64_t i1 = i0 + grain < f->n_loci ? i0 + grain : f->n_loci;
        for (uint64_t i = i0; i < i1; ++i) {
            int rc = orc_genotype_locus(f->rd, f->ix, f->lc[i], f->ls[i], f->le[i], f->minlen,
                                        f->support, f->unphased, &f->p1[i], &f->p2[i], f->op_visits);
            if (rc != ORC_OK) {
                pthread_mutex_lock(&f->mu);
                if (f->rc == ORC_OK || i < f->rc_locus) { f->rc = rc; f->rc_locus = i; }
                pthread_mutex_unlock(&f->mu);
            }
        }
    }
    return NULL;
}

int orc_genotype_loci(const orc_reads *rd, int32_t n_contigs, uint64_t n_loci,
                      const int32_t *locus_contig, const uint32_t *locus_start,
                      const uint32_t *locus_end, uint32_t minlen, size_t support,
                      int unphased, int threads, double *phase1, double *phase2,
                      uint64_t *op_visits)
{
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    orc_index *ix = orc_index_build(rd, n_contigs);
    fanout f;
    memset(&f, 0, sizeof(f));
    f.rd = rd; f.ix = ix; f.n_loci = n_loci;
    f.lc = locus_contig; f.ls = locus_start; f.le = locus_end;
    f.minlen = minlen; f.support = support; f.unphased = unphased;
    f.p1 = phase1; f.p2 = phase2; f.op_visits = op_visits;
    f.rc = ORC_OK;
    pthread_mutex_init(&f.mu, NULL);
    if (threads == 1) {
        fanout_worker(&f);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
        for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, fanout_worker, &f);
        for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
        free(th);
    }
    pthread_mutex_destroy(&f.mu);
    orc_index_free(ix);
    return f.rc;
}

/* ---- human_sort 0.2.2 `compare` (call site call.rs:35) --------------------
 * The crate is a Cargo.lock dependency (human-sort 0.2.2), not vendored.
 * Published algorithm, restated: walk both strings; when both current chars
 * are numeric take the maximal digit run on each side as a u32 and compare
 * the numbers; otherwise compare the two chars and advance both; the first
 * difference decides. If either string runs out first, fall back to plain
 * byte-wise string comparison of the whole strings. */
static uint32_t take_numeric(const unsigned char **p)
{
    uint32_t sum = 0;
    while (**p >= '0' && **p <= '9') {
        sum = sum * 10u + (uint32_t)(**p - '0');     /* release build: wrapping */
        ++*p;
    }
    return sum;
}

int orc_human_compare(const char *a, const char *b)
{
    const unsigned char *x = (const unsigned char *)a, *y = (const unsigned char *)b;
    while (*x && *y) {
        int xd = (*x >= '0' && *x <= '9'), yd = (*y >= '0' && *y <= '9');
        if (xd && yd) {
            uint32_t nx = take_numeric(&x), ny = take_numeric(&y);
            if (nx != ny) return nx < ny ? -1 : 1;
        } else {
            if (*x != *y) return *x < *y ? -1 : 1;
            ++x;
            ++y;
        }
    }
    int c = strcmp(a, b);
    return (c > 0) - (c < 0);
}

/* ---- Rust f64 Display on {NaN} U (1/2)Z (call.rs:57-65) ------------------- */
int orc_format_f64(double v, char *buf, size_t buflen)
{
    if (isnan(v)) return snprintf(buf, buflen, "NaN");
    double twice = v * 2.0;
    long long t = (long long)twice;
    if ((t & 1) == 0) return snprintf(buf, buflen, "%lld", t / 2);
    /* odd number of halves: x.5 ; C truncation toward zero matches the digits */
    long long whole = t / 2;
    if (t < 0 && whole == 0) return snprintf(buf, buflen, "-0.5");
    return snprintf(buf, buflen, "%lld.5", whole);
}

/* ---- repeats.rs:96-115 ----------------------------------------------------- */
int orc_validate_interval(int64_t start, int64_t end, int64_t chrom_len)
{
    if (end < start) return ORC_PANIC_BAD_INTERVAL;              /* repeats.rs:102-104 */
    if (chrom_len < 0 || !(end < chrom_len)) return ORC_PANIC_BAD_INTERVAL; /* repeats.rs:108-114 */
    return ORC_OK;
}
