/*
 * oracle.h -- CPU restatement of inquiSTR's `call` hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under inquistr_b200/ may include, link or
 * execute this. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg use it, as the checker or as the timed
 * CPU baseline, never as the product.
 *
 * PARITY UNPINNED: the reference (wdecoster/inquiSTR v0.13.0) ships no golden
 * output for `call`, its test BAM is absent from the checkout and no Rust
 * toolchain exists in this image, so this restatement could not be checked
 * against a run of the reference. It is kept line-traceable instead: every
 * function cites the src/call.rs / src/repeats.rs lines it follows, and the
 * known-answer vectors in tests/test_oracle_kat.py are derived by hand from
 * those lines.
 *
 * The cohort `outlier` rows (oracle_cohort.c) ARE pinned: the reference's unit
 * tests outlier.rs:147-168 hold one known-answer vector per method, replayed in
 * tests/test_cohort.py.
 */
#ifndef INQ_ORACLE_H
#define INQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes that stand in for the reference's panics */
#define ORC_OK 0
#define ORC_PANIC_BAD_HP (-1)         /* call.rs:358  HP not in {0,1,2} -> unwrap on None */
#define ORC_PANIC_MEDIAN_EMPTY (-2)   /* call.rs:516  support==0 and no values -> index underflow */
#define ORC_PANIC_START_LT_10 (-3)    /* call.rs:285  u32 underflow of start-10 (treated as rejected input) */
#define ORC_PANIC_BAD_INTERVAL (-4)   /* repeats.rs:102-114 */
#define ORC_PANIC_BAD_SA (-5)         /* call.rs:431,439-450  is_accidental_2d on an SA tag it cannot digest */
#define ORC_ERR_ARG (-5)

/* SoA view of aligned reads: exactly what the reference reads off each
 * htslib record on this path (call.rs:297-299,349-352,380-382,422-423,483). */
typedef struct {
    uint64_t n_reads;
    const int32_t *contig;      /* tid */
    const int32_t *ref_start;   /* 0-based pos (reference_start) */
    const int32_t *ref_end;     /* bam_endpos (reference_end) */
    const uint8_t *mapq;
    const uint8_t *hp;          /* HP tag value, 0xFF = tag absent (call.rs:482-491) */
    const uint8_t *flags;       /* bit0: is_accidental_2d(read) (call.rs:415-459); bit1: the read has an S op and
                                   an SA tag on which is_accidental_2d panics (call.rs:431,439-450) */
    const uint64_t *cigar_off;  /* n_reads+1 */
    const uint32_t *cigar;      /* BAM packed words: len<<4 | op, ops MIDNSHP=X */
} orc_reads;

/* one read x one window: call.rs:377-413. returns the signed length sum,
 * *clip = 1 when the result is Call::Clip, 0 when Call::Span. */
int64_t orc_call_from_cigar(int32_t ref_start, const uint32_t *cigar, uint64_t n_cigar,
                            uint32_t minlen, uint32_t start_ext, uint32_t end_ext,
                            int accidental_2d, int *clip);

/* call.rs:497-522. values[i]/clip[i] describe Call::Span/Clip entries.
 * returns NaN when n < support; *panicked set when the reference would panic. */
double orc_median_str_length(const int64_t *values, const uint8_t *clip, size_t n,
                             size_t support, int *panicked);

/* call.rs:461-477 */
int64_t orc_cigar_to_rlen(const char *cigar_text);

/* call.rs:415-459. sa == NULL means no SA tag. ref_start/ref_end 0-based record coords. */
int orc_is_accidental_2d(int is_reverse, const char *sa, int64_t ref_start, int64_t ref_end);

/* call.rs:279-327 (unphased != 0) and call.rs:329-374 (unphased == 0) for one
 * locus on contig `tid`. The read set is the one htslib's fetch(tid,start_ext,end_ext)
 * yields: pos < end_ext && endpos > start_ext, any flag.
 * `order`/`pmax` come from orc_index_build. Returns ORC_OK or a panic code. */
typedef struct orc_index orc_index;
orc_index *orc_index_build(const orc_reads *reads, int32_t n_contigs);
void orc_index_free(orc_index *ix);

int orc_genotype_locus(const orc_reads *reads, const orc_index *ix, int32_t tid,
                       uint32_t start, uint32_t end, uint32_t minlen, size_t support,
                       int unphased, double *phase1, double *phase2,
                       uint64_t *op_visits /* nullable: += n_cigar per walked pair */);

/* call.rs:103-158: all loci, `threads` worker threads pulling loci off a shared
 * counter (the reference's par_bridge fan-out). Output in input order.
 * Returns first panic code seen (lowest locus index), ORC_OK otherwise. */
int orc_genotype_loci(const orc_reads *reads, int32_t n_contigs, uint64_t n_loci,
                      const int32_t *locus_contig, const uint32_t *locus_start,
                      const uint32_t *locus_end, uint32_t minlen, size_t support,
                      int unphased, int threads, double *phase1, double *phase2,
                      uint64_t *op_visits);

/* human_sort 0.2.2 compare (call.rs:35 call site; algorithm restated from the
 * published crate, see oracle.c). returns <0,0,>0 */
int orc_human_compare(const char *a, const char *b);

/* Rust `{}` formatting of an f64 restricted to NaN and multiples of 0.5
 * (call.rs:57-65). Writes a NUL-terminated string, returns its length. */
int orc_format_f64(double v, char *buf, size_t buflen);

/* repeats.rs:96-115 validation. chrom_len < 0 means contig not in header. */
int orc_validate_interval(int64_t start, int64_t end, int64_t chrom_len);

/* ---- cohort `outlier` rows (src/outlier.rs); implemented in oracle_cohort.c ---- */
int orc_repeat_lengths(const float *in, size_t n, uint32_t minsize, float *out);          /* outlier.rs:75-97 */
void orc_std_deviation_and_mean(const float *data, size_t n, float *mean, float *sd);     /* outlier.rs:18-31 */
void orc_zscore_outliers(const float *values, size_t n, float cutoff, uint8_t *flag);     /* outlier.rs:99-113 */
int64_t orc_mode(const float *values, size_t n);                                           /* outlier.rs:133-145 */
int orc_dbscan_outliers(const float *values, size_t n, size_t mincluster, uint8_t *flag); /* outlier.rs:115-131 */

#ifdef __cplusplus
}
#endif
#endif
