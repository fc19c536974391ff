/* inqcohort.h -- C ABI of the cohort `outlier` rows, part of libinqcall.so.
 *
 * Replaces the per-line body of the reference's `outlier` loop (src/outlier.rs:41-71, wdecoster/inquiSTR
 * v0.13.0) for a whole combined matrix at once: rows = loci, columns = the haplotype columns
 * `<sample>_H1, <sample>_H2, ...` of the `combine` TSV (combine.rs:27-59), values = f32 as parsed by
 * outlier.rs:76-80. There is no CPU implementation behind these entry points.
 *
 *   get_repeat_lengths   outlier.rs:75-97   NaN -> 0, row dropped when max < minsize
 *   z_score_outliers     outlier.rs:18-31,99-113  f32 sequential mean / population sd, (v - mean) / sd >= cutoff
 *   dbscan_outliers      outlier.rs:115-145 + dbscan 0.3.1   eps = max(2 * mode, 10), min_points = ilog2(columns),
 *                        outliers = Noise
 * What the caller keeps: the sample names (suffix stripping outlier.rs:112,130), --subset filtering
 * (outlier.rs:59-67) and the TSV text.
 */
#ifndef INQCOHORT_H
#define INQCOHORT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INQ_OUTLIER_ZSCORE 0       /* outlier.rs Method::Zscore */
#define INQ_OUTLIER_DBSCAN 1       /* outlier.rs Method::Dbscan */

#define INQ_ERR_NO_MODE (-15)      /* dbscan: a kept row without a positive value (the reference panics, outlier.rs:144) */
#define INQ_ERR_HITS_CAP (-16)     /* more outliers than `cap`: *n_hits holds the number needed */

/* Flags the outliers of every row of `values` (host memory, row-major n_rows x n_cols, NaN allowed).
 *   row_kept : n_rows bytes, 1 = the row passed the minsize test (nullable)
 *   hits     : up to `cap` entries (row << 32 | column), sorted ascending = the reference's print order
 *   n_hits   : number of outliers found (also set on INQ_ERR_HITS_CAP)
 * Returns 0 or a negative INQ_ERR_* code (inqcall.h); inq_cohort_last_error() describes the failure. */
int inq_outlier(int device, int method, uint64_t n_rows, uint32_t n_cols, const float *values,
                uint32_t minsize, float zscore_cutoff, uint8_t *row_kept,
                uint64_t *n_hits, uint64_t *hits, uint64_t cap, float *ms_kernel /* nullable */);

const char *inq_cohort_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
