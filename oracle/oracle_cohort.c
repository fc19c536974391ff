/* oracle_cohort.c -- CPU restatement of the reference's `outlier` row computations
 * (src/outlier.rs, wdecoster/inquiSTR v0.13.0).
 *
 * TEST INFRASTRUCTURE ONLY: nothing under inquistr_b200/ may call, link or import this file.
 *
 * Pinned by the reference's own unit tests: outlier.rs:147-157 (test_dbscan_outliers) and
 * outlier.rs:159-168 (test_z_score_outliers) -- tests/test_oracle_kat.py replays both vectors.
 * The dbscan crate (0.3.1, Cargo.lock:534-537) is not vendored under /root/reference; its published
 * algorithm (Model::run / expand_cluster / range_query with `distance < eps`, euclidean distance in
 * f64) is restated below from memory of the crate source and cannot be re-verified offline.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* outlier.rs:75-97: NaN -> 0.0; the row is dropped when its maximum is below minsize.
 * out[n] receives the cleaned values. Returns 1 when the row is kept. */
int orc_repeat_lengths(const float *in, size_t n, uint32_t minsize, float *out)
{
    float mx = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        out[i] = isnan(in[i]) ? 0.0f : in[i];
        /* max_by(partial_cmp): the LAST maximal element; only its value is used */
        if (i == 0 || !(out[i] < mx)) mx = out[i];
    }
    if (n == 0) return 0;                          /* the reference unwraps None: no columns is a panic */
    return mx < (float)minsize ? 0 : 1;
}

/* outlier.rs:18-31: f32 throughout, sequential sums, population variance */
void orc_std_deviation_and_mean(const float *data, size_t n, float *mean, float *sd)
{
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += data[i];
    const float count = (float)n;
    const float m = sum / count;
    float var = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        const float diff = m - data[i];
        const float sq = diff * diff;
        var += sq;
    }
    var = var / count;
    *mean = m;
    *sd = sqrtf(var);
}

/* outlier.rs:99-113: flag[i] = ((v - mean) / sd) >= cutoff (only expansions; NaN compares false) */
void orc_zscore_outliers(const float *values, size_t n, float cutoff, uint8_t *flag)
{
    float mean, sd;
    orc_std_deviation_and_mean(values, n, &mean, &sd);
    for (size_t i = 0; i < n; ++i) {
        const float z = (values[i] - mean) / sd;
        flag[i] = (z >= cutoff) ? 1 : 0;
    }
}

/* outlier.rs:133-145: most frequent `value as usize` among values > 0. The reference takes
 * max_by_key over a HashMap, so ties are broken by hash iteration order (random per process);
 * the restatement takes the smallest value among the most frequent. Returns -1 when no value
 * is positive (the reference panics: "No mode found for repeat"). */
static int cmp_u64(const void *a, const void *b)
{
    const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
int64_t orc_mode(const float *values, size_t n)
{
    uint64_t *t = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i)
        if (values[i] > 0.0f) {
            /* Rust `as usize` saturates */
            t[m++] = values[i] >= 18446744073709551615.0f ? UINT64_MAX : (uint64_t)values[i];
        }
    if (m == 0) { free(t); return -1; }
    qsort(t, m, sizeof(uint64_t), cmp_u64);
    uint64_t best = t[0];
    size_t best_n = 0, run = 0;
    for (size_t i = 0; i < m; ++i) {
        run = (i && t[i] == t[i - 1]) ? run + 1 : 1;
        if (run > best_n) { best_n = run; best = t[i]; }
    }
    free(t);
    return (int64_t)best;
}

/* dbscan 0.3.1 Model::run on 1-D points, as called from outlier.rs:115-131:
 * eps = max(2 * mode, 10) as f64, min_points = mincluster; flag[i] = 1 for Noise.
 * Returns 0, or -1 when there is no mode. */
int orc_dbscan_outliers(const float *values, size_t n, size_t mincluster, uint8_t *flag)
{
    const int64_t mode = orc_mode(values, n);
    if (mode < 0) return -1;
    const uint64_t twice = 2u * (uint64_t)mode;
    const double eps = (double)(twice > 10 ? twice : 10);
    enum { NOISE = 0, EDGE = 1, CORE = 2 };
    uint8_t *cls = (uint8_t *)calloc(n ? n : 1, 1), *visited = (uint8_t *)calloc(n ? n : 1, 1);
    size_t *queue = (size_t *)malloc((n ? n : 1) * sizeof(size_t)), *nb = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    for (size_t idx = 0; idx < n; ++idx) {
        if (visited[idx]) continue;
        visited[idx] = 1;
        size_t qn = 0;
        queue[qn++] = idx;
        while (qn) {                                   /* expand_cluster */
            const size_t ind = queue[--qn];
            size_t k = 0;
            for (size_t j = 0; j < n; ++j) {           /* range_query: euclidean distance in f64, strict < */
                const double d = (double)values[ind] - (double)values[j];
                if (sqrt(d * d) < eps) nb[k++] = j;
            }
            if (k < mincluster) continue;
            cls[ind] = CORE;
            for (size_t t = 0; t < k; ++t) {
                const size_t j = nb[t];
                if (cls[j] == NOISE) cls[j] = EDGE;
                if (visited[j]) continue;
                visited[j] = 1;
                queue[qn++] = j;
            }
        }
    }
    for (size_t i = 0; i < n; ++i) flag[i] = cls[i] == NOISE;
    free(cls); free(visited); free(queue); free(nb);
    return 0;
}
