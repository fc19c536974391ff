// inq_cohort.cuh -- sm_100a kernels of the cohort `outlier` rows (reference: src/outlier.rs, v0.13.0).
//
// z-score: the reference sums a row's f32 values left to right (outlier.rs:19,22-28), so bit-exact
// parity needs the same order: one thread owns one row and adds sequentially with __fadd_rn / __fmul_rn
// (never contracted to FMA); the CTA stages 32-column tiles through shared memory so that global loads
// stay coalesced although consecutive threads own consecutive rows.
// dbscan: one CTA per row; sort, then the order-independent form of dbscan 0.3.1 in one dimension:
// core(i) = #{j : |x_i - x_j| < eps} >= min_points, noise(i) = !core(i) && no core point within eps.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace inqc {

struct CohortCounters {
    unsigned long long n_hits;
    unsigned int no_mode;              // a kept row had no positive value (dbscan)
    unsigned int no_mode_row_lo, no_mode_row_hi;
};

constexpr int kZRows = 128;            // rows (= threads) per CTA
constexpr int kZTile = 32;             // columns per shared-memory tile

__device__ __forceinline__ float clean(float v) { return (v != v) ? 0.0f : v; }   // outlier.rs:81-84

// loads the tile [row0, row0 + kZRows) x [c0, c0 + kZTile) (NaN -> 0) with coalesced 128-byte row segments
__device__ __forceinline__ void load_tile(const float *__restrict__ m, uint64_t n_rows, uint32_t n_cols, uint64_t row0,
                                          uint32_t c0, float (*tile)[kZTile + 1])
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (uint32_t r = warp; r < (uint32_t)kZRows; r += nwarp) {
        const uint64_t row = row0 + r;
        const uint32_t c = c0 + lane;
        tile[r][lane] = (row < n_rows && c < n_cols) ? clean(m[row * n_cols + c]) : 0.0f;
    }
}

__global__ void __launch_bounds__(kZRows)
k_outlier_zscore(const float *__restrict__ m, uint64_t n_rows, uint32_t n_cols, float minsize, float cutoff,
                 uint8_t *__restrict__ row_kept, unsigned long long *__restrict__ hits, uint64_t cap,
                 CohortCounters *__restrict__ ctr)
{
    __shared__ float tile[kZRows][kZTile + 1];
    __shared__ float s_mean[kZRows], s_sd[kZRows];
    __shared__ uint8_t s_kept[kZRows];
    const uint64_t row0 = (uint64_t)blockIdx.x * kZRows, row = row0 + threadIdx.x;
    // pass 1: sequential f32 sum and the maximum (outlier.rs:19, 87-90)
    float sum = 0.0f, mx = 0.0f;
    for (uint32_t c0 = 0; c0 < n_cols; c0 += kZTile) {
        __syncthreads();
        load_tile(m, n_rows, n_cols, row0, c0, tile);
        __syncthreads();
        const uint32_t n = min((uint32_t)kZTile, n_cols - c0);
        for (uint32_t c = 0; c < n; ++c) {
            const float v = tile[threadIdx.x][c];
            sum = __fadd_rn(sum, v);
            mx = (c0 + c == 0 || !(v < mx)) ? v : mx;
        }
    }
    const float count = (float)n_cols;
    const float mean = __fdiv_rn(sum, count);
    const bool kept = row < n_rows && !(mx < minsize);
    // pass 2: population variance around the f32 mean (outlier.rs:22-29)
    float var = 0.0f;
    for (uint32_t c0 = 0; c0 < n_cols; c0 += kZTile) {
        __syncthreads();
        load_tile(m, n_rows, n_cols, row0, c0, tile);
        __syncthreads();
        const uint32_t n = min((uint32_t)kZTile, n_cols - c0);
        for (uint32_t c = 0; c < n; ++c) {
            const float diff = __fsub_rn(mean, tile[threadIdx.x][c]);
            var = __fadd_rn(var, __fmul_rn(diff, diff));
        }
    }
    const float sd = __fsqrt_rn(__fdiv_rn(var, count));
    s_mean[threadIdx.x] = mean;
    s_sd[threadIdx.x] = sd;
    s_kept[threadIdx.x] = kept ? 1 : 0;
    if (row < n_rows && row_kept) row_kept[row] = kept ? 1 : 0;
    // pass 3: flags, element-parallel (no order dependence): (v - mean) / sd >= cutoff (outlier.rs:109)
    for (uint32_t c0 = 0; c0 < n_cols; c0 += kZTile) {
        __syncthreads();
        load_tile(m, n_rows, n_cols, row0, c0, tile);
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        for (uint32_t r = warp; r < (uint32_t)kZRows; r += nwarp) {
            const uint32_t c = c0 + lane;
            if (!s_kept[r] || c >= n_cols) continue;
            const float z = __fdiv_rn(__fsub_rn(tile[r][lane], s_mean[r]), s_sd[r]);
            if (z >= cutoff) {
                const unsigned long long slot = atomicAdd(&ctr->n_hits, 1ull);
                if (slot < cap) hits[slot] = ((row0 + r) << 32) | c;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- dbscan
constexpr int kDbThreads = 128;

__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if ((int)lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
        if (w < warp) base += s_warp[w];
        tot += s_warp[w];
    }
    __syncthreads();
    *total = tot;
    return base + x - v;
}

// neighbours of sorted element i: the contiguous run [lo, hi) with |key[i] - key[j]| < eps, compared in f64
// like dbscan's euclidean distance (range_query's `distance < eps`)
__device__ __forceinline__ void eps_range(const float *key, uint32_t n, uint32_t i, double eps, uint32_t *lo, uint32_t *hi)
{
    const double x = (double)key[i];
    uint32_t a = 0, b = i;                               // first j <= i with x - key[j] < eps
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (!(x - (double)key[mid] < eps)) a = mid + 1; else b = mid; }
    *lo = a;
    a = i; b = n;                                        // first j >= i with key[j] - x >= eps
    while (a < b) { const uint32_t mid = (a + b) >> 1; if ((double)key[mid] - x < eps) a = mid + 1; else b = mid; }
    *hi = a;
}

// one CTA per row. Shared memory: key[n2] f32, idx[n2] u16, pc[n2 + 1] u32 (prefix count of core points)
__global__ void __launch_bounds__(kDbThreads)
k_outlier_dbscan(const float *__restrict__ m, uint64_t n_rows, uint32_t n_cols, uint32_t n2, float minsize,
                 uint32_t min_points, uint8_t *__restrict__ row_kept, unsigned long long *__restrict__ hits,
                 uint64_t cap, CohortCounters *__restrict__ ctr)
{
    extern __shared__ unsigned char db_smem[];
    float *key = reinterpret_cast<float *>(db_smem);
    uint32_t *pc = reinterpret_cast<uint32_t *>(key + n2);
    uint16_t *idx = reinterpret_cast<uint16_t *>(pc + n2 + 1);
    __shared__ uint32_t s_warp[kDbThreads / 32];
    __shared__ float s_wmax[kDbThreads / 32];
    __shared__ unsigned long long s_best;
    const uint32_t tid = threadIdx.x;
    for (uint64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        __syncthreads();
        if (tid == 0) s_best = 0ull;
        float mx = -INFINITY;
        for (uint32_t i = tid; i < n2; i += blockDim.x) {
            float v = INFINITY;                              // padding sorts last
            if (i < n_cols) { v = clean(m[row * n_cols + i]); mx = (v > mx) ? v : mx; }
            key[i] = v;
            idx[i] = (uint16_t)i;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) { const float o = __shfl_xor_sync(0xffffffffu, mx, d); mx = (o > mx) ? o : mx; }
        if ((tid & 31) == 0) s_wmax[tid >> 5] = mx;
        __syncthreads();
        float row_max = s_wmax[0];
        for (uint32_t w = 1; w < (blockDim.x >> 5); ++w) row_max = (s_wmax[w] > row_max) ? s_wmax[w] : row_max;
        const bool kept = !(row_max < minsize);
        if (tid == 0 && row_kept) row_kept[row] = kept ? 1 : 0;
        if (!kept) continue;

        // ---- bitonic sort by value, carrying the column index
        for (uint32_t k = 2; k <= n2; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (uint32_t i = tid; i < n2; i += blockDim.x) {
                    const uint32_t p = i ^ j;
                    if (p > i) {
                        const bool up = (i & k) == 0;
                        const float a = key[i], b = key[p];
                        if ((a > b) == up) {
                            key[i] = b; key[p] = a;
                            const uint16_t t = idx[i]; idx[i] = idx[p]; idx[p] = t;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- mode of `value as usize` over the positive values (outlier.rs:133-145): the sorted order
        //      is also the order of the truncated values, so a value's count is a pair of binary searches.
        //      Ties: the smallest value (the reference's HashMap order is random).
        unsigned long long best = 0ull;                      // (count << 32) | ~rank  -> max = most frequent, then smallest
        for (uint32_t i = tid; i < n_cols; i += blockDim.x) {
            const float v = key[i];
            if (!(v > 0.0f)) continue;
            const float t = truncf(v);
            uint32_t lo = 0, hi = n_cols;                    // first j with trunc(key[j]) >= t  (key[j] >= t)
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (key[mid] < t) lo = mid + 1; else hi = mid; }
            uint32_t lo2 = lo, hi2 = n_cols;                 // first j with key[j] >= t + 1
            const float t1 = t + 1.0f;
            while (lo2 < hi2) { const uint32_t mid = (lo2 + hi2) >> 1; if (key[mid] < t1) lo2 = mid + 1; else hi2 = mid; }
            // positives only: for t == 0 the run starts at the first positive value
            uint32_t first = lo;
            if (t == 0.0f) { uint32_t a = 0, b = n_cols; while (a < b) { const uint32_t mid = (a + b) >> 1; if (!(key[mid] > 0.0f)) a = mid + 1; else b = mid; } first = a; }
            const uint32_t cnt = lo2 - first;
            const unsigned long long cand = ((unsigned long long)cnt << 32) | (0xFFFFFFFFu - first);
            best = cand > best ? cand : best;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
            best = o > best ? o : best;
        }
        if ((tid & 31) == 0 && best) atomicMax(&s_best, best);
        __syncthreads();
        if (s_best == 0ull) {                                // no positive value: the reference panics
            if (tid == 0) {
                atomicExch(&ctr->no_mode, 1u);
                ctr->no_mode_row_lo = (unsigned int)row;
                ctr->no_mode_row_hi = (unsigned int)(row >> 32);
            }
            continue;
        }
        const uint32_t mode_first = 0xFFFFFFFFu - (uint32_t)(s_best & 0xFFFFFFFFull);
        const float mode_f = truncf(key[mode_first]);
        // eps = max(2 * mode, 10) as f64 (outlier.rs:118); mode < 2^24 is exact in f32 and doubling is exact in f64
        const double eps = fmax(2.0 * (double)mode_f, 10.0);

        // ---- core points: |x_i - x_j| < eps in f64 (dbscan's euclidean distance), j over the whole row
        __syncthreads();
        for (uint32_t i = tid; i < n2; i += blockDim.x) {
            uint32_t lo = 0, hi = 0;
            if (i < n_cols) eps_range(key, n_cols, i, eps, &lo, &hi);
            pc[i] = (i < n_cols && hi - lo >= min_points) ? 1u : 0u;
        }
        __syncthreads();
        // exclusive prefix count of core points over the sorted order: chunked block scan
        uint32_t carry = 0;
        for (uint32_t base = 0; base < n2; base += blockDim.x) {
            const uint32_t i = base + tid;
            const uint32_t v = i < n2 ? pc[i] : 0u;
            uint32_t tot;
            const uint32_t ex = block_scan_excl(v, s_warp, &tot);
            if (i < n2) pc[i] = carry + ex;
            carry += tot;
        }
        if (tid == 0) pc[n2] = carry;
        __syncthreads();
        // ---- noise = not core and no core point within eps (Edge otherwise)
        for (uint32_t i = tid; i < n_cols; i += blockDim.x) {
            uint32_t lo, hi;
            eps_range(key, n_cols, i, eps, &lo, &hi);
            const bool core = hi - lo >= min_points;
            const bool near_core = pc[hi] - pc[lo] > 0u;
            if (!core && !near_core) {
                const unsigned long long slot = atomicAdd(&ctr->n_hits, 1ull);
                if (slot < cap) hits[slot] = (row << 32) | idx[i];
            }
        }
    }
}

}  // namespace inqc
