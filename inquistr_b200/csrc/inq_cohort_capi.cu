// inq_cohort_capi.cu -- extern "C" entry points of include/inqcohort.h (part of libinqcall.so)
#include "../../include/inqcall.h"
#include "../../include/inqcohort.h"
#include "inq_cohort.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_cohort_error;

int cfail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_cohort_error = buf;
    return code;
}

struct Guard {                         // frees everything on every exit path
    void *d_m = nullptr, *d_kept = nullptr, *d_hits = nullptr, *d_ctr = nullptr;
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Guard()
    {
        if (d_m) cudaFree(d_m);
        if (d_kept) cudaFree(d_kept);
        if (d_hits) cudaFree(d_hits);
        if (d_ctr) cudaFree(d_ctr);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (s) cudaStreamDestroy(s);
    }
};

#define CC_TRY(call)                                                                            \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) return cfail(INQ_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace

extern "C" {

const char *inq_cohort_last_error(void) { return g_cohort_error.c_str(); }

int inq_outlier(int device, int method, uint64_t n_rows, uint32_t n_cols, const float *values, uint32_t minsize,
                float zscore_cutoff, uint8_t *row_kept, uint64_t *n_hits, uint64_t *hits, uint64_t cap, float *ms_kernel)
{
    using namespace inqc;
    if (n_hits) *n_hits = 0;
    if (ms_kernel) *ms_kernel = 0.f;
    if (method != INQ_OUTLIER_ZSCORE && method != INQ_OUTLIER_DBSCAN) return cfail(INQ_ERR_ARG, "inq_outlier: unknown method %d", method);
    if (n_rows == 0) return INQ_OK;
    if (n_cols == 0) return cfail(INQ_ERR_ARG, "inq_outlier: a row without value columns (the reference panics, outlier.rs:87-91)");
    if (!values || !n_hits || (cap && !hits)) return cfail(INQ_ERR_ARG, "inq_outlier: NULL array");
    if (n_rows > 0xFFFFFFFFull) return cfail(INQ_ERR_TOO_LARGE, "inq_outlier: more than 2^32-1 rows");
    if (method == INQ_OUTLIER_DBSCAN && n_cols > 4096) return cfail(INQ_ERR_TOO_LARGE, "inq_outlier: dbscan supports at most 4096 columns");
    CC_TRY(cudaSetDevice(device));
    Guard g;
    CC_TRY(cudaStreamCreateWithFlags(&g.s, cudaStreamNonBlocking));
    CC_TRY(cudaEventCreate(&g.e0));
    CC_TRY(cudaEventCreate(&g.e1));
    const size_t bytes = (size_t)n_rows * n_cols * sizeof(float);
    CC_TRY(cudaMalloc(&g.d_m, bytes));
    CC_TRY(cudaMalloc(&g.d_kept, n_rows));
    CC_TRY(cudaMalloc(&g.d_hits, std::max<uint64_t>(cap, 1) * sizeof(unsigned long long)));
    CC_TRY(cudaMalloc(&g.d_ctr, sizeof(CohortCounters)));
    CC_TRY(cudaMemcpyAsync(g.d_m, values, bytes, cudaMemcpyHostToDevice, g.s));
    CC_TRY(cudaMemsetAsync(g.d_ctr, 0, sizeof(CohortCounters), g.s));
    CC_TRY(cudaEventRecord(g.e0, g.s));
    if (method == INQ_OUTLIER_ZSCORE) {
        const uint32_t stride = n_cols | 1u;
        const size_t res_smem = (size_t)kZResRows * stride * sizeof(float);
        int sms = 0, smem_max = 0;
        CC_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        CC_TRY(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const size_t per_warp = zw_warp_bytes(n_cols);
        const int zw_warps = (int)std::min<size_t>(kZwMaxWarps, ((size_t)smem_max - 256) / per_warp);
        uint32_t zw_dbg = 0;
        bool force_rows = false;
#ifdef INQ_TIMING_EXPERIMENTS
        if (const char *e = getenv("INQ_ZW_DEBUG")) zw_dbg = (uint32_t)atoi(e);
        if (const char *e = getenv("INQ_ZSCORE_ROWS")) force_rows = atoi(e) != 0;
#endif
        if (n_cols % 4 == 0 && n_cols <= 32u * kZwMaskWords && zw_warps >= 2 && !force_rows) {
            // every warp its own pipeline over groups of 32 rows (bulk async copies, one lane per row)
            const size_t smem = (size_t)zw_warps * per_warp + (size_t)zw_warps * 12 + 128;
            CC_TRY(cudaFuncSetAttribute(k_outlier_zscore_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const uint64_t groups = (n_rows + 31) / 32;
            const unsigned grid = (unsigned)std::min<uint64_t>((groups + zw_warps - 1) / zw_warps, (uint64_t)sms);
            k_outlier_zscore_warp<<<grid, zw_warps * 32, smem, g.s>>>((const float *)g.d_m, n_rows, n_cols, (float)minsize, zscore_cutoff,
                                                                     (uint8_t *)g.d_kept, (unsigned long long *)g.d_hits, cap, (CohortCounters *)g.d_ctr, zw_dbg);
        } else if (res_smem <= 200u * 1024u) {
            // 32 rows fit in shared memory: the matrix is read once
            CC_TRY(cudaFuncSetAttribute(k_outlier_zscore_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem));
            const unsigned grid = (unsigned)((n_rows + kZResRows - 1) / kZResRows);
            k_outlier_zscore_rows<<<grid, kZResThreads, res_smem, g.s>>>((const float *)g.d_m, n_rows, n_cols, stride, (float)minsize, zscore_cutoff,
                                                                        (uint8_t *)g.d_kept, (unsigned long long *)g.d_hits, cap, (CohortCounters *)g.d_ctr);
        } else {
        const unsigned grid = (unsigned)((n_rows + kZRows - 1) / kZRows);
        k_outlier_zscore<<<grid, kZRows, 0, g.s>>>((const float *)g.d_m, n_rows, n_cols, (float)minsize, zscore_cutoff,
                                                   (uint8_t *)g.d_kept, (unsigned long long *)g.d_hits, cap, (CohortCounters *)g.d_ctr);
        }
    } else {
        uint32_t n2 = 32;
        while (n2 < n_cols) n2 <<= 1;
        uint32_t min_points = 0;                              // samples.len().ilog2(), outlier.rs:39
        while ((2ull << min_points) <= n_cols) ++min_points;
        int sms = 0;
        CC_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        // 64..1024 padded columns: one warp per row, sorted in registers (n2 / 32 keys per lane). Otherwise the
        // shared-memory network: one warp for tiny rows, the 128-thread CTA for rows wider than 1024 columns.
        const int ch = (n2 >= 64 && n2 <= 1024) ? (int)(n2 / 32) : 0;
        const size_t nk = ch ? (size_t)n_cols + (n_cols >> 5) + 1 : n_cols;
        const size_t smem = ((size_t)n_cols * 2 + 1 + nk) * sizeof(uint32_t) + 16;   // rng[n], pc[n + 1], key[nk]
        const unsigned threads = n_cols <= 1024 ? 32u : (unsigned)kDbThreads;
        auto launch = [&](auto kernel) -> int {
            int per_sm = 8;
            CC_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CC_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)threads, smem));
            const unsigned grid = (unsigned)std::min<uint64_t>(n_rows, (uint64_t)sms * std::max(per_sm, 1));
            kernel<<<grid, threads, smem, g.s>>>((const float *)g.d_m, n_rows, n_cols, n2, (float)minsize, min_points,
                                                 (uint8_t *)g.d_kept, (unsigned long long *)g.d_hits, cap, (CohortCounters *)g.d_ctr);
            return INQ_OK;
        };
        int lrc;
        switch (ch) {
        case 2: lrc = launch(k_outlier_dbscan<2>); break;
        case 4: lrc = launch(k_outlier_dbscan<4>); break;
        case 8: lrc = launch(k_outlier_dbscan<8>); break;
        case 16: lrc = launch(k_outlier_dbscan<16>); break;
        case 32: lrc = launch(k_outlier_dbscan<32>); break;
        default: lrc = launch(k_outlier_dbscan<0>); break;
        }
        if (lrc != INQ_OK) return lrc;
    }
    CC_TRY(cudaGetLastError());
    CC_TRY(cudaEventRecord(g.e1, g.s));
    CohortCounters h;
    CC_TRY(cudaMemcpyAsync(&h, g.d_ctr, sizeof(h), cudaMemcpyDeviceToHost, g.s));
    if (row_kept) CC_TRY(cudaMemcpyAsync(row_kept, g.d_kept, n_rows, cudaMemcpyDeviceToHost, g.s));
    CC_TRY(cudaStreamSynchronize(g.s));
    if (ms_kernel) cudaEventElapsedTime(ms_kernel, g.e0, g.e1);
    if (h.no_mode)
        return cfail(INQ_ERR_NO_MODE, "row %llu passes the minsize test but has no positive value: no mode (the reference panics, outlier.rs:144)",
                     ((unsigned long long)h.no_mode_row_hi << 32) | h.no_mode_row_lo);
    *n_hits = h.n_hits;
    const uint64_t n = std::min<uint64_t>(h.n_hits, cap);
    if (n) {
        CC_TRY(cudaMemcpy(hits, g.d_hits, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        std::sort(hits, hits + n);                            // (row, column) = the reference's print order
    }
    if (h.n_hits > cap) return cfail(INQ_ERR_HITS_CAP, "inq_outlier: %llu outliers, capacity %llu", (unsigned long long)h.n_hits, (unsigned long long)cap);
    return INQ_OK;
}

}  // extern "C"
