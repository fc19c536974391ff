// inq_capi.cu -- extern "C" boundary of libinqcall.so (see include/inqcall.h).
// Host-side orchestration only: device memory, H2D/D2H copies, kernel launches and timing.
// No CPU implementation of the hot path lives here: without a CUDA device every entry fails.
#include "../../include/inqcall.h"
#include "inq_device.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace inq;

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    uint64_t cap = 0;   // elements
};

enum { EV_START = 0, EV_INDEX, EV_JOIN0, EV_JOIN, EV_CIGAR0, EV_CIGAR, EV_FIXUP, EV_SCAN, EV_PAIRS, EV_MEDIAN, EV_D2H, EV_H2D0, EV_H2D1, EV_COUNT };

}  // namespace

struct inq_ctx {
    int device = 0;
    int sm_count = 0;
    int scan_ctas_per_sm = 1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_chunk[kMedianChunks + 1] = {};
    uint32_t scan_debug = 0, pair_debug = 0;     // timing experiments (INQ_SCAN_DEBUG / INQ_PAIR_DEBUG, read once): results are wrong when != 0
    cudaStream_t stream_join = nullptr;   // the join runs next to the CIGAR scan (latency-bound vs ALU-bound)
    std::string err;

    // locus catalog
    int32_t n_contigs = 0;
    int64_t L = 0;
    DevBuf<int64_t> contig_off;
    DevBuf<int32_t> lstart, lend, lpmax;

    // reads
    uint64_t R = 0, C = 0;
    DevBuf<int32_t> contig, rs, re;
    DevBuf<uint8_t> mapq, hp, flags;
    DevBuf<uint64_t> cig_off;
    DevBuf<uint32_t> cigar;

    // work buffers
    DevBuf<uint32_t> cand_lo, cand_n, ev_off, delta, lcnt, seg_off, big_list;
    DevBuf<unsigned long long> cursor;
    DevBuf<uint32_t> wt_sbase, tile_first;
    DevBuf<uint2> rd_pre, wt;
    DevBuf<uint64_t> desc_scan, desc_wt, vals;
    DevBuf<uint2> evraw;
    CUtensorMap tmap;                 // 2-D view of the packed CIGAR stream: rows of 32 words, 128B swizzle
    const void *tmap_base = nullptr;
    uint64_t tmap_rows = 0;
    DevBuf<uint2> events;
    DevBuf<int64_t> t1, t2;
    DevBuf<uint8_t> valid;
    DevCounters *d_ctr = nullptr;
    DevCounters *h_ctr = nullptr;     // pinned
    uint32_t *h_total = nullptr;      // pinned

    cudaEvent_t ev[EV_COUNT] = {};
    float ms_h2d = 0.f;
    uint64_t last_n_events = 0;
};

namespace {

int fail(inq_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define CU_TRY(ctx, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? INQ_ERR_NOMEM : INQ_ERR_CUDA,   \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// grow-only device buffer; keeps the first `keep` elements
template <typename T>
int ensure(inq_ctx *ctx, DevBuf<T> &b, uint64_t need, uint64_t keep = 0, double growth = 1.0)
{
    if (need <= b.cap) return INQ_OK;
    uint64_t cap = std::max<uint64_t>(need, (uint64_t)(b.cap * growth));
    T *np = nullptr;
    CU_TRY(ctx, cudaMalloc(&np, std::max<uint64_t>(cap, 1) * sizeof(T)));
    if (keep && b.p) {
        cudaError_t e = cudaMemcpyAsync(np, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(np); return fail(ctx, INQ_ERR_CUDA, "device copy failed: %s", cudaGetErrorString(e)); }
    }
    if (b.p) cudaFree(b.p);
    b.p = np;
    b.cap = cap;
    return INQ_OK;
}

template <typename T>
void release(DevBuf<T> &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

#define TRY(x) do { int rc_ = (x); if (rc_ != INQ_OK) return rc_; } while (0)

uint64_t round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }

typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (re)build the TMA descriptor of the CIGAR stream: uint32 [rows][32], box = one 16 KB tile, 128B swizzle
int make_tensor_map(inq_ctx *ctx, uint64_t n_words_padded)
{
    const uint64_t rows = n_words_padded / 32;
    if (ctx->tmap_base == ctx->cigar.p && ctx->tmap_rows == rows) return INQ_OK;
    static tmap_encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(ctx, INQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        encode = (tmap_encode_fn)fn;
    }
    const cuuint64_t gdim[2] = {32, rows};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)(kWarpTileWords / 32)};    // one 2 KB warp tile
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, ctx->cigar.p, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, INQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    ctx->tmap_base = ctx->cigar.p;
    ctx->tmap_rows = rows;
    return INQ_OK;
}

int reserve_reads(inq_ctx *ctx, uint64_t nR, uint64_t nC, double growth)
{
    const uint64_t R = ctx->R, C = ctx->C;
    TRY(ensure(ctx, ctx->contig, nR, R, growth));
    TRY(ensure(ctx, ctx->rs, nR, R, growth));
    TRY(ensure(ctx, ctx->re, nR, R, growth));
    TRY(ensure(ctx, ctx->mapq, nR, R, growth));
    TRY(ensure(ctx, ctx->hp, nR, R, growth));
    TRY(ensure(ctx, ctx->flags, nR, R, growth));
    TRY(ensure(ctx, ctx->cig_off, nR + 1, R + 1, growth));
    // CIGAR stream is padded with zero words up to a tile boundary (+1 tile of slack)
    TRY(ensure(ctx, ctx->cigar, round_up(nC, kTileWords) + kTileWords, C, growth));
    // per warp tile: the first read that starts in it (entries below C / kWarpTileWords stay valid across pushes)
    TRY(ensure(ctx, ctx->tile_first, ctx->cigar.cap / kWarpTileWords + 2, C / kWarpTileWords, 1.0));
    return INQ_OK;
}

}  // namespace

extern "C" {

const char *inq_version(void) { return "inquistr-b200 0.1.0 (sm_100a)"; }

const char *inq_last_error(const inq_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int inq_ctx_create(int device, inq_ctx **out)
{
    if (!out) return fail(nullptr, INQ_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, INQ_ERR_CUDA, "no CUDA device available (%s); libinqcall has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, INQ_ERR_ARG, "device %d out of range [0,%d)", device, n);
    inq_ctx *ctx = new inq_ctx();
    ctx->device = device;
    auto bail = [&](const char *what, cudaError_t err) {
        int rc = fail(nullptr, INQ_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
        inq_ctx_destroy(ctx);
        return rc;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major < 10) {
        fail(nullptr, INQ_ERR_CUDA, "device %d is sm_%d%d; libinqcall is built for sm_100a only", device, prop.major, prop.minor);
        inq_ctx_destroy(ctx);
        return INQ_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream_join, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (int i = 0; i < EV_COUNT; ++i)
        if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i <= kMedianChunks; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_ctr, sizeof(DevCounters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMallocHost(&ctx->h_ctr, sizeof(DevCounters))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost(&ctx->h_total, 64)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaFuncSetAttribute(k_cigar_scan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_cigar_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmemBytes)) != cudaSuccess)
        return bail("cudaFuncSetAttribute(k_cigar_scan)", e);
    if ((e = cudaFuncSetAttribute(k_pair_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairSmem))) != cudaSuccess)
        return bail("cudaFuncSetAttribute(k_pair_eval)", e);
    int occ = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cigar_scan<false>, kCtaThreads, kScanSmemBytes)) != cudaSuccess)
        return bail("occupancy(k_cigar_scan)", e);
    ctx->scan_ctas_per_sm = std::max(1, occ);
    if (const char *dbg = getenv("INQ_SCAN_DEBUG")) ctx->scan_debug = (uint32_t)atoi(dbg);
    if (const char *dbg = getenv("INQ_PAIR_DEBUG")) ctx->pair_debug = (uint32_t)atoi(dbg);
    *out = ctx;
    return INQ_OK;
}

void inq_ctx_destroy(inq_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    release(ctx->contig_off); release(ctx->lstart); release(ctx->lend); release(ctx->lpmax);
    release(ctx->contig); release(ctx->rs); release(ctx->re);
    release(ctx->mapq); release(ctx->hp); release(ctx->flags);
    release(ctx->cig_off); release(ctx->cigar);
    release(ctx->cand_lo); release(ctx->cand_n); release(ctx->ev_off);
    release(ctx->rd_pre); release(ctx->wt); release(ctx->wt_sbase); release(ctx->tile_first);
    release(ctx->desc_wt); release(ctx->evraw);
    release(ctx->delta); release(ctx->lcnt); release(ctx->seg_off); release(ctx->cursor); release(ctx->big_list);
    release(ctx->desc_scan); release(ctx->vals);
    release(ctx->events); release(ctx->t1); release(ctx->t2); release(ctx->valid);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_total) cudaFreeHost(ctx->h_total);
    for (int i = 0; i < EV_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i <= kMedianChunks; ++i)
        if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    if (ctx->stream_join) cudaStreamDestroy(ctx->stream_join);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int inq_host_alloc(size_t bytes, void **out)
{
    if (!out) return INQ_ERR_ARG;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? INQ_ERR_NOMEM : INQ_ERR_CUDA; }
    return INQ_OK;
}

int inq_host_free(void *p)
{
    if (!p) return INQ_OK;
    return cudaFreeHost(p) == cudaSuccess ? INQ_OK : INQ_ERR_CUDA;
}

int inq_set_loci(inq_ctx *ctx, int32_t n_contigs, const int64_t *contig_locus_offsets, const int32_t *start,
                 const int32_t *end)
{
    if (!ctx) return INQ_ERR_ARG;
    if (n_contigs < 0 || !contig_locus_offsets) return fail(ctx, INQ_ERR_ARG, "inq_set_loci: bad contig table");
    const int64_t L = contig_locus_offsets[n_contigs];
    if (contig_locus_offsets[0] != 0 || L < 0) return fail(ctx, INQ_ERR_ARG, "inq_set_loci: offsets must start at 0");
    for (int32_t c = 0; c < n_contigs; ++c)
        if (contig_locus_offsets[c + 1] < contig_locus_offsets[c]) return fail(ctx, INQ_ERR_ARG, "inq_set_loci: offsets not monotonic");
    if (L > 0 && (!start || !end)) return fail(ctx, INQ_ERR_ARG, "inq_set_loci: start/end are NULL");
    if ((uint64_t)L * 2 + 2 > 0x7FFFFFFFull) return fail(ctx, INQ_ERR_TOO_LARGE, "inq_set_loci: too many loci (%lld)", (long long)L);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->n_contigs = 0;
    ctx->L = 0;
    TRY(ensure(ctx, ctx->contig_off, (uint64_t)n_contigs + 1));
    TRY(ensure(ctx, ctx->lstart, (uint64_t)L));
    TRY(ensure(ctx, ctx->lend, (uint64_t)L));
    TRY(ensure(ctx, ctx->lpmax, (uint64_t)L));
    TRY(ensure(ctx, ctx->delta, (uint64_t)L + 2));
    TRY(ensure(ctx, ctx->lcnt, (uint64_t)L + 3));
    TRY(ensure(ctx, ctx->seg_off, (uint64_t)L + 2));
    TRY(ensure(ctx, ctx->cursor, (uint64_t)L + 1));
    TRY(ensure(ctx, ctx->big_list, (uint64_t)L + 1));
    TRY(ensure(ctx, ctx->desc_scan, 2 * (((uint64_t)L + 2 + kXsTile - 1) / kXsTile + 1)));
    TRY(ensure(ctx, ctx->t1, (uint64_t)L));
    TRY(ensure(ctx, ctx->t2, (uint64_t)L));
    TRY(ensure(ctx, ctx->valid, (uint64_t)L));
    cudaStream_t s = ctx->stream;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->contig_off.p, contig_locus_offsets, ((size_t)n_contigs + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    if (L) {
        CU_TRY(ctx, cudaMemcpyAsync(ctx->lstart.p, start, (size_t)L * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemcpyAsync(ctx->lend.p, end, (size_t)L * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DevCounters), s));
        k_locus_check<<<std::min<int64_t>((L + 255) / 256, 4096), 256, 0, s>>>(n_contigs, ctx->contig_off.p, ctx->lstart.p, ctx->lend.p, &ctx->d_ctr->flags);
        if (n_contigs > 0) k_locus_pmax<<<n_contigs, 1024, 0, s>>>(ctx->contig_off.p, ctx->lend.p, ctx->lpmax.p);
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(ctx, cudaStreamSynchronize(s));
    if (L) {
        const unsigned f = ctx->h_ctr->flags;
        if (f & 4u) return fail(ctx, INQ_ERR_LOCUS_ORDER, "a locus has end < start (the reference panics, repeats.rs:102-104)");
        if (f & 2u) return fail(ctx, INQ_ERR_LOCUS_ORDER, "loci must be sorted by start within each contig");
        if (f & 1u) return fail(ctx, INQ_ERR_LOCUS_START, "a locus has start < 10: start-10 underflows u32 in the reference (call.rs:285)");
    }
    ctx->n_contigs = n_contigs;
    ctx->L = L;
    return INQ_OK;
}

int inq_reserve_reads(inq_ctx *ctx, uint64_t n_reads, uint64_t n_cigar_words)
{
    if (!ctx) return INQ_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return reserve_reads(ctx, std::max(n_reads, ctx->R), std::max(n_cigar_words, ctx->C), 1.0);
}

int inq_clear_reads(inq_ctx *ctx)
{
    if (!ctx) return INQ_ERR_ARG;
    ctx->R = 0;
    ctx->C = 0;
    return INQ_OK;
}

int inq_push_reads(inq_ctx *ctx, uint64_t n, const int32_t *contig, const int32_t *ref_start, const int32_t *ref_end,
                   const uint8_t *mapq, const uint8_t *hp, const uint8_t *flags, const uint64_t *cigar_off,
                   const uint32_t *cigar_words)
{
    if (!ctx) return INQ_ERR_ARG;
    if (n == 0) return INQ_OK;
    if (!contig || !ref_start || !ref_end || !mapq || !hp || !flags || !cigar_off)
        return fail(ctx, INQ_ERR_ARG, "inq_push_reads: NULL array");
    if (cigar_off[0] != 0) return fail(ctx, INQ_ERR_ARG, "inq_push_reads: cigar_off[0] must be 0");
    const uint64_t nw = cigar_off[n];
    if (nw && !cigar_words) return fail(ctx, INQ_ERR_ARG, "inq_push_reads: cigar_words is NULL");
    if (ctx->R + n >= 0xFFFFFFFFull) return fail(ctx, INQ_ERR_TOO_LARGE, "inq_push_reads: more than 2^32-2 reads");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(reserve_reads(ctx, ctx->R + n, ctx->C + nw, 1.5));
    cudaStream_t s = ctx->stream;
    const uint64_t R0 = ctx->R, C0 = ctx->C;
    CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_H2D0], s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->contig.p + R0, contig, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->rs.p + R0, ref_start, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->re.p + R0, ref_end, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->mapq.p + R0, mapq, n, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->hp.p + R0, hp, n, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->flags.p + R0, flags, n, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->cig_off.p + R0, cigar_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (nw) CU_TRY(ctx, cudaMemcpyAsync(ctx->cigar.p + C0, cigar_words, nw * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (C0) {
        k_rebase_offsets<<<(unsigned)std::min<uint64_t>((n + 1 + 255) / 256, 8192), 256, 0, s>>>(ctx->cig_off.p + R0, n + 1, C0);
        CU_TRY(ctx, cudaGetLastError());
    }
    const uint64_t C1 = C0 + nw, Cpad = round_up(C1, kTileWords);
    if (Cpad > C1) CU_TRY(ctx, cudaMemsetAsync(ctx->cigar.p + C1, 0, (Cpad - C1) * sizeof(uint32_t), s));
    {
        // reads starting per warp tile, for the tiles that gained words (k_cigar_scan leaves the tile-local
        // prefix of every read start in rd_pre); cig_off[R0 + n] is the sentinel and counts as a start
        const uint64_t t0 = C0 / kWarpTileWords, t1 = Cpad / kWarpTileWords;
        k_tile_first<<<(unsigned)((t1 - t0 + 1 + 255) / 256), 256, 0, s>>>(ctx->cig_off.p, R0 + n + 1, t0, t1, ctx->tile_first.p);
        CU_TRY(ctx, cudaGetLastError());
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_H2D1], s));
    CU_TRY(ctx, cudaStreamSynchronize(s));       // host arrays may be reused by the caller after return
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[EV_H2D0], ctx->ev[EV_H2D1]);
    ctx->ms_h2d = (R0 == 0 ? 0.f : ctx->ms_h2d) + ms;
    ctx->R = R0 + n;
    ctx->C = C1;
    return INQ_OK;
}

int inq_genotype(inq_ctx *ctx, uint32_t minlen, uint32_t support, int unphased, int64_t *twice_h1, int64_t *twice_h2,
                 uint8_t *valid_mask, inq_stats *stats)
{
    if (!ctx) return INQ_ERR_ARG;
    const int64_t L = ctx->L;
    const uint64_t R = ctx->R, C = ctx->C;
    if (L > 0 && (!twice_h1 || !twice_h2 || !valid_mask)) return fail(ctx, INQ_ERR_ARG, "inq_genotype: NULL output array");
    if (minlen >= (1u << 28)) minlen = (1u << 28) - 1;       // BAM op lengths have 28 bits: nothing is longer
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t ntiles = (uint32_t)((C + kTileWords - 1) / kTileWords);
    if ((C + kTileWords - 1) / kTileWords > 0x7FFFFFFFull) return fail(ctx, INQ_ERR_TOO_LARGE, "too many CIGAR words");
    const uint32_t loc_scan_tiles = (uint32_t)(((uint64_t)L + 1 + kXsTile - 1) / kXsTile);
    const uint64_t n_wt = (uint64_t)ntiles * (kTileWords / kWarpTileWords);    // 512-word warp tiles
    const uint32_t wt_scan_tiles = (uint32_t)((n_wt + kXsTile - 1) / kXsTile);
    TRY(ensure(ctx, ctx->cand_lo, R));
    TRY(ensure(ctx, ctx->cand_n, R));
    TRY(ensure(ctx, ctx->rd_pre, R + 1));
    TRY(ensure(ctx, ctx->wt, n_wt + 2));
    TRY(ensure(ctx, ctx->wt_sbase, n_wt + 1));
    TRY(ensure(ctx, ctx->desc_wt, 2 * ((uint64_t)wt_scan_tiles + 1)));
    if (ntiles) TRY(make_tensor_map(ctx, (uint64_t)ntiles * kTileWords));
    const unsigned scan_grid = (unsigned)std::min<uint64_t>((n_wt + kScanWarps - 1) / kScanWarps, (uint64_t)ctx->sm_count * ctx->scan_ctas_per_sm);
    const uint64_t raw_slack = (uint64_t)ctx->sm_count * ctx->scan_ctas_per_sm * kScanWarps * kEvChunk;
    // event storage is sized speculatively (1/16 of the words; checked and regrown after the run); it hands
    // out kEvChunk-slot chunks, so every resident warp may strand one chunk
    if (ctx->evraw.cap == 0) TRY(ensure(ctx, ctx->evraw, C / 16 + 4096 + raw_slack));

    ReadView rv{ctx->contig.p, ctx->rs.p, ctx->re.p, ctx->mapq.p, ctx->hp.p, ctx->flags.p, ctx->cig_off.p, R};
    LocusView lv{ctx->contig_off.p, ctx->lstart.p, ctx->lend.p, ctx->lpmax.p, ctx->n_contigs};
    uint32_t launches = 0;
    const bool work = R && L;

    for (int attempt = 0; attempt < 4; ++attempt) {
        launches = 0;
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_START], s));
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DevCounters), s));
        if (L) {
            CU_TRY(ctx, cudaMemsetAsync(ctx->delta.p, 0, ((uint64_t)L + 2) * sizeof(uint32_t), s));
            CU_TRY(ctx, cudaMemsetAsync(ctx->seg_off.p, 0, ((uint64_t)L + 2) * sizeof(uint32_t), s));
            CU_TRY(ctx, cudaMemsetAsync(ctx->cursor.p, 0, ((uint64_t)L + 1) * sizeof(unsigned long long), s));
            CU_TRY(ctx, cudaMemsetAsync(ctx->desc_scan.p, 0, 2 * ((uint64_t)loc_scan_tiles + 1) * sizeof(uint64_t), s));
        }
        if (!ntiles) CU_TRY(ctx, cudaMemsetAsync(ctx->wt.p, 0, 2 * sizeof(uint2), s));      // no CIGAR words at all
        if (wt_scan_tiles) CU_TRY(ctx, cudaMemsetAsync(ctx->desc_wt.p, 0, 2 * ((uint64_t)wt_scan_tiles + 1) * sizeof(uint64_t), s));
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_INDEX], s));

        auto launch_scan = [&]() -> int {
        // K2 first: the persistent scan kernel takes its one CTA per SM (and nearly all of its registers);
        // the join (K1), on a second stream, fills SMs as scan CTAs retire and overlaps the scan's tail
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_CIGAR0], s));
        if (ntiles && L) {
            ScanParams sp;
            sp.tile_first = ctx->tile_first.p; sp.cig_off = ctx->cig_off.p; sp.rd_pre = ctx->rd_pre.p; sp.wt = ctx->wt.p;
            sp.wt_sbase = ctx->wt_sbase.p; sp.evraw = ctx->evraw.p; sp.ctr = ctx->d_ctr; sp.raw_cap = ctx->evraw.cap;
            sp.n_wt = n_wt; sp.neg1 = 0xFFFFFFFFu;
            sp.thr = (std::min<uint32_t>(minlen, (1u << 28) - 1u) << 4) | 15u;      // BAM op lengths have 28 bits
            sp.debug = ctx->scan_debug;
            if (sp.thr >> 31) k_cigar_scan<true><<<scan_grid, kCtaThreads, kScanSmemBytes, s>>>(ctx->tmap, sp);
            else k_cigar_scan<false><<<scan_grid, kCtaThreads, kScanSmemBytes, s>>>(ctx->tmap, sp);
            ++launches;
        }

            return INQ_OK;
        };
        auto launch_join = [&]() -> int {
        // K1: candidate ranges + difference array, then the per-locus segment offsets (second stream)
        cudaStream_t sj = ctx->stream_join;
        CU_TRY(ctx, cudaStreamWaitEvent(sj, ctx->ev[EV_INDEX], 0));
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_JOIN0], sj));
        if (work) {
            k_join_ranges<<<(unsigned)((R + 255) / 256), 256, 0, sj>>>(rv, lv, unphased, ctx->cand_lo.p, ctx->cand_n.p, ctx->delta.p, ctx->d_ctr);
            const unsigned g = std::min<unsigned>(loc_scan_tiles, (unsigned)ctx->sm_count * 4);
            // lcnt[i+1] = number of candidate reads of locus i ; seg_off = exclusive scan of those counts
            k_exclusive_scan<<<g, kXsThreads, 0, sj>>>(ctx->delta.p, ctx->lcnt.p, (uint64_t)L + 1, loc_scan_tiles, ctx->desc_scan.p,
                                                       &ctx->d_ctr->scan_counter[2], nullptr);
            k_exclusive_scan<<<g, kXsThreads, 0, sj>>>(ctx->lcnt.p + 1, ctx->seg_off.p, (uint64_t)L, loc_scan_tiles,
                                                       ctx->desc_scan.p + loc_scan_tiles + 1, &ctx->d_ctr->scan_counter[3], &ctx->d_ctr->flags);
            launches += 3;
        }
            return INQ_OK;
        };
        // measured: scan first 4.35 ms/step, join first 4.39 (the join's CTAs delay the scan's), serial 4.41
        TRY(launch_scan());
        TRY(launch_join());
        cudaStream_t sj = ctx->stream_join;
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_JOIN], sj));
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_CIGAR], s));
        if (ntiles && L) {
            const unsigned g = std::min<unsigned>(wt_scan_tiles, (unsigned)ctx->sm_count * 4);
            k_exclusive_scan2<<<g, kXsThreads, 0, s>>>(ctx->wt.p, n_wt, wt_scan_tiles, ctx->desc_wt.p, ctx->desc_wt.p + wt_scan_tiles + 1,
                                                       &ctx->d_ctr->scan_counter[0], &ctx->d_ctr->flags);
            ++launches;
        }
        // total number of events = last entry of the exclusive scan
        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_total + 2, ctx->wt.p + n_wt, sizeof(uint2), cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_FIXUP], s));
        CU_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev[EV_JOIN], 0));
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_SCAN], s));
        CU_TRY(ctx, cudaGetLastError());

        // the call buffer holds one slot per candidate; its size is only known on the device. It is
        // sized from the previous run when there was one (checked afterwards), otherwise read back now.
        if (work) CU_TRY(ctx, cudaMemcpyAsync(ctx->h_total, ctx->seg_off.p + L, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (work && ctx->vals.cap == 0) {
            CU_TRY(ctx, cudaStreamSynchronize(s));
            TRY(ensure(ctx, ctx->vals, (uint64_t)*ctx->h_total + 1));
        }

        // K2b: filter + window sums + scatter
        if (work) {
            k_pair_eval<<<(unsigned)((R + 255) / 256), 256, sizeof(PairSmem), s>>>(rv, lv, unphased, ctx->cand_lo.p, ctx->cand_n.p,
                                                                   EventSource{ctx->wt.p, ctx->rd_pre.p, ctx->wt_sbase.p, ctx->evraw.p, ctx->evraw.cap},
                                                                   ctx->seg_off.p, ctx->cursor.p,
                                                                   ctx->vals.p, ctx->vals.cap, ctx->d_ctr, ctx->pair_debug);
            ++launches;
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_PAIRS], s));

        // K3: medians, in chunks of the catalog: the result copy of one chunk (second stream, copy engine)
        // runs under the median kernels of the next
        if (L) {
            const int nchunk = L >= (1 << 16) ? kMedianChunks : 1;
            cudaStream_t sc = ctx->stream_join;
            for (int c = 0; c < nchunk; ++c) {
                const uint32_t l0 = (uint32_t)((uint64_t)L * c / nchunk), l1 = (uint32_t)((uint64_t)L * (c + 1) / nchunk);
                k_locus_median<<<(unsigned)(((uint64_t)(l1 - l0) * 32 + 255) / 256), 256, 0, s>>>(l0, l1, c, unphased, support, ctx->seg_off.p, ctx->cursor.p,
                                                                                                  ctx->vals.p, ctx->vals.cap, ctx->t1.p, ctx->t2.p, ctx->valid.p,
                                                                                                  ctx->big_list.p, ctx->d_ctr);
                k_locus_median_big<<<(unsigned)ctx->sm_count * 2, kBigThreads, 0, s>>>(l0, c, unphased, support, ctx->seg_off.p, ctx->cursor.p, ctx->vals.p,
                                                                                      ctx->vals.cap, ctx->t1.p, ctx->t2.p, ctx->valid.p, ctx->big_list.p, ctx->d_ctr);
                launches += 2;
                CU_TRY(ctx, cudaEventRecord(ctx->ev_chunk[c], s));
                CU_TRY(ctx, cudaStreamWaitEvent(sc, ctx->ev_chunk[c], 0));
                const size_t n = l1 - l0;
                CU_TRY(ctx, cudaMemcpyAsync(twice_h1 + l0, ctx->t1.p + l0, n * sizeof(int64_t), cudaMemcpyDeviceToHost, sc));
                CU_TRY(ctx, cudaMemcpyAsync(twice_h2 + l0, ctx->t2.p + l0, n * sizeof(int64_t), cudaMemcpyDeviceToHost, sc));
                CU_TRY(ctx, cudaMemcpyAsync(valid_mask + l0, ctx->valid.p + l0, n, cudaMemcpyDeviceToHost, sc));
            }
            CU_TRY(ctx, cudaEventRecord(ctx->ev_chunk[kMedianChunks], sc));
        }
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_MEDIAN], s));
        CU_TRY(ctx, cudaGetLastError());

        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, s));
        if (L) CU_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev_chunk[kMedianChunks], 0));
        CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_D2H], s));
        CU_TRY(ctx, cudaStreamSynchronize(s));

        const unsigned f = ctx->h_ctr->flags;
        bool retry = false;
        if (f & kFlagEventOverflow) {
            // the event buffers were sized speculatively; the scan still counted every event and slot
            const uint64_t need_raw = ctx->h_ctr->ev_alloc + raw_slack;
            release(ctx->evraw);
            TRY(ensure(ctx, ctx->evraw, need_raw));
            retry = true;
        }
        if (work && (uint64_t)*ctx->h_total + 1 > ctx->vals.cap) {
            release(ctx->vals);
            TRY(ensure(ctx, ctx->vals, (uint64_t)*ctx->h_total + 1));
            retry = true;
        }
        if (retry) continue;
        if (f & kFlagCountOverflow) return fail(ctx, INQ_ERR_TOO_LARGE, "pair or event count exceeds 2^32");
        if (f & kFlagValsOverflow) return fail(ctx, INQ_ERR_STATE, "internal: call buffer overflow");
        if (f & kFlagBadHp)
            return fail(ctx, INQ_ERR_BAD_HP, "read %llu passes the phased filter but carries HP %u, outside {0,1,2} (the reference panics, call.rs:358)",
                        (unsigned long long)ctx->h_ctr->bad_hp_read, ctx->h_ctr->bad_hp_value);
        if (f & kFlagMedianEmpty)
            return fail(ctx, INQ_ERR_MEDIAN_EMPTY, "support == 0 with a bucket without usable calls (the reference panics, call.rs:516)");
        break;
    }
    if (ctx->h_ctr->flags & kFlagEventOverflow) return fail(ctx, INQ_ERR_STATE, "internal: event list overflow persists");
    ctx->last_n_events = (ntiles && L) ? ctx->h_total[3] : 0;      // .y of wt[n_wt]

    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_loci = (uint64_t)L;
        stats->n_reads = R;
        stats->n_cigar_words = C;
        auto total = [&](int k) { uint64_t t = 0; for (int i = 0; i < kStatSlots; ++i) t += ctx->h_ctr->stat[i][k]; return t; };
        stats->n_cigar_words_joined = total(ST_WORDS_JOINED);
        stats->n_reads_joined = total(ST_READS_JOINED);
        stats->n_pairs = total(ST_PAIRS);
        stats->n_candidates = total(ST_CANDIDATES);
        stats->n_events = ctx->last_n_events;
        stats->op_visits = total(ST_OP_VISITS);
        stats->n_kernel_launches = launches;
        stats->n_tiles = ntiles;
        auto el = [&](int a, int b) { float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]); return ms; };
        stats->ms_total = el(EV_START, EV_MEDIAN);
        stats->ms_index = el(EV_START, EV_INDEX);
        stats->ms_join = el(EV_JOIN0, EV_JOIN);          // runs concurrently with the CIGAR scan
        stats->ms_cigar = el(EV_CIGAR0, EV_CIGAR);
        stats->ms_fixup = el(EV_CIGAR, EV_FIXUP);
        stats->ms_scan = el(EV_FIXUP, EV_SCAN);
        stats->ms_pairs = el(EV_SCAN, EV_PAIRS);
        stats->ms_median = el(EV_PAIRS, EV_MEDIAN);
        stats->ms_d2h = el(EV_MEDIAN, EV_D2H);
        stats->ms_h2d = ctx->ms_h2d;
    }
    return INQ_OK;
}

int inq_debug_events(inq_ctx *ctx, uint64_t *n_events, uint32_t *event_pos, int32_t *event_val, uint64_t cap,
                     uint32_t *read_event_off)
{
    if (!ctx) return INQ_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const uint64_t E = ctx->last_n_events, R = ctx->R;
    if (n_events) *n_events = E;
    // the genotyping pass reads events straight from the scan kernel's warp-tile storage; the per-read
    // lists in CIGAR order with absolute anchors are only materialised here
    TRY(ensure(ctx, ctx->ev_off, R + 1));
    if (ctx->events.cap < E + 1) { release(ctx->events); TRY(ensure(ctx, ctx->events, E + 1)); }
    if (E) {
        const uint64_t fix_warps = (R + 1 + 30) / 31;
        k_read_fixup<<<(unsigned)((fix_warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->cig_off.p, ctx->rs.p, R, ctx->wt.p, ctx->rd_pre.p,
                                                                                       ctx->wt_sbase.p, ctx->evraw.p, ctx->evraw.cap,
                                                                                       ctx->events.p, ctx->events.cap, ctx->ev_off.p, ctx->d_ctr);
        CU_TRY(ctx, cudaGetLastError());
    } else {
        CU_TRY(ctx, cudaMemsetAsync(ctx->ev_off.p, 0, (R + 1) * sizeof(uint32_t), ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (read_event_off)
        CU_TRY(ctx, cudaMemcpy(read_event_off, ctx->ev_off.p, (R + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    const uint64_t n = std::min(E, cap);
    if (n && (event_pos || event_val)) {
        uint2 *tmp = (uint2 *)malloc(n * sizeof(uint2));
        if (!tmp) return fail(ctx, INQ_ERR_NOMEM, "host allocation failed");
        cudaError_t e = cudaMemcpy(tmp, ctx->events.p, n * sizeof(uint2), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { free(tmp); return fail(ctx, INQ_ERR_CUDA, "cudaMemcpy: %s", cudaGetErrorString(e)); }
        for (uint64_t i = 0; i < n; ++i) {
            if (event_pos) event_pos[i] = tmp[i].x;
            if (event_val) event_val[i] = (int32_t)tmp[i].y;
        }
        free(tmp);
    }
    return INQ_OK;
}

}  // extern "C"
