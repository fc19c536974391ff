// inq_capi.cu -- extern "C" boundary of libinqcall.so (see include/inqcall.h).
// Host-side orchestration only: device memory, H2D/D2H copies, kernel launches and timing.
// No CPU implementation of the hot path lives here: without a CUDA device every entry fails.
//
// One genotyping pass is a small DAG over four streams:
//   S0  counter memset, k_cigar_scan(k) per range k of the CIGAR stream; with one range (the default) also the prefix
//       sum over its warp-tile totals (k_exclusive_scan2)
//   S1  k_zero, k_join_ranges + two offset scans (next to scan(0)), then -- with several ranges k_exclusive_scan2(k)
//       and -- k_pair_eval(k) as soon as range k is scanned
//   S2  k_locus_median (+ the CTA-path kernel where the last pass had deep loci) over the catalog chunk that became
//       complete with pair(k); the counters go home from here
//   S3  k_push_results: finished chunks stored into the caller's pinned arrays (copy engines when they are not mapped)
// The steady state (same reads, same parameters, same pinned outputs) is replayed from a CUDA graph; the first call
// of a shape and every retry (speculative buffer regrown, CTA-path guess wrong) launch the same DAG directly.
#include "../../include/inqcall.h"
#include "inq_device.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace inq;

// shape of the pass (compile-time; the defaults are what the sweeps in profiles/README.md picked)
#ifndef INQ_STREAM_PRIORITIES
#define INQ_STREAM_PRIORITIES 1       // bit i: stream Si is high priority
#endif
#ifndef INQ_SCAN_FIRST
#define INQ_SCAN_FIRST 0              // 1: enqueue the scan kernels before the join chain
#endif
#ifndef INQ_BIG_ON_S3
#define INQ_BIG_ON_S3 0               // 1: CTA-path medians run on the copy stream instead of between two chunks
#endif

namespace {

thread_local std::string g_create_error;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    uint64_t cap = 0;   // elements
};

// driver entry points (libcuda is not linked; they are fetched through the runtime)
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*mem_reserve_fn)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*mem_create_fn)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
typedef CUresult (*mem_map_fn)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*mem_access_fn)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
typedef CUresult (*mem_unmap_fn)(CUdeviceptr, size_t);
typedef CUresult (*mem_release_fn)(CUmemGenericAllocationHandle);
typedef CUresult (*mem_free_fn)(CUdeviceptr, size_t);
typedef CUresult (*mem_gran_fn)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);

struct DriverApi {
    tmap_encode_fn tmap_encode = nullptr;
    mem_reserve_fn reserve = nullptr;
    mem_create_fn create = nullptr;
    mem_map_fn map = nullptr;
    mem_access_fn set_access = nullptr;
    mem_unmap_fn unmap = nullptr;
    mem_release_fn release = nullptr;
    mem_free_fn addr_free = nullptr;
    mem_gran_fn granularity = nullptr;
    bool vmm = false;
};

void *driver_symbol(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    return fn;
}

const DriverApi &driver()
{
    static DriverApi d = [] {
        DriverApi a;
        a.tmap_encode = (tmap_encode_fn)driver_symbol("cuTensorMapEncodeTiled");
        a.reserve = (mem_reserve_fn)driver_symbol("cuMemAddressReserve");
        a.create = (mem_create_fn)driver_symbol("cuMemCreate");
        a.map = (mem_map_fn)driver_symbol("cuMemMap");
        a.set_access = (mem_access_fn)driver_symbol("cuMemSetAccess");
        a.unmap = (mem_unmap_fn)driver_symbol("cuMemUnmap");
        a.release = (mem_release_fn)driver_symbol("cuMemRelease");
        a.addr_free = (mem_free_fn)driver_symbol("cuMemAddressFree");
        a.granularity = (mem_gran_fn)driver_symbol("cuMemGetAllocationGranularity");
        a.vmm = a.reserve && a.create && a.map && a.set_access && a.unmap && a.release && a.addr_free && a.granularity;
        return a;
    }();
    return d;
}

// The packed CIGAR stream (the one multi-GB buffer) lives in a reserved virtual address range that is
// backed by physical chunks as reads arrive: growing it neither copies the stream nor moves its base
// address (the TMA tensor map and a captured graph stay valid).
struct VmStream {
    CUdeviceptr base = 0;
    size_t reserved = 0, mapped = 0, gran = 0;
    std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks;
};

enum {
    EV_START = 0, EV_END, EV_INDEX, EV_JOIN0, EV_JOIN, EV_MED0, EV_MED1, EV_D2H, EV_H2D0, EV_H2D1,
    EV_SCAN0,                                   // + 2k, 2k+1: start / end of k_cigar_scan(k)
    EV_XS0 = EV_SCAN0 + 2 * kMaxRanges,         // + 2k, 2k+1: start / end of k_exclusive_scan2(k)
    EV_PAIR0 = EV_XS0 + 2 * kMaxRanges,         // + 2k, 2k+1: start / end of k_pair_eval(k)
    EV_COUNT = EV_PAIR0 + 2 * kMaxRanges
};
// dependency-only events (no timing)
enum {
    DEP_FORK = 0, DEP_JOIN_DONE, DEP_S1_DONE, DEP_S2_DONE, DEP_S3_DONE, DEP_ZEROED, DEP_SCAN_ONLY,
    DEP_SCANNED,                                // + k: range k scanned
    DEP_PAIRED = DEP_SCANNED + kMaxRanges,      // + k: pair(k) done
    DEP_CHUNK = DEP_PAIRED + kMaxRanges,        // + c: median chunk c done
    DEP_COUNT = DEP_CHUNK + kMaxMedianChunks
};

struct MedianChunk {
    uint32_t l0, l1;
    int after_range;                            // runs once pair(after_range) is done
};

// How the pass is cut up; rebuilt whenever the reads or the catalog change.
struct Plan {
    bool valid = false;
    uint64_t data_gen = 0;
    int K = 1;
    uint64_t tile_end[kMaxRanges + 1] = {};     // warp-tile boundaries, tile_end[0] = 0
    uint64_t read_end[kMaxRanges + 1] = {};     // reads [read_end[k], read_end[k+1]) are evaluated after range k
    int n_chunks = 0;
    MedianChunk chunk[kMaxMedianChunks];
    bool reads_sorted = false;
    // the CTA-path median kernel of a chunk is only launched where the last completed pass over the same reads had
    // deep loci (it is empty almost everywhere and costs a launch on the S2 chain per chunk); verified after every
    // pass from the per-chunk counters, like the speculative buffer sizes
    bool big_known = false;
    bool need_big[kMaxMedianChunks] = {};
};

struct GraphKey {
    uint64_t data_gen = 0, buf_gen = 0;
    uint32_t minlen = 0, support = 0;
    int unphased = 0, timing = 0;
    const void *o1 = nullptr, *o2 = nullptr, *ov = nullptr;
    bool operator==(const GraphKey &o) const
    {
        return data_gen == o.data_gen && buf_gen == o.buf_gen && minlen == o.minlen && support == o.support &&
               unphased == o.unphased && timing == o.timing && o1 == o.o1 && o2 == o.o2 && ov == o.ov;
    }
};

struct ProbeOut {                               // k_plan_probe result per range boundary
    uint32_t read_end;
    uint32_t contig;                            // of read `read_end` (0xFFFFFFFF if none)
    int32_t ref_start;
    uint32_t pad;
};

}  // namespace

struct inq_ctx {
    int device = 0;
    int sm_count = 0;
    int scan_ctas_per_sm = 1;
    cudaStream_t stream = nullptr;        // S0
    cudaStream_t stream_join = nullptr;   // S1: join, pair evaluation
    cudaStream_t stream_med = nullptr;    // S2: medians
    cudaStream_t stream_copy = nullptr;   // S3: result copies
    uint32_t scan_debug = 0, pair_debug = 0;     // only honoured by builds with -DINQ_TIMING_EXPERIMENTS
    std::string err;

    // options (inq_set_option)
    int opt_ranges = 0;                   // 0 = automatic
    int64_t opt_min_range_tiles = 64 * 1024;
    int opt_max_ranges = 1;               // measured (profiles/README.md, r2 sweeps): co-scheduling pair/median CTAs under the scan is zero-sum on B200
    int opt_graph = 1;
    int opt_timing = 1;
    int opt_median_pieces = 12;       // median chunks per pass (the last one is a quarter piece)
    int64_t opt_min_piece = 1 << 16;  // ... but no piece smaller than this many loci
    unsigned long long *join_trace = nullptr;   // experiments (INQ_JOIN_TRACE=file): per-CTA start / end / SM of k_join_ranges
    uint64_t join_trace_ctas = 0;
    int opt_evict_first = 1;          // k_cigar_scan loads the stream with the L2 evict-first policy
    int opt_join_coop = 1;            // k_join_ranges: warp-cooperative lower bounds (0: two scalar binary searches per read)
    int opt_push_ctas = 8;            // CTAs of k_push_results (PCIe-bound stores: a few SMs' worth is plenty)
    int opt_push_kernel = 1;          // results leave through k_push_results (0: three copy-engine operations per chunk)

    // locus catalog
    int32_t n_contigs = 0;
    int64_t L = 0;
    DevBuf<int64_t> contig_off;
    DevBuf<int32_t> lstart, lend, lpmax;
    std::vector<int64_t> h_contig_off;    // host copies for the plan (which loci are complete after which read)
    std::vector<int32_t> h_pmax, h_lstart;
    void *h_route_stage = nullptr;        // pinned staging of inq_push_reads_routed when the kept reads are scattered

    // reads
    uint64_t R = 0, C = 0;
    DevBuf<int32_t> contig, rs, re;
    DevBuf<uint8_t> mapq, hp, flags;
    DevBuf<uint64_t> cig_off;
    DevBuf<uint32_t> cigar;               // cudaMalloc'ed fallback when the driver has no VMM API
    VmStream vm;
    uint32_t *d_unsorted = nullptr;       // != 0: the pushed reads are not sorted by (contig, ref_start)

    // work buffers
    DevBuf<uint32_t> cand_lo, cand_n, ev_off, delta, lcnt, seg_off, big_list;
    DevBuf<unsigned long long> cursor;
    DevBuf<uint32_t> wt_sbase, tile_first;
    DevBuf<uint2> rd_pre, wt, wtot;
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
    ProbeOut *d_probe = nullptr, *h_probe = nullptr;
    uint64_t *d_probe_tiles = nullptr;
    void *h_stage = nullptr;          // pinned staging of the results when the caller's arrays are pageable
    size_t h_stage_bytes = 0;

    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t dep[DEP_COUNT] = {};
    float ms_h2d = 0.f;
    uint64_t last_n_events = 0;

    uint64_t data_gen = 1, buf_gen = 1;
    Plan plan;
    cudaGraphExec_t graph = nullptr;
    GraphKey graph_key, last_key;
    bool have_last_key = false;
    uint32_t graph_launches = 0;
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

void drop_graph(inq_ctx *ctx)
{
    if (ctx->graph) cudaGraphExecDestroy(ctx->graph);
    ctx->graph = nullptr;
}

// grow-only device buffer; keeps the first `keep` elements
template <typename T>
int ensure(inq_ctx *ctx, DevBuf<T> &b, uint64_t need, uint64_t keep = 0, double growth = 1.0)
{
    if (need <= b.cap) return INQ_OK;
    uint64_t cap = std::max<uint64_t>(need, (uint64_t)(b.cap * growth));
    T *np = nullptr;
    CU_TRY(ctx, cudaMalloc(&np, std::max<uint64_t>(cap, 1) * sizeof(T)));
    if (keep && b.p) {
        cudaError_t e = cudaMemcpyAsync(np, b.p, std::min<uint64_t>(keep, b.cap) * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(np); return fail(ctx, INQ_ERR_CUDA, "device copy failed: %s", cudaGetErrorString(e)); }
    }
    if (b.p) cudaFree(b.p);
    b.p = np;
    b.cap = cap;
    ++ctx->buf_gen;                   // a captured graph holds the old pointer
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

void vm_release(VmStream &vm)
{
    const DriverApi &d = driver();
    size_t off = 0;
    for (auto &c : vm.chunks) {
        d.unmap(vm.base + off, c.second);
        d.release(c.first);
        off += c.second;
    }
    vm.chunks.clear();
    if (vm.base) d.addr_free(vm.base, vm.reserved);
    vm = VmStream();
}

// make sure the first `bytes` of the stream are backed by device memory
int vm_ensure(inq_ctx *ctx, size_t bytes)
{
    VmStream &vm = ctx->vm;
    const DriverApi &d = driver();
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = ctx->device;
    if (!vm.base) {
        if (d.granularity(&vm.gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !vm.gran)
            return fail(ctx, INQ_ERR_CUDA, "cuMemGetAllocationGranularity failed");
        size_t total = 0, free_b = 0;
        CU_TRY(ctx, cudaMemGetInfo(&free_b, &total));
        vm.reserved = round_up(total, vm.gran);              // address space only: the whole device, nothing committed
        if (d.reserve(&vm.base, vm.reserved, 1u << 21, 0, 0) != CUDA_SUCCESS) { vm.base = 0; return fail(ctx, INQ_ERR_NOMEM, "cuMemAddressReserve(%zu) failed", vm.reserved); }
    }
    if (bytes <= vm.mapped) return INQ_OK;
    if (bytes > vm.reserved) return fail(ctx, INQ_ERR_NOMEM, "CIGAR stream of %zu bytes exceeds the device", bytes);
    // geometric chunk sizes keep the number of mappings small (64 MB .. 2 GB)
    size_t want = std::max<size_t>(bytes - vm.mapped, std::min<size_t>(std::max<size_t>(vm.mapped / 2, 64u << 20), 2048ull << 20));
    want = std::min(round_up(want, vm.gran), vm.reserved - vm.mapped);
    CUmemGenericAllocationHandle h;
    if (d.create(&h, want, &prop, 0) != CUDA_SUCCESS) {
        want = round_up(bytes - vm.mapped, vm.gran);         // no room for the slack: take exactly what is needed
        if (d.create(&h, want, &prop, 0) != CUDA_SUCCESS) return fail(ctx, INQ_ERR_NOMEM, "cuMemCreate(%zu) failed", want);
    }
    if (d.map(vm.base + vm.mapped, want, 0, h, 0) != CUDA_SUCCESS) { d.release(h); return fail(ctx, INQ_ERR_CUDA, "cuMemMap failed"); }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof(acc));
    acc.location = prop.location;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.set_access(vm.base + vm.mapped, want, &acc, 1) != CUDA_SUCCESS) {
        d.unmap(vm.base + vm.mapped, want);
        d.release(h);
        return fail(ctx, INQ_ERR_CUDA, "cuMemSetAccess failed");
    }
    vm.chunks.emplace_back(h, want);
    vm.mapped += want;
    return INQ_OK;
}

uint32_t *cigar_ptr(inq_ctx *ctx) { return driver().vmm ? reinterpret_cast<uint32_t *>(ctx->vm.base) : ctx->cigar.p; }
uint64_t cigar_cap_words(inq_ctx *ctx) { return driver().vmm ? ctx->vm.mapped / 4 : ctx->cigar.cap; }

// (re)build the TMA descriptor of the CIGAR stream: uint32 [rows][32], box = one 4 KB warp tile, 128B swizzle
int make_tensor_map(inq_ctx *ctx, uint64_t n_words_padded)
{
    const uint64_t rows = n_words_padded / 32;
    if (ctx->tmap_base == cigar_ptr(ctx) && ctx->tmap_rows == rows) return INQ_OK;
    if (!driver().tmap_encode) return fail(ctx, INQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {32, rows};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, (cuuint32_t)(kWarpTileWords / 32)};    // one warp tile
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = driver().tmap_encode(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, cigar_ptr(ctx), gdim, gstride, box, estride,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, INQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    ctx->tmap_base = cigar_ptr(ctx);
    ctx->tmap_rows = rows;
    ++ctx->buf_gen;                   // the tensor map is a kernel argument of a captured graph
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
    // CIGAR stream is padded with zero words up to a tile boundary (+1 tile of slack). What must survive a
    // regrow is everything inq_push_reads left behind: the words, their zero padding, and the tile_first
    // entries of every warp tile up to the padded end plus the sentinel slot.
    const uint64_t Cpad = R ? round_up(C, kTileWords) : 0;
    const uint64_t need_words = round_up(nC, kTileWords) + kTileWords;
    if (driver().vmm) TRY(vm_ensure(ctx, need_words * sizeof(uint32_t)));
    else TRY(ensure(ctx, ctx->cigar, need_words, Cpad, growth));
    TRY(ensure(ctx, ctx->tile_first, std::max(cigar_cap_words(ctx), need_words) / kWarpTileWords + 2, R ? Cpad / kWarpTileWords + 1 : 0, 1.0));
    return INQ_OK;
}

// ---- small kernels of the orchestration layer ---------------------------------------------------

// reads [r0, r0 + n) continue a (contig, ref_start)-sorted sequence? contig compares unsigned so that
// unmapped reads (tid -1) sort last, as in a coordinate-sorted BAM
__global__ void k_check_sorted(const int32_t *__restrict__ contig, const int32_t *__restrict__ rs, uint64_t r0, uint64_t n,
                               uint32_t *__restrict__ unsorted)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = r0 + i;
        if (r == 0) continue;
        const uint32_t ca = (uint32_t)contig[r - 1], cb = (uint32_t)contig[r];
        if (ca > cb || (ca == cb && rs[r - 1] > rs[r])) *unsorted = 1u;
    }
}

// everything a pass needs zeroed besides the counters, in one launch (six small memsets cost ~15 us of launch latency
// on the critical path of a small workload)
struct ZeroJob { void *p; uint64_t bytes; };
struct ZeroJobs { ZeroJob j[6]; int n; };
// The results of one median chunk go home through a kernel that stores into the caller's pinned (mapped) arrays:
// one launch in place of three copy-engine operations, whose ~10 us apiece of fixed latency is what a chunk's
// transfer costs at the per-rank sizes of a multi-GPU run. 16-byte streaming stores between the aligned edges
// of each array (source and destination share the index; when they do not share the alignment: bytes).
__device__ __forceinline__ void push_bytes(unsigned char *__restrict__ dst, const unsigned char *__restrict__ src, size_t n, uint32_t tid, uint32_t stride)
{
    if (((reinterpret_cast<uintptr_t>(dst) ^ reinterpret_cast<uintptr_t>(src)) & 15u) != 0) {
        if ((((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | n) & 7u) == 0)) {
            for (size_t i = tid; i < n / 8; i += stride) __stcs(reinterpret_cast<unsigned long long *>(dst) + i, reinterpret_cast<const unsigned long long *>(src)[i]);
        } else {
            for (size_t i = tid; i < n; i += stride) dst[i] = src[i];
        }
        return;
    }
    const size_t head = min(n, (size_t)((16u - (reinterpret_cast<uintptr_t>(src) & 15u)) & 15u));
    const size_t nv = (n - head) / 16, tail0 = head + nv * 16;
    if (tid < head) dst[tid] = src[tid];
    const uint4 *sv = reinterpret_cast<const uint4 *>(src + head);
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    for (size_t i = tid; i < nv; i += stride) __stcs(dv + i, sv[i]);
    if (tid < n - tail0) dst[tail0 + tid] = src[tail0 + tid];
}

__global__ void __launch_bounds__(256)
k_push_results(uint32_t l0, uint32_t l1, const int64_t *__restrict__ t1, const int64_t *__restrict__ t2, const uint8_t *__restrict__ valid,
               int64_t *__restrict__ o1, int64_t *__restrict__ o2, uint8_t *__restrict__ ov)
{
    const uint32_t stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = l1 - l0;
    push_bytes(reinterpret_cast<unsigned char *>(o1 + l0), reinterpret_cast<const unsigned char *>(t1 + l0), n * 8, tid, stride);
    push_bytes(reinterpret_cast<unsigned char *>(o2 + l0), reinterpret_cast<const unsigned char *>(t2 + l0), n * 8, tid, stride);
    push_bytes(ov + l0, valid + l0, n, tid, stride);
}

__global__ void k_zero(ZeroJobs jobs)
{
    for (int k = 0; k < jobs.n; ++k) {
        uint4 *p = static_cast<uint4 *>(jobs.j[k].p);
        const uint64_t n16 = jobs.j[k].bytes / 16;
        for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// per range boundary k: the first read that is NOT entirely inside warp tiles [0, tile_end[k]) and where it starts
__global__ void k_plan_probe(const uint32_t *__restrict__ tile_first, const uint64_t *__restrict__ tile_end, int K, uint64_t R,
                             const int32_t *__restrict__ contig, const int32_t *__restrict__ rs, ProbeOut *__restrict__ out)
{
    const int k = threadIdx.x;
    if (k >= K) return;
    // tile_first[t] = first entry of cig_off[0, R] that is >= t * kWarpTileWords: every read before entry f - 1 has
    // its end (= the next read's start) below the boundary
    const uint32_t f = tile_first[tile_end[k]];
    uint64_t b = f ? f - 1u : 0u;
    if (b > R) b = R;
    ProbeOut o;
    o.read_end = (uint32_t)b;
    o.contig = b < R ? (uint32_t)contig[b] : 0xFFFFFFFFu;
    o.ref_start = b < R ? rs[b] : 0;
    o.pad = 0;
    out[k] = o;
}

struct RunParams {
    uint32_t minlen, support;
    int unphased;
    int64_t *o1, *o2;
    uint8_t *ov;
    int timing;                             // 0: none; 1: pass + CIGAR scan (3 event records); 2: every stage (~10, ~5 us apiece in a replayed graph)
    int64_t *d1 = nullptr, *d2 = nullptr;   // device views of o1/o2/ov (mapped pinned memory); null: use the copy engines
    uint8_t *dv = nullptr;
};

// Cut the pass into ranges (see the file header). Needs one small device read-back, so it is cached until
// the reads or the catalog change.
int build_plan(inq_ctx *ctx, uint64_t n_wt)
{
    Plan &pl = ctx->plan;
    if (pl.valid && pl.data_gen == ctx->data_gen) return INQ_OK;
    pl = Plan();
    const uint64_t R = ctx->R;
    const int64_t L = ctx->L;
    int K = ctx->opt_ranges > 0 ? ctx->opt_ranges
                                : (int)std::min<uint64_t>((uint64_t)ctx->opt_max_ranges, n_wt / (uint64_t)std::max<int64_t>(1, ctx->opt_min_range_tiles));
    K = std::max(1, std::min(K, kMaxRanges));
    if ((uint64_t)K > n_wt) K = (int)std::max<uint64_t>(1, n_wt);
    pl.K = K;
    for (int k = 0; k <= K; ++k) pl.tile_end[k] = n_wt * (uint64_t)k / (uint64_t)K;
    pl.read_end[0] = 0;
    pl.read_end[K] = R;
    std::vector<int64_t> done(K + 1, 0);
    done[K] = L;
    if (R && L) {
        cudaStream_t s = ctx->stream;
        if (K > 1) {
            CU_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_tiles, pl.tile_end + 1, (K - 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
            k_plan_probe<<<1, 32, 0, s>>>(ctx->tile_first.p, ctx->d_probe_tiles, K - 1, R, ctx->contig.p, ctx->rs.p, ctx->d_probe);
            CU_TRY(ctx, cudaGetLastError());
            CU_TRY(ctx, cudaMemcpyAsync(ctx->h_probe, ctx->d_probe, (K - 1) * sizeof(ProbeOut), cudaMemcpyDeviceToHost, s));
        }
        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_total + 8, ctx->d_unsorted, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        pl.reads_sorted = ctx->h_total[8] == 0;
        for (int k = 1; k < K; ++k) {
            const ProbeOut &o = ctx->h_probe[k - 1];
            pl.read_end[k] = std::max<uint64_t>(pl.read_end[k - 1], std::min<uint64_t>(o.read_end, R));
            // Sorted reads: every read from read_end[k] on starts at (o.contig, o.ref_start) or later, so a locus whose
            // running max of end (+10) lies at or before that start cannot be a candidate of any of them
            // (k_join_ranges / candidate_range): its calls are complete once pair(k - 1) is done.
            int64_t d = 0;
            if (pl.reads_sorted) {
                if (o.contig >= (uint32_t)ctx->n_contigs) d = L;
                else {
                    const int64_t a = ctx->h_contig_off[o.contig], b = ctx->h_contig_off[o.contig + 1];
                    d = std::upper_bound(ctx->h_pmax.begin() + a, ctx->h_pmax.begin() + b, (int64_t)o.ref_start - 10,
                                         [](int64_t v, int32_t p) { return v < (int64_t)p; }) - ctx->h_pmax.begin();
                }
            }
            done[k] = std::max(done[k - 1], std::min<int64_t>(d, L));
        }
    }
    // median chunks: the loci that became complete with pair(k - 1), cut into pieces so that the copy of one piece
    // runs under the medians of the next; the very last piece is small (its copy is the exposed one)
    const int64_t piece = std::max<int64_t>(ctx->opt_min_piece, (L + ctx->opt_median_pieces - 1) / ctx->opt_median_pieces);
    const int64_t last_piece = std::max<int64_t>(ctx->opt_min_piece / 4, L / (4 * ctx->opt_median_pieces));
    for (int k = 1; k <= K; ++k) {
        int64_t a = done[k - 1], b = done[k];
        if (k == K && L == 0) break;
        while (a < b) {
            int64_t e = std::min(b, a + piece);
            if (k == K) {
                // the last group ends with a short piece
                if (b - a > 2 * last_piece && e > b - last_piece) e = b - last_piece;
            } else if (b - e < piece / 4) {
                e = b;                                        // no tiny trailing piece
            }
            if (pl.n_chunks == kMaxMedianChunks - 1) e = b;
            pl.chunk[pl.n_chunks++] = MedianChunk{(uint32_t)a, (uint32_t)e, k - 1};
            a = e;
        }
    }
    pl.valid = true;
    pl.data_gen = ctx->data_gen;
    return INQ_OK;
}

// Enqueue one whole pass on the four streams, starting and ending on S0. `capturing`: inside a stream
// capture (timing events are recorded as external event nodes).
int enqueue_pass(inq_ctx *ctx, const RunParams &rp, bool capturing, uint32_t *n_launches)
{
    const Plan &pl = ctx->plan;
    const int64_t L = ctx->L;
    const uint64_t R = ctx->R, C = ctx->C;
    cudaStream_t s0 = ctx->stream, s1 = ctx->stream_join, s2 = ctx->stream_med, s3 = ctx->stream_copy;
    const uint32_t ntiles = (uint32_t)((C + kTileWords - 1) / kTileWords);
    const uint64_t n_wt = (uint64_t)ntiles * (kTileWords / kWarpTileWords);
    const uint32_t loc_scan_tiles = (uint32_t)(((uint64_t)L + 1 + kXsTile - 1) / kXsTile);
    const bool work = R && L;
    uint32_t launches = 0;
    auto stamp = [&](int e, cudaStream_t s, int level = 2) -> cudaError_t {
        if (rp.timing < level) return cudaSuccess;
        return cudaEventRecordWithFlags(ctx->ev[e], s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
    };
    ReadView rv{ctx->contig.p, ctx->rs.p, ctx->re.p, ctx->mapq.p, ctx->hp.p, ctx->flags.p, ctx->cig_off.p, R};
    LocusView lv{ctx->contig_off.p, ctx->lstart.p, ctx->lend.p, ctx->lpmax.p, ctx->n_contigs};

    // S0 only needs the counters zeroed before the scan starts; everything the join / pair / median kernels need
    // zeroed is cleared on S1, off the scan's critical path
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DevCounters), s0));
    CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_FORK], s0));
    // (EV_START / EV_END are plain event records on S0 around the pass -- or around the graph launch --, made by
    // inq_genotype: no nodes of the graph)

    auto enqueue_join = [&]() -> int {
        // ---- S1: K1 candidate ranges + difference array, then the per-locus segment offsets
        CU_TRY(ctx, cudaStreamWaitEvent(s1, ctx->dep[DEP_FORK], 0));
        {
            // (buffers come from cudaMalloc: 256-byte aligned; the sizes are rounded up to 16 bytes inside their capacity)
            ZeroJobs z;
            z.n = 0;
            auto add = [&](void *p, uint64_t bytes, uint64_t cap_bytes) { if (p && bytes) z.j[z.n++] = ZeroJob{p, std::min((bytes + 15) / 16 * 16, cap_bytes / 16 * 16)}; };
            uint64_t biggest = 0;
            if (L) {
                add(ctx->delta.p, ((uint64_t)L + 2) * sizeof(uint32_t), ctx->delta.cap * sizeof(uint32_t));
                add(ctx->seg_off.p, ((uint64_t)L + 2) * sizeof(uint32_t), ctx->seg_off.cap * sizeof(uint32_t));
                add(ctx->cursor.p, ((uint64_t)L + 1) * sizeof(unsigned long long), ctx->cursor.cap * sizeof(unsigned long long));
                add(ctx->desc_scan.p, 2 * ((uint64_t)loc_scan_tiles + 1) * sizeof(uint64_t), ctx->desc_scan.cap * sizeof(uint64_t));
                biggest = ((uint64_t)L + 1) * sizeof(unsigned long long);
            }
            if (!ntiles) add(ctx->wt.p, 2 * sizeof(uint2), ctx->wt.cap * sizeof(uint2));      // no CIGAR words at all
            if (n_wt) add(ctx->desc_wt.p, ctx->desc_wt.cap * sizeof(uint64_t), ctx->desc_wt.cap * sizeof(uint64_t));
            if (z.n) {
                const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((std::max<uint64_t>(biggest, 4096) / 16 + 255) / 256, (uint64_t)ctx->sm_count * 4));
                k_zero<<<g, 256, 0, s1>>>(z);
                CU_TRY(ctx, cudaGetLastError());
            }
            CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_ZEROED], s1));
            CU_TRY(ctx, stamp(EV_INDEX, s1));
        }
        if (work) {
            k_join_ranges<<<(unsigned)((R + 255) / 256), 256, 0, s1>>>(rv, lv, rp.unphased, ctx->opt_join_coop, ctx->cand_lo.p, ctx->cand_n.p, ctx->delta.p, ctx->d_ctr, ctx->join_trace);
            CU_TRY(ctx, stamp(EV_JOIN0, s1));
            const unsigned g = std::min<unsigned>(loc_scan_tiles, (unsigned)ctx->sm_count * 4);
            // lcnt[i+1] = number of candidate reads of locus i ; seg_off = exclusive scan of those counts
            k_exclusive_scan<<<g, kXsThreads, 0, s1>>>(ctx->delta.p, ctx->lcnt.p, (uint64_t)L + 1, loc_scan_tiles, ctx->desc_scan.p,
                                                       &ctx->d_ctr->scan_counter[2], nullptr);
            k_exclusive_scan<<<g, kXsThreads, 0, s1>>>(ctx->lcnt.p + 1, ctx->seg_off.p, (uint64_t)L, loc_scan_tiles,
                                                       ctx->desc_scan.p + loc_scan_tiles + 1, &ctx->d_ctr->scan_counter[3], &ctx->d_ctr->flags);
            launches += 3;
            CU_TRY(ctx, cudaGetLastError());
            // the call buffer holds one slot per candidate; its size is only known on the device (checked after the pass)
            // (read home on S2, idle until the first medians: a copy-engine operation in front of k_pair_eval costs the chain ~10 us)
            CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_JOIN_DONE], s1));
            CU_TRY(ctx, cudaStreamWaitEvent(s2, ctx->dep[DEP_JOIN_DONE], 0));
            CU_TRY(ctx, cudaMemcpyAsync(ctx->h_total, ctx->seg_off.p + L, sizeof(uint32_t), cudaMemcpyDeviceToHost, s2));
        }
        CU_TRY(ctx, stamp(EV_JOIN, s1));

        return INQ_OK;
    };
    uint64_t desc_base = 0;
    auto enqueue_xscan2 = [&](int k, cudaStream_t st) -> int {
        const uint64_t t0 = pl.tile_end[k], t1 = pl.tile_end[k + 1];
        if (t1 > t0 && L) {
            const uint64_t n = t1 - t0;
            const uint32_t xt = (uint32_t)((n + kXsTile - 1) / kXsTile);
            const unsigned g = std::min<unsigned>(xt, (unsigned)ctx->sm_count * 4);
            uint64_t *dx = ctx->desc_wt.p + desc_base, *dy = dx + xt + 1;
            desc_base += 2 * ((uint64_t)xt + 1);
            k_exclusive_scan2<<<g, kXsThreads, 0, st>>>(ctx->wtot.p + t0, ctx->wt.p + t0, n, xt, dx, dy, &ctx->d_ctr->wt_scan_counter[k],
                                                        ctx->d_ctr->wt_carry[k], ctx->d_ctr->wt_carry[k + 1], &ctx->d_ctr->flags);
            ++launches;
            CU_TRY(ctx, cudaGetLastError());
        }
        return INQ_OK;
    };
    const bool xs_on_s0 = pl.K == 1 && !INQ_SCAN_FIRST;
    auto enqueue_scans = [&]() -> int {
        // ---- S0: K2, range after range, nothing in between (the scan of a range depends on nothing but the reads)
        for (int k = 0; k < pl.K; ++k) {
            const uint64_t t0 = pl.tile_end[k], t1 = pl.tile_end[k + 1];
            CU_TRY(ctx, stamp(EV_SCAN0 + 2 * k, s0));
            if (t1 > t0 && L) {
                ScanParams sp;
                sp.tile_first = ctx->tile_first.p; sp.cig_off = ctx->cig_off.p; sp.rd_pre = ctx->rd_pre.p; sp.wt = ctx->wtot.p;
                sp.wt_sbase = ctx->wt_sbase.p; sp.evraw = ctx->evraw.p; sp.ctr = ctx->d_ctr; sp.raw_cap = ctx->evraw.cap;
                sp.wt_begin = t0; sp.n_wt = t1; sp.neg1 = 0xFFFFFFFFu;
                sp.thr = (std::min<uint32_t>(rp.minlen, (1u << 28) - 1u) << 4) | 15u;      // BAM op lengths have 28 bits
                sp.evict_first = (uint32_t)ctx->opt_evict_first;
                sp.debug = ctx->scan_debug;
                const unsigned grid = (unsigned)std::min<uint64_t>((t1 - t0 + kScanWarps - 1) / kScanWarps, (uint64_t)ctx->sm_count * ctx->scan_ctas_per_sm);
                if (sp.thr >> 31) k_cigar_scan<true><<<grid, kCtaThreads, kScanSmemBytes, s0>>>(ctx->tmap, sp);
                else k_cigar_scan<false><<<grid, kCtaThreads, kScanSmemBytes, s0>>>(ctx->tmap, sp);
                ++launches;
                CU_TRY(ctx, cudaGetLastError());
            }
            if (xs_on_s0) {
                // one range: the prefix sum over the warp-tile totals follows the scan on its own stream, so the join chain
                // on S1 (slow under the scan) has until the end of it; the scan's end is time-stamped on the idle S3
                CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_SCAN_ONLY], s0));
                if (rp.timing >= 1) {
                    CU_TRY(ctx, cudaStreamWaitEvent(s3, ctx->dep[DEP_SCAN_ONLY], 0));
                    CU_TRY(ctx, stamp(EV_SCAN0 + 2 * k + 1, s3, 1));
                }
                CU_TRY(ctx, cudaStreamWaitEvent(s0, ctx->dep[DEP_ZEROED], 0));       // desc_wt is cleared by k_zero
                TRY(enqueue_xscan2(k, s0));
                CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_SCANNED + k], s0));
            } else {
                // the dependants hang off the kernel, not off the event-record node that follows it
                CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_SCANNED + k], s0));
                CU_TRY(ctx, stamp(EV_SCAN0 + 2 * k + 1, s0, k == pl.K - 1 ? 1 : 2));
            }
        }

        return INQ_OK;
    };
    // which of the two is enqueued first decides whose CTAs the SMs see first
    if (INQ_SCAN_FIRST) { TRY(enqueue_scans()); TRY(enqueue_join()); }
    else { TRY(enqueue_join()); TRY(enqueue_scans()); }

    // ---- S1: prefix sum over the range's warp-tile totals, then K2b ; S2: K3 per finished catalog chunk ; S3: result copies
    int c = 0;
    int64_t *o1 = rp.o1, *o2 = rp.o2;
    uint8_t *ov = rp.ov;
    for (int k = 0; k < pl.K; ++k) {
        const uint64_t r0 = pl.read_end[k], r1 = pl.read_end[k + 1];
        CU_TRY(ctx, cudaStreamWaitEvent(s1, ctx->dep[DEP_SCANNED + k], 0));
        if (!xs_on_s0) TRY(enqueue_xscan2(k, s1));
        CU_TRY(ctx, stamp(EV_PAIR0 + 2 * k, s1));
        if (work && r1 > r0) {
            const unsigned threads = kPairWarps * 32;
            k_pair_eval<<<(unsigned)((r1 - r0 + threads - 1) / threads), threads, sizeof(PairSmem), s1>>>(
                rv, r0, r1, lv, rp.unphased, ctx->cand_lo.p, ctx->cand_n.p,
                EventSource{ctx->wt.p, ctx->rd_pre.p, ctx->wt_sbase.p, ctx->evraw.p, ctx->evraw.cap}, ctx->seg_off.p, ctx->cursor.p,
                ctx->vals.p, ctx->vals.cap, ctx->d_ctr, ctx->pair_debug);
            ++launches;
            CU_TRY(ctx, cudaGetLastError());
        }
        CU_TRY(ctx, stamp(EV_PAIR0 + 2 * k + 1, s1));
        CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_PAIRED + k], s1));
        bool waited = false;
        for (; c < pl.n_chunks && pl.chunk[c].after_range == k; ++c) {
            const uint32_t l0 = pl.chunk[c].l0, l1 = pl.chunk[c].l1;
            if (!waited) {
                CU_TRY(ctx, cudaStreamWaitEvent(s2, ctx->dep[DEP_PAIRED + k], 0));
                // the first chunk also needs the (empty) segments of loci without reads: wait for the join if nothing was paired
                waited = true;
            }
            if (c == 0) CU_TRY(ctx, stamp(EV_MED0, s2));
            k_locus_median<<<(unsigned)((((uint64_t)(l1 - l0) + kMedianLociPerWarp - 1) / kMedianLociPerWarp * 32 + 255) / 256), 256, 0, s2>>>(l0, l1, c, rp.unphased, rp.support, ctx->seg_off.p, ctx->cursor.p,
                                                                                              ctx->vals.p, ctx->vals.cap, ctx->t1.p, ctx->t2.p, ctx->valid.p,
                                                                                              ctx->big_list.p, ctx->d_ctr);
            CU_TRY(ctx, cudaGetLastError());
            // the CTA path for the (rare) deep loci of the chunk: between two chunks on S2, or on the copy stream
            cudaStream_t sb = INQ_BIG_ON_S3 ? s3 : s2;
            if (INQ_BIG_ON_S3) {
                CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_CHUNK + c], s2));
                CU_TRY(ctx, cudaStreamWaitEvent(s3, ctx->dep[DEP_CHUNK + c], 0));
            }
            if (!pl.big_known || pl.need_big[c]) {
                k_locus_median_big<<<(unsigned)ctx->sm_count * 2, kBigThreads, 0, sb>>>(l0, c, rp.unphased, rp.support, ctx->seg_off.p, ctx->cursor.p, ctx->vals.p,
                                                                                      ctx->vals.cap, ctx->t1.p, ctx->t2.p, ctx->valid.p, ctx->big_list.p, ctx->d_ctr);
                ++launches;
            }
            ++launches;
            CU_TRY(ctx, cudaGetLastError());
            if (!INQ_BIG_ON_S3) {
                CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_CHUNK + c], s2));
                CU_TRY(ctx, cudaStreamWaitEvent(s3, ctx->dep[DEP_CHUNK + c], 0));
            }
            const size_t n = l1 - l0;
            if (rp.d1) {
                const unsigned g = (unsigned)std::min<size_t>((size_t)ctx->opt_push_ctas, (n + 2047) / 2048);
                k_push_results<<<std::max(g, 1u), 256, 0, s3>>>(l0, l1, ctx->t1.p, ctx->t2.p, ctx->valid.p, rp.d1, rp.d2, rp.dv);
                ++launches;
                CU_TRY(ctx, cudaGetLastError());
            } else {
                CU_TRY(ctx, cudaMemcpyAsync(o1 + l0, ctx->t1.p + l0, n * sizeof(int64_t), cudaMemcpyDeviceToHost, s3));
                CU_TRY(ctx, cudaMemcpyAsync(o2 + l0, ctx->t2.p + l0, n * sizeof(int64_t), cudaMemcpyDeviceToHost, s3));
                CU_TRY(ctx, cudaMemcpyAsync(ov + l0, ctx->valid.p + l0, n, cudaMemcpyDeviceToHost, s3));
            }
        }
    }
    if (pl.n_chunks == 0) {                                   // nothing to do on S2/S3: keep them in the DAG for the join below
        CU_TRY(ctx, cudaStreamWaitEvent(s2, ctx->dep[DEP_FORK], 0));
        CU_TRY(ctx, stamp(EV_MED0, s2));
        CU_TRY(ctx, cudaStreamWaitEvent(s3, ctx->dep[DEP_FORK], 0));
    }
    CU_TRY(ctx, stamp(EV_MED1, s2));
    // total number of events = the last prefix
    if (n_wt && L) CU_TRY(ctx, cudaMemcpyAsync(ctx->h_total + 2, ctx->wt.p + n_wt, sizeof(uint2), cudaMemcpyDeviceToHost, s1));
    CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_S1_DONE], s1));
    // the counters go home on S2 next to the last result copy (S3) instead of after it: S2 has seen every median kernel,
    // S1 every join / prefix / pair kernel, and those waited for the scans
    cudaStream_t sctr = INQ_BIG_ON_S3 ? s3 : s2;              // (the stream that runs the CTA-path medians, which also raise flags)
    CU_TRY(ctx, cudaStreamWaitEvent(sctr, ctx->dep[DEP_S1_DONE], 0));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, sctr));
    CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_S2_DONE], s2));
    CU_TRY(ctx, stamp(EV_D2H, s3));
    CU_TRY(ctx, cudaEventRecord(ctx->dep[DEP_S3_DONE], s3));

    // ---- join everything on S0
    CU_TRY(ctx, cudaStreamWaitEvent(s0, ctx->dep[DEP_S1_DONE], 0));
    CU_TRY(ctx, cudaStreamWaitEvent(s0, ctx->dep[DEP_S2_DONE], 0));
    CU_TRY(ctx, cudaStreamWaitEvent(s0, ctx->dep[DEP_S3_DONE], 0));
    *n_launches = launches;
    return INQ_OK;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

const char *inq_version(void) { return "inquistr-b200 0.2.0 (sm_100a)"; }

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
    // the HBM-bound scan gets the SMs first; pair / median CTAs fill what its one CTA per SM leaves free
    int prio_lo = 0, prio_hi = 0;
    if ((e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)) != cudaSuccess) return bail("cudaDeviceGetStreamPriorityRange", e);
    cudaStream_t *streams[4] = {&ctx->stream, &ctx->stream_join, &ctx->stream_med, &ctx->stream_copy};
    for (int i = 0; i < 4; ++i) {
        // INQ_STREAM_PRIORITIES bit i: stream Si gets the high priority (default: only the scan stream S0)
        int prio_mask = INQ_STREAM_PRIORITIES;
        if (const char *pm = getenv("INQ_STREAM_PRIORITY_MASK")) prio_mask = atoi(pm);       // experiments
        const bool hi = (prio_mask >> i) & 1;
        if ((e = cudaStreamCreateWithPriority(streams[i], cudaStreamNonBlocking, hi ? prio_hi : prio_lo)) != cudaSuccess) return bail("cudaStreamCreate", e);
    }
    for (int i = 0; i < EV_COUNT; ++i)
        if ((e = cudaEventCreate(&ctx->ev[i])) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int i = 0; i < DEP_COUNT; ++i)
        if ((e = cudaEventCreateWithFlags(&ctx->dep[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->d_ctr, sizeof(DevCounters))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_unsorted, sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMemset(ctx->d_unsorted, 0, sizeof(uint32_t))) != cudaSuccess) return bail("cudaMemset", e);
    if ((e = cudaMalloc(&ctx->d_probe, kMaxRanges * sizeof(ProbeOut))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_probe_tiles, kMaxRanges * sizeof(uint64_t))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMallocHost(&ctx->h_ctr, sizeof(DevCounters))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost(&ctx->h_total, 64)) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost(&ctx->h_probe, kMaxRanges * sizeof(ProbeOut))) != cudaSuccess) return bail("cudaMallocHost", e);
    memset(ctx->h_ctr, 0, sizeof(DevCounters));
    if ((e = cudaFuncSetAttribute(k_cigar_scan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmemBytes)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_cigar_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanSmemBytes)) != cudaSuccess)
        return bail("cudaFuncSetAttribute(k_cigar_scan)", e);
    if ((e = cudaFuncSetAttribute(k_pair_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairSmem))) != cudaSuccess)
        return bail("cudaFuncSetAttribute(k_pair_eval)", e);
    int occ = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cigar_scan<false>, kCtaThreads, kScanSmemBytes)) != cudaSuccess)
        return bail("occupancy(k_cigar_scan)", e);
    ctx->scan_ctas_per_sm = std::max(1, occ);
#ifdef INQ_TIMING_EXPERIMENTS
    if (const char *dbg = getenv("INQ_SCAN_DEBUG")) ctx->scan_debug = (uint32_t)atoi(dbg);
    if (const char *dbg = getenv("INQ_PAIR_DEBUG")) ctx->pair_debug = (uint32_t)atoi(dbg);
#endif
    *out = ctx;
    return INQ_OK;
}

void inq_ctx_destroy(inq_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    drop_graph(ctx);
    release(ctx->contig_off); release(ctx->lstart); release(ctx->lend); release(ctx->lpmax);
    release(ctx->contig); release(ctx->rs); release(ctx->re);
    release(ctx->mapq); release(ctx->hp); release(ctx->flags);
    release(ctx->cig_off); release(ctx->cigar);
    if (ctx->vm.base) vm_release(ctx->vm);
    release(ctx->cand_lo); release(ctx->cand_n); release(ctx->ev_off);
    release(ctx->rd_pre); release(ctx->wt); release(ctx->wtot); release(ctx->wt_sbase); release(ctx->tile_first);
    release(ctx->desc_wt); release(ctx->evraw);
    release(ctx->delta); release(ctx->lcnt); release(ctx->seg_off); release(ctx->cursor); release(ctx->big_list);
    release(ctx->desc_scan); release(ctx->vals);
    release(ctx->events); release(ctx->t1); release(ctx->t2); release(ctx->valid);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->d_unsorted) cudaFree(ctx->d_unsorted);
    if (ctx->d_probe) cudaFree(ctx->d_probe);
    if (ctx->d_probe_tiles) cudaFree(ctx->d_probe_tiles);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_total) cudaFreeHost(ctx->h_total);
    if (ctx->h_probe) cudaFreeHost(ctx->h_probe);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_route_stage) cudaFreeHost(ctx->h_route_stage);
    for (int i = 0; i < EV_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < DEP_COUNT; ++i)
        if (ctx->dep[i]) cudaEventDestroy(ctx->dep[i]);
    cudaStream_t streams[4] = {ctx->stream_copy, ctx->stream_med, ctx->stream_join, ctx->stream};
    for (auto s : streams)
        if (s) cudaStreamDestroy(s);
    delete ctx;
}

int inq_set_option(inq_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return INQ_ERR_ARG;
    const std::string n(name);
    if (n == "ranges") {
        if (value < 0 || value > kMaxRanges) return fail(ctx, INQ_ERR_ARG, "ranges must be in [0, %d]", kMaxRanges);
        ctx->opt_ranges = (int)value;
    } else if (n == "max_ranges") {
        if (value < 1 || value > kMaxRanges) return fail(ctx, INQ_ERR_ARG, "max_ranges must be in [1, %d]", kMaxRanges);
        ctx->opt_max_ranges = (int)value;
    } else if (n == "min_range_tiles") {
        if (value < 1) return fail(ctx, INQ_ERR_ARG, "min_range_tiles must be positive");
        ctx->opt_min_range_tiles = value;
    } else if (n == "graph") ctx->opt_graph = value != 0;
    else if (n == "timing") {
        if (value < 0 || value > 2) return fail(ctx, INQ_ERR_ARG, "timing must be 0 (off), 1 (pass and CIGAR scan) or 2 (every stage)");
        ctx->opt_timing = (int)value;
    }
    else if (n == "push_kernel") ctx->opt_push_kernel = value != 0;
    else if (n == "join_coop") ctx->opt_join_coop = value != 0;
    else if (n == "evict_first") ctx->opt_evict_first = value != 0;
    else if (n == "push_ctas") {
        if (value < 1 || value > 1024) return fail(ctx, INQ_ERR_ARG, "push_ctas must be in [1, 1024]");
        ctx->opt_push_ctas = (int)value;
    }
    else if (n == "median_pieces") {
        if (value < 1 || value > kMaxMedianChunks / 2) return fail(ctx, INQ_ERR_ARG, "median_pieces must be in [1, %d]", kMaxMedianChunks / 2);
        ctx->opt_median_pieces = (int)value;
    } else if (n == "min_piece") {
        if (value < 256) return fail(ctx, INQ_ERR_ARG, "min_piece must be at least 256 loci");
        ctx->opt_min_piece = value;
    }
    else return fail(ctx, INQ_ERR_ARG, "unknown option '%s'", name);
    ctx->plan.valid = false;
    drop_graph(ctx);
    ctx->have_last_key = false;
    return INQ_OK;
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
    ++ctx->data_gen;
    TRY(ensure(ctx, ctx->contig_off, (uint64_t)n_contigs + 1));
    TRY(ensure(ctx, ctx->lstart, (uint64_t)L));
    TRY(ensure(ctx, ctx->lend, (uint64_t)L));
    TRY(ensure(ctx, ctx->lpmax, (uint64_t)L));
    TRY(ensure(ctx, ctx->delta, (uint64_t)L + 2 + 4));            // (+ slack: k_zero clears in 16-byte units)
    TRY(ensure(ctx, ctx->lcnt, (uint64_t)L + 3));
    TRY(ensure(ctx, ctx->seg_off, (uint64_t)L + 2 + 4));
    TRY(ensure(ctx, ctx->cursor, (uint64_t)L + 1 + 2));
    TRY(ensure(ctx, ctx->big_list, (uint64_t)L + 1));
    TRY(ensure(ctx, ctx->desc_scan, 2 * (((uint64_t)L + 2 + kXsTile - 1) / kXsTile + 1) + 2));
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
    // host copy of the running max of `end` per contig (the plan's "which loci are complete" test); overlaps the copies above
    ctx->h_contig_off.assign(contig_locus_offsets, contig_locus_offsets + n_contigs + 1);
    ctx->h_pmax.resize((size_t)L);
    ctx->h_lstart.assign(start, start + (L ? L : 0));
    for (int32_t c = 0; c < n_contigs; ++c) {
        int32_t m = INT32_MIN;
        for (int64_t i = contig_locus_offsets[c]; i < contig_locus_offsets[c + 1]; ++i) { m = std::max(m, end[i]); ctx->h_pmax[(size_t)i] = m; }
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
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    ctx->R = 0;
    ctx->C = 0;
    ++ctx->data_gen;
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_unsorted, 0, sizeof(uint32_t), ctx->stream));
    return INQ_OK;
}

// Appends n reads whose CIGAR words are words[cigar_off[0] .. cigar_off[n]) (cigar_off need not start at 0).
// The copies are asynchronous; the caller synchronises ctx->stream before the host arrays may change.
static int push_impl(inq_ctx *ctx, uint64_t n, const int32_t *contig, const int32_t *ref_start, const int32_t *ref_end,
                     const uint8_t *mapq, const uint8_t *hp, const uint8_t *flags, const uint64_t *cigar_off, const uint32_t *words)
{
    const uint64_t w0 = cigar_off[0], nw = cigar_off[n] - w0;
    if (ctx->R + n >= 0xFFFFFFFFull) return fail(ctx, INQ_ERR_TOO_LARGE, "inq_push_reads: more than 2^32-2 reads");
    TRY(reserve_reads(ctx, ctx->R + n, ctx->C + nw, 1.5));
    ++ctx->data_gen;
    cudaStream_t s = ctx->stream;
    const uint64_t R0 = ctx->R, C0 = ctx->C;
    uint32_t *cig = cigar_ptr(ctx);
    CU_TRY(ctx, cudaMemcpyAsync(ctx->contig.p + R0, contig, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->rs.p + R0, ref_start, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->re.p + R0, ref_end, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->mapq.p + R0, mapq, n, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->hp.p + R0, hp, n, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->flags.p + R0, flags, n, cudaMemcpyHostToDevice, s));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->cig_off.p + R0, cigar_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (nw) CU_TRY(ctx, cudaMemcpyAsync(cig + C0, words + w0, nw * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    if (C0 != w0) {                                              // batch-local offsets -> offsets into the device stream (mod 2^64)
        k_rebase_offsets<<<(unsigned)std::min<uint64_t>((n + 1 + 255) / 256, 8192), 256, 0, s>>>(ctx->cig_off.p + R0, n + 1, C0 - w0);
        CU_TRY(ctx, cudaGetLastError());
    }
    k_check_sorted<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, s>>>(ctx->contig.p, ctx->rs.p, R0, n, ctx->d_unsorted);
    CU_TRY(ctx, cudaGetLastError());
    const uint64_t C1 = C0 + nw, Cpad = round_up(C1, kTileWords);
    if (Cpad > C1) CU_TRY(ctx, cudaMemsetAsync(cig + C1, 0, (Cpad - C1) * sizeof(uint32_t), s));
    {
        // reads starting per warp tile, for the tiles that gained words (k_cigar_scan leaves the tile-local
        // prefix of every read start in rd_pre); cig_off[R0 + n] is the sentinel and counts as a start
        const uint64_t t0 = C0 / kWarpTileWords, t1 = Cpad / kWarpTileWords;
        k_tile_first<<<(unsigned)((t1 - t0 + 1 + 255) / 256), 256, 0, s>>>(ctx->cig_off.p, R0 + n + 1, t0, t1, ctx->tile_first.p);
        CU_TRY(ctx, cudaGetLastError());
    }
    ctx->R = R0 + n;
    ctx->C = C1;
    return INQ_OK;
}

static int check_push_args(inq_ctx *ctx, uint64_t n, const int32_t *contig, const int32_t *ref_start, const int32_t *ref_end,
                           const uint8_t *mapq, const uint8_t *hp, const uint8_t *flags, const uint64_t *cigar_off, const uint32_t *cigar_words)
{
    if (!contig || !ref_start || !ref_end || !mapq || !hp || !flags || !cigar_off) return fail(ctx, INQ_ERR_ARG, "inq_push_reads: NULL array");
    if (cigar_off[0] != 0) return fail(ctx, INQ_ERR_ARG, "inq_push_reads: cigar_off[0] must be 0");
    if (cigar_off[n] && !cigar_words) return fail(ctx, INQ_ERR_ARG, "inq_push_reads: cigar_words is NULL");
    return INQ_OK;
}

int inq_push_reads(inq_ctx *ctx, uint64_t n, const int32_t *contig, const int32_t *ref_start, const int32_t *ref_end,
                   const uint8_t *mapq, const uint8_t *hp, const uint8_t *flags, const uint64_t *cigar_off,
                   const uint32_t *cigar_words)
{
    if (!ctx) return INQ_ERR_ARG;
    if (n == 0) return INQ_OK;
    TRY(check_push_args(ctx, n, contig, ref_start, ref_end, mapq, hp, flags, cigar_off, cigar_words));
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool first = ctx->R == 0;
    CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_H2D0], s));
    TRY(push_impl(ctx, n, contig, ref_start, ref_end, mapq, hp, flags, cigar_off, cigar_words));
    CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_H2D1], s));
    CU_TRY(ctx, cudaStreamSynchronize(s));       // host arrays may be reused by the caller after return
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[EV_H2D0], ctx->ev[EV_H2D1]);
    ctx->ms_h2d = (first ? 0.f : ctx->ms_h2d) + ms;
    return INQ_OK;
}

// Push only the reads of a batch that can matter to THIS context's catalog (see include/inqcall.h). With N contexts
// holding N contiguous catalog shards, handing every batch to every context routes the reads (reads at a cut go to
// both sides) without the caller knowing the cuts -- the host-side half of SURVEY 8e.
int inq_push_reads_routed(inq_ctx *ctx, uint64_t n, const int32_t *contig, const int32_t *ref_start, const int32_t *ref_end,
                          const uint8_t *mapq, const uint8_t *hp, const uint8_t *flags, const uint64_t *cigar_off,
                          const uint32_t *cigar_words, uint32_t filter, int host_threads, uint64_t *n_taken)
{
    if (!ctx) return INQ_ERR_ARG;
    if (n_taken) *n_taken = 0;
    if (n == 0) return INQ_OK;
    TRY(check_push_args(ctx, n, contig, ref_start, ref_end, mapq, hp, flags, cigar_off, cigar_words));
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const int nt = std::max(1, std::min(host_threads > 0 ? host_threads : (int)std::thread::hardware_concurrency(), 256));
    // 1. which reads would htslib's fetch return for some locus of this catalog (pos < end+10 && endpos > start-10,
    //    call.rs:285-288), minus the ones the per-read half of the filters rejects everywhere (call.rs:297-300,350-352)
    std::vector<uint8_t> keep(n);
    const int64_t *coff = ctx->h_contig_off.data();
    const int32_t *ls = ctx->h_lstart.data(), *pm = ctx->h_pmax.data();
    const int32_t n_contigs = ctx->n_contigs;
    auto mark = [&](uint64_t a, uint64_t b) {
        for (uint64_t i = a; i < b; ++i) {
            uint8_t k = 0;
            const int32_t c = contig[i];
            if (c >= 0 && c < n_contigs && coff[c] != coff[c + 1] && !((filter & INQ_ROUTE_DROP_LOW_MAPQ) && mapq[i] <= 10) &&
                !((filter & INQ_ROUTE_DROP_NO_HP) && hp[i] == INQ_HP_ABSENT)) {
                const int32_t *b0 = ls + coff[c], *b1 = ls + coff[c + 1];
                const int64_t hi = std::lower_bound(b0, b1, (int32_t)std::min<int64_t>((int64_t)ref_end[i] + 10, INT32_MAX)) - ls;
                k = hi != coff[c] && (int64_t)pm[hi - 1] + 10 > (int64_t)ref_start[i];
            }
            keep[i] = k;
        }
    };
    auto parallel = [&](uint64_t total, auto fn) {
        const uint64_t per = (total + nt - 1) / nt;
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t)
            if ((uint64_t)t * per < total) th.emplace_back([&, t] { fn((uint64_t)t * per, std::min(total, (uint64_t)(t + 1) * per)); });
        fn(0, std::min(total, per));
        for (auto &x : th) x.join();
    };
    parallel(n, mark);
    // 2. runs of kept reads. Short gaps are bridged (a read that reaches nothing is harmless on the device: the join
    //    gives it no candidates, and shipping it is cheaper than gathering around it): coordinate-sorted input and a
    //    contiguous catalog range give one run per contig
    constexpr uint64_t kBridge = 256;
    std::vector<std::pair<uint64_t, uint64_t>> runs;
    uint64_t taken = 0;
    for (uint64_t i = 0; i < n;) {
        if (!keep[i]) { ++i; continue; }
        uint64_t j = i + 1;
        while (j < n && keep[j]) ++j;
        if (!runs.empty() && i - runs.back().second <= kBridge) runs.back().second = j;
        else runs.emplace_back(i, j);
        i = j;
    }
    for (const auto &r : runs) taken += r.second - r.first;
    if (n_taken) *n_taken = taken;
    if (!taken) return INQ_OK;
    cudaStream_t s = ctx->stream;
    const bool first = ctx->R == 0;
    CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_H2D0], s));
    if (runs.size() <= 1024) {
        // coordinate-sorted input and a contiguous catalog range: a handful of runs, copied straight from the caller's arrays
        for (const auto &r : runs)
            TRY(push_impl(ctx, r.second - r.first, contig + r.first, ref_start + r.first, ref_end + r.first, mapq + r.first, hp + r.first,
                          flags + r.first, cigar_off + r.first, cigar_words));
    } else {
        // scattered: gather into pinned staging chunks (all host threads), one push per chunk
        constexpr uint64_t kStageWords = 16u << 20, kStageReads = 1u << 20;
        const size_t meta_bytes = kStageReads * 12 + (kStageReads + 1) * 8 + kStageReads * 3;
        if (!ctx->h_route_stage) CU_TRY(ctx, cudaMallocHost(&ctx->h_route_stage, kStageWords * 4 + meta_bytes + 64));
        uint32_t *sw = static_cast<uint32_t *>(ctx->h_route_stage);
        uint8_t *m = reinterpret_cast<uint8_t *>(sw + kStageWords);
        uint64_t *so = reinterpret_cast<uint64_t *>(m); m += (kStageReads + 1) * 8;
        int32_t *sc = reinterpret_cast<int32_t *>(m); m += kStageReads * 4;
        int32_t *ss = reinterpret_cast<int32_t *>(m); m += kStageReads * 4;
        int32_t *se = reinterpret_cast<int32_t *>(m); m += kStageReads * 4;
        uint8_t *sq = m; m += kStageReads;
        uint8_t *sh = m; m += kStageReads;
        uint8_t *sf = m;
        std::vector<uint64_t> idx;
        idx.reserve(kStageReads);
        size_t ri = 0;
        uint64_t pos_in_run = 0;
        while (ri < runs.size()) {
            idx.clear();
            uint64_t words = 0;
            so[0] = 0;
            while (ri < runs.size() && idx.size() < kStageReads) {
                const uint64_t i = runs[ri].first + pos_in_run;
                const uint64_t nw = cigar_off[i + 1] - cigar_off[i];
                if (nw > kStageWords) return fail(ctx, INQ_ERR_TOO_LARGE, "inq_push_reads_routed: a read has more CIGAR words than the staging buffer");
                if (words + nw > kStageWords) break;
                so[idx.size() + 1] = words + nw;
                idx.push_back(i);
                words += nw;
                if (++pos_in_run == runs[ri].second - runs[ri].first) { ++ri; pos_in_run = 0; }
            }
            const uint64_t cnt = idx.size();
            parallel(cnt, [&](uint64_t a, uint64_t b) {
                for (uint64_t k = a; k < b; ++k) {
                    const uint64_t i = idx[k];
                    sc[k] = contig[i]; ss[k] = ref_start[i]; se[k] = ref_end[i];
                    sq[k] = mapq[i]; sh[k] = hp[i]; sf[k] = flags[i];
                    memcpy(sw + so[k], cigar_words + cigar_off[i], (size_t)(cigar_off[i + 1] - cigar_off[i]) * 4);
                }
            });
            TRY(push_impl(ctx, cnt, sc, ss, se, sq, sh, sf, so, sw));
            CU_TRY(ctx, cudaStreamSynchronize(s));               // the staging buffer is refilled next
        }
    }
    CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_H2D1], s));
    CU_TRY(ctx, cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[EV_H2D0], ctx->ev[EV_H2D1]);
    ctx->ms_h2d = (first ? 0.f : ctx->ms_h2d) + ms;
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
    unphased = unphased != 0;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t ntiles = (uint32_t)((C + kTileWords - 1) / kTileWords);
    if ((C + kTileWords - 1) / kTileWords > 0x7FFFFFFFull) return fail(ctx, INQ_ERR_TOO_LARGE, "too many CIGAR words");
    const uint64_t n_wt = (uint64_t)ntiles * (kTileWords / kWarpTileWords);
    const bool work = R && L;
    TRY(ensure(ctx, ctx->cand_lo, R));
    TRY(ensure(ctx, ctx->cand_n, R));
    TRY(ensure(ctx, ctx->rd_pre, R + 1));
    TRY(ensure(ctx, ctx->wt, n_wt + 2));
    TRY(ensure(ctx, ctx->wtot, n_wt + 2));
    TRY(ensure(ctx, ctx->wt_sbase, n_wt + 1));
    TRY(ensure(ctx, ctx->desc_wt, 2 * ((n_wt + kXsTile - 1) / kXsTile + 2 * kMaxRanges + 1)));       // even: k_zero clears it in 16-byte units
    if (ntiles) TRY(make_tensor_map(ctx, (uint64_t)ntiles * kTileWords));
    const uint64_t raw_slack = (uint64_t)ctx->sm_count * ctx->scan_ctas_per_sm * kScanWarps * kEvChunk;
    // event storage is sized speculatively (1/16 of the words; checked and regrown after the run); it hands out
    // kEvChunk-slot chunks and a warp strands the remainder of its chunk whenever a tile does not fit any more
    if (ctx->evraw.cap == 0) TRY(ensure(ctx, ctx->evraw, C / 16 + 4096 + raw_slack));
    TRY(build_plan(ctx, n_wt));
    if (getenv("INQ_JOIN_TRACE") && R) {
        const uint64_t ctas = (R + 255) / 256;
        if (ctas != ctx->join_trace_ctas) {
            if (ctx->join_trace) cudaFree(ctx->join_trace);
            ctx->join_trace = nullptr;
            CU_TRY(ctx, cudaMalloc(&ctx->join_trace, (ctas * 3 + 8) * sizeof(unsigned long long)));
            ctx->join_trace_ctas = ctas;
            drop_graph(ctx);
            ctx->have_last_key = false;
        }
    }

    // results go straight to the caller's arrays when those are pinned, otherwise through a pinned staging buffer
    const bool direct_out = L == 0 || (is_pinned(twice_h1) && is_pinned(twice_h2) && is_pinned(valid_mask));
    RunParams rp{minlen, support, unphased, twice_h1, twice_h2, valid_mask, ctx->opt_timing};
    if (!direct_out) {
        const size_t need = (size_t)L * 17 + 64;
        if (need > ctx->h_stage_bytes) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr;
            ctx->h_stage_bytes = 0;
            CU_TRY(ctx, cudaMallocHost(&ctx->h_stage, need));
            ctx->h_stage_bytes = need;
        }
        rp.o1 = static_cast<int64_t *>(ctx->h_stage);
        rp.o2 = rp.o1 + L;
        rp.ov = reinterpret_cast<uint8_t *>(rp.o2 + L);
    }

    if (ctx->opt_push_kernel && L) {
        void *a = nullptr, *b = nullptr, *c = nullptr;
        if (cudaHostGetDevicePointer(&a, rp.o1, 0) == cudaSuccess && cudaHostGetDevicePointer(&b, rp.o2, 0) == cudaSuccess &&
            cudaHostGetDevicePointer(&c, rp.ov, 0) == cudaSuccess) {
            rp.d1 = static_cast<int64_t *>(a);
            rp.d2 = static_cast<int64_t *>(b);
            rp.dv = static_cast<uint8_t *>(c);
        } else {
            cudaGetLastError();                               // pinned but not mapped: the copy engines take it
        }
    }

    // the call buffer holds one slot per candidate; on the first call of a context its size is read back from a
    // join-only pre-pass, afterwards it is sized from the previous run and checked after the pass
    if (work && ctx->vals.cap == 0) {
        const uint32_t loc_scan_tiles = (uint32_t)(((uint64_t)L + 1 + kXsTile - 1) / kXsTile);
        ReadView rv{ctx->contig.p, ctx->rs.p, ctx->re.p, ctx->mapq.p, ctx->hp.p, ctx->flags.p, ctx->cig_off.p, R};
        LocusView lv{ctx->contig_off.p, ctx->lstart.p, ctx->lend.p, ctx->lpmax.p, ctx->n_contigs};
        CU_TRY(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DevCounters), s));
        CU_TRY(ctx, cudaMemsetAsync(ctx->delta.p, 0, ((uint64_t)L + 2) * sizeof(uint32_t), s));
        CU_TRY(ctx, cudaMemsetAsync(ctx->desc_scan.p, 0, 2 * ((uint64_t)loc_scan_tiles + 1) * sizeof(uint64_t), s));
        k_join_ranges<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(rv, lv, unphased, ctx->opt_join_coop, ctx->cand_lo.p, ctx->cand_n.p, ctx->delta.p, ctx->d_ctr, ctx->join_trace);
        const unsigned g = std::min<unsigned>(loc_scan_tiles, (unsigned)ctx->sm_count * 4);
        k_exclusive_scan<<<g, kXsThreads, 0, s>>>(ctx->delta.p, ctx->lcnt.p, (uint64_t)L + 1, loc_scan_tiles, ctx->desc_scan.p,
                                                  &ctx->d_ctr->scan_counter[2], nullptr);
        k_exclusive_scan<<<g, kXsThreads, 0, s>>>(ctx->lcnt.p + 1, ctx->seg_off.p, (uint64_t)L, loc_scan_tiles,
                                                  ctx->desc_scan.p + loc_scan_tiles + 1, &ctx->d_ctr->scan_counter[3], &ctx->d_ctr->flags);
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaMemcpyAsync(ctx->h_total, ctx->seg_off.p + L, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU_TRY(ctx, cudaStreamSynchronize(s));
        TRY(ensure(ctx, ctx->vals, (uint64_t)*ctx->h_total + 1));
    }

    uint32_t launches = 0;
    bool used_graph = false, done = false;
    for (int attempt = 0; attempt < 4 && !done; ++attempt) {
        const GraphKey key{ctx->data_gen, ctx->buf_gen, minlen, support, unphased, rp.timing, rp.o1, rp.o2, rp.ov};
        used_graph = false;
        if (ctx->graph && !(ctx->graph_key == key)) drop_graph(ctx);
        if (ctx->opt_graph && !ctx->graph && ctx->have_last_key && ctx->last_key == key) {
            // second call with the same shape: capture the DAG once, replay it from now on
            cudaGraph_t g = nullptr;
            CU_TRY(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int rc = enqueue_pass(ctx, rp, true, &ctx->graph_launches);
            cudaError_t e = cudaStreamEndCapture(s, &g);
            if (rc != INQ_OK) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) return fail(ctx, INQ_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&ctx->graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) { ctx->graph = nullptr; return fail(ctx, INQ_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
            ctx->graph_key = key;
        }
        if (rp.timing >= 1) CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_START], s));
        if (ctx->graph) {
            CU_TRY(ctx, cudaGraphLaunch(ctx->graph, s));
            launches = ctx->graph_launches;
            used_graph = true;
        } else {
            TRY(enqueue_pass(ctx, rp, false, &launches));
        }
        if (rp.timing >= 1) CU_TRY(ctx, cudaEventRecord(ctx->ev[EV_END], s));
        ctx->last_key = key;
        ctx->have_last_key = true;
        {
            // a blocking synchronize wakes the host tens of microseconds late; a pass takes 0.1 - 3 ms, so poll first
            cudaError_t q = cudaErrorNotReady;
            for (int spin = 0; spin < 200000 && (q = cudaStreamQuery(s)) == cudaErrorNotReady; ++spin) {}
            if (q != cudaSuccess && q != cudaErrorNotReady) CU_TRY(ctx, q);
            CU_TRY(ctx, cudaStreamSynchronize(s));
        }

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
        if (retry) { ctx->have_last_key = false; continue; }       // buffers moved: the next attempt launches directly
        if (ctx->plan.big_known) {
            // a chunk whose CTA-path kernel was left out turned out to have deep loci (other parameters than the pass the
            // guess came from): run again with it
            bool again = false;
            for (int c = 0; c < ctx->plan.n_chunks; ++c)
                if (ctx->h_ctr->big_count[c] && !ctx->plan.need_big[c]) { ctx->plan.need_big[c] = true; again = true; }
            if (again) { drop_graph(ctx); ctx->have_last_key = false; continue; }
        }
        if (f & kFlagCountOverflow) return fail(ctx, INQ_ERR_TOO_LARGE, "pair or event count exceeds 2^32");
        if (f & kFlagValsOverflow) return fail(ctx, INQ_ERR_STATE, "internal: call buffer overflow");
        if (f & kFlagBadSa)
            return fail(ctx, INQ_ERR_BAD_SA, "read %llu passes the filter and has a soft clip, but its SA tag is not a string or cannot be split/parsed "
                        "(the reference panics in is_accidental_2d, call.rs:431,439-450)", (unsigned long long)ctx->h_ctr->bad_sa_read);
        if (f & kFlagBadHp)
            return fail(ctx, INQ_ERR_BAD_HP, "read %llu passes the phased filter but carries HP %u, outside {0,1,2} (the reference panics, call.rs:358)",
                        (unsigned long long)ctx->h_ctr->bad_hp_read, ctx->h_ctr->bad_hp_value);
        if (f & kFlagMedianEmpty)
            return fail(ctx, INQ_ERR_MEDIAN_EMPTY, "support == 0 with a bucket without usable calls (the reference panics, call.rs:516)");
        if (!ctx->plan.big_known) {
            for (int c = 0; c < ctx->plan.n_chunks; ++c) ctx->plan.need_big[c] = ctx->h_ctr->big_count[c] != 0;
            ctx->plan.big_known = true;
        }
        done = true;
    }
    if (!done) return fail(ctx, INQ_ERR_STATE, "internal: speculative buffers still too small after 4 attempts");
    ctx->last_n_events = (ntiles && L) ? ctx->h_total[3] : 0;      // .y of wt[n_wt]
    if (ctx->join_trace) {
        // experiments: dump the trace of the last pass (binary u64 triples: start ns, end ns, SM), plus the pass start
        if (const char *path = getenv("INQ_JOIN_TRACE")) {
            std::vector<unsigned long long> h(ctx->join_trace_ctas * 3);
            CU_TRY(ctx, cudaMemcpy(h.data(), ctx->join_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            if (FILE *f = fopen(path, "wb")) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
        }
    }
    if (!direct_out && L) {
        memcpy(twice_h1, rp.o1, (size_t)L * sizeof(int64_t));
        memcpy(twice_h2, rp.o2, (size_t)L * sizeof(int64_t));
        memcpy(valid_mask, rp.ov, (size_t)L);
    }

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
        stats->n_ranges = (uint32_t)ctx->plan.K;
        stats->used_graph = used_graph ? 1u : 0u;
        stats->reads_sorted = ctx->plan.reads_sorted ? 1u : 0u;
        stats->n_median_chunks = (uint32_t)ctx->plan.n_chunks;
        if (rp.timing) {
            auto el = [&](int a, int b) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; } return ms; };
            stats->ms_total = el(EV_START, EV_END);
            if (rp.timing >= 2) {
                stats->ms_join = el(EV_START, EV_JOIN);          // memsets + join + segment offsets, under the CIGAR scan
                stats->ms_index = el(EV_START, EV_INDEX);        // ... of which the zeroing kernel
                if (getenv("INQ_DEBUG_TIMING"))
                    fprintf(stderr, "[inq timing] zero done %.3f  join done %.3f  offsets done %.3f  scan done %.3f ms after the start\n",
                            el(EV_START, EV_INDEX), work ? el(EV_START, EV_JOIN0) : 0.f, el(EV_START, EV_JOIN), el(EV_START, EV_SCAN0 + 2 * (ctx->plan.K - 1) + 1));
                for (int k = 0; k < ctx->plan.K; ++k) {
                    stats->ms_cigar += el(EV_SCAN0 + 2 * k, EV_SCAN0 + 2 * k + 1);
                    stats->ms_pairs += el(EV_PAIR0 + 2 * k, EV_PAIR0 + 2 * k + 1);
                }
                // prefix scan over the last range's warp-tile totals (incl. waiting for the join stream)
                stats->ms_fixup = el(EV_SCAN0 + 2 * (ctx->plan.K - 1) + 1, EV_PAIR0 + 2 * (ctx->plan.K - 1));
                stats->ms_median = el(EV_MED0, EV_MED1);
                stats->ms_d2h = el(EV_MED1, EV_D2H);
            } else {
                // S0 runs the K scans back to back right after the pass starts: start of the pass -> end of the last scan
                // (a few us of fork on top of the kernels: the conservative side for a bandwidth figure)
                stats->ms_cigar = el(EV_START, EV_SCAN0 + 2 * (ctx->plan.K - 1) + 1);
            }
            stats->ms_scan = el(EV_SCAN0 + 2 * (ctx->plan.K - 1) + 1, EV_END);      // what is left exposed after the last range is scanned
        }
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
