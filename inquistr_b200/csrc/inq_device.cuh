// inq_device.cuh -- sm_100a kernels of the `inquiSTR call` hot path.
//
// Reference semantics being reproduced (files under /root/reference/src, v0.13.0):
//   K1  k_join_count      read x locus overlap join + filter   call.rs:288,297-301,338,349-355
//   K2  k_cigar_scan      segmented CIGAR scan -> event list   call.rs:377-413 (position cursor + op tests)
//   K2b k_pair_eval       per pair window sum + bucket scatter call.rs:388-403 (window test), 304,358
//   K3  k_locus_median*   sort / split / support / median      call.rs:308-321,365-369,497-522
//
// Layout in HBM (all structure-of-arrays):
//   reads : contig/ref_start/ref_end int32[R], mapq/hp/flags u8[R], cig_off u64[R+1], cigar u32[C]
//   loci  : start/end/pmax_end int32[L] sorted by (contig,start), contig_off int64[n_contigs+1]
//   events: uint2{pos1 (1-based anchor, u32), val = (signed len << 1) | is_softclip}[E], ev_off u32[R+1]
//   buckets: 2 per locus (H1,H2 | unphased: all,unused); cnt/off u32[2L+1]; vals u64[P]
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace inq {

// ----------------------------------------------------------------------------------------------
// constants
constexpr int kWarp = 32;
constexpr int kTileWords = 8192;            // CIGAR words per tile (32 KB)
constexpr int kScanThreads = 512;           // 16 compute warps, 512 words per warp, 16 per lane
constexpr int kScanStages = 3;              // smem ring of bulk-copied tiles
constexpr int kQuadsPerTile = kTileWords / 4;
constexpr int kWarpsPerScanCta = kScanThreads / kWarp;
constexpr int kQuadsPerWarp = kQuadsPerTile / kWarpsPerScanCta;   // 128
constexpr int kSlabs = kQuadsPerWarp / kWarp;                     // 4 x (32 lanes x uint4)

constexpr uint32_t kFlagBadHp = 1u << 0;
constexpr uint32_t kFlagMedianEmpty = 1u << 1;
constexpr uint32_t kFlagEventOverflow = 1u << 2;
constexpr uint32_t kFlagValsOverflow = 1u << 3;
constexpr uint32_t kFlagCountOverflow = 1u << 4;

constexpr uint64_t kDescInvalid = 0ull;
constexpr uint64_t kDescAggregate = 1ull << 62;
constexpr uint64_t kDescPrefix = 2ull << 62;
constexpr uint64_t kDescValueMask = (1ull << 62) - 1;

constexpr int64_t kCallBias = 1ll << 61;    // keys are (call + bias) << 1 | clip, 63 bits
constexpr uint64_t kKeyHapBit = 1ull << 63;
constexpr uint64_t kKeyInf = ~0ull;

// device-side counters, one struct per ctx
struct DevCounters {
    unsigned long long n_candidates;
    unsigned long long n_pairs;
    unsigned long long n_reads_joined;
    unsigned long long n_words_joined;
    unsigned long long op_visits;
    unsigned long long n_events;
    unsigned int flags;
    unsigned int tile_counter;
    unsigned int scan_counter;
    unsigned int big_count;
    unsigned int big_cursor;
    unsigned int pad;
};

// ----------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// same, but lets the hardware park the warp for up to ~`ns` per probe instead of spinning on issue slots
__device__ __forceinline__ void mbar_wait_parked(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// CIGAR word decode. BAM ops MIDNSHP=X -> 0..8; M,D,N,=,X consume the reference
// (call.rs:384-392,404): bit mask 0b1_1000_1101.
__device__ __forceinline__ uint32_t cig_consume(uint32_t w)
{
    return ((0x18Du >> (w & 15u)) & 1u) ? (w >> 4) : 0u;
}
// op is I, D or S and longer than minlen (call.rs:388,394,400 -- strict `>`)
__device__ __forceinline__ bool cig_is_event(uint32_t w, uint32_t minlen)
{
    return (((0x16u >> (w & 15u)) & 1u) != 0u) && ((w >> 4) > minlen);
}

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, d);
        if ((int)lane_id() >= d) v += n;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// first index i in [lo,hi) with a[i] >= key  (a sorted ascending)
__device__ __forceinline__ int lower_bound_i32(const int32_t *__restrict__ a, int lo, int hi, int64_t key)
{
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if ((int64_t)__ldg(a + mid) >= key) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ----------------------------------------------------------------------------------------------
// locus prefix-max of end, one CTA per contig (used for the phased join's lower bound)
__global__ void k_locus_pmax(const int64_t *__restrict__ contig_off, const int32_t *__restrict__ end,
                             int32_t *__restrict__ pmax)
{
    __shared__ int32_t wmax[32];
    __shared__ int32_t carry_s;
    const int64_t a = contig_off[blockIdx.x], b = contig_off[blockIdx.x + 1];
    if (threadIdx.x == 0) carry_s = INT32_MIN;
    __syncthreads();
    for (int64_t base = a; base < b; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        int32_t v = i < b ? end[i] : INT32_MIN;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int32_t n = __shfl_up_sync(0xffffffffu, v, d);
            if ((int)lane_id() >= d) v = max(v, n);
        }
        if (lane_id() == 31) wmax[threadIdx.x >> 5] = v;
        __syncthreads();
        int32_t pre = carry_s;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pre = max(pre, wmax[w]);
        v = max(v, pre);
        if (i < b) pmax[i] = v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = v;
        __syncthreads();
    }
}

// validate the catalog: sorted by start within contig, start >= 10, end >= start
__global__ void k_locus_check(int n_contigs, const int64_t *__restrict__ contig_off,
                              const int32_t *__restrict__ start, const int32_t *__restrict__ end,
                              unsigned int *__restrict__ bad /* bit0 start<10, bit1 order, bit2 end<start */)
{
    int64_t L = contig_off[n_contigs];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned int f = 0;
        if (start[i] < 10) f |= 1u;
        if (end[i] < start[i]) f |= 4u;
        if (i > 0 && start[i - 1] > start[i]) {
            // allowed only across a contig boundary
            int lo = 0, hi = n_contigs;          // find contig of i: last c with off[c] <= i
            while (lo < hi) { int m = (lo + hi) >> 1; if (contig_off[m + 1] <= i) lo = m + 1; else hi = m; }
            if (contig_off[lo] != i) f |= 2u;
        }
        if (f) atomicOr(bad, f);
    }
}

__global__ void k_rebase_offsets(uint64_t *__restrict__ off, uint64_t n, uint64_t base)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        off[i] += base;
}

// ----------------------------------------------------------------------------------------------
// tile index: tile_first[t] = first read r with cig_off[r] >= t * kTileWords  (t < ntiles),
// tile_first[ntiles] = R. Reads whose CIGAR starts inside tile t are [tile_first[t], tile_first[t+1]).
__global__ void k_tile_index(const uint64_t *__restrict__ cig_off, uint64_t R, uint32_t ntiles,
                             uint32_t *__restrict__ tile_first)
{
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > R) return;
    if (r == R) tile_first[ntiles] = (uint32_t)R;
    uint64_t t_lo = (r == 0) ? 0 : cig_off[r - 1] / kTileWords + 1;
    uint64_t t_hi = cig_off[r] / kTileWords;
    if (r == R) t_hi = ntiles;                       // tiles that start past the last offset
    for (uint64_t t = t_lo; t <= t_hi && t < ntiles; ++t) tile_first[t] = (uint32_t)r;
}

// ----------------------------------------------------------------------------------------------
// K1: read x locus overlap join, counting pass. One thread per read, warp-uniform candidate loop,
// warp-aggregated atomics into the per-bucket counters.
//
// unphased (call.rs:297-300): keep iff ref_start <= start_ext && ref_end >= end_ext && mapq > 10
// phased   (call.rs:350-352): keep iff HP present && !(start_ext < ref_start && ref_end < end_ext)
//                             && mapq > 10, among reads fetch() yields (pos < end_ext && endpos > start_ext)
struct ReadView {
    const int32_t *contig, *rs, *re;
    const uint8_t *mapq, *hp, *flags;
    const uint64_t *cig_off;
    uint64_t R;
};
struct LocusView {
    const int64_t *contig_off;
    const int32_t *start, *end, *pmax;
    int n_contigs;
};

__device__ __forceinline__ bool pair_passes(bool unphased, int32_t rs, int32_t re, int32_t lstart, int32_t lend)
{
    // windows as the reference computes them (call.rs:285-286), compared after `as u32` casts
    const uint32_t start_ext = (uint32_t)lstart - 10u, end_ext = (uint32_t)lend + 10u;
    const uint32_t urs = (uint32_t)rs, ure = (uint32_t)re;
    if (unphased) return !(start_ext < urs || ure < end_ext);
    const bool fetched = ((int64_t)rs < (int64_t)end_ext) && ((int64_t)re > (int64_t)start_ext);
    return fetched && !(start_ext < urs && ure < end_ext);
}

__device__ __forceinline__ void candidate_range(bool unphased, const LocusView &lv, int c, int32_t rs, int32_t re,
                                                int &lo, int &hi)
{
    const int l0 = (int)lv.contig_off[c], l1 = (int)lv.contig_off[c + 1];
    if (unphased) {
        lo = lower_bound_i32(lv.start, l0, l1, (int64_t)rs + 10);          // start - 10 >= rs
        hi = lower_bound_i32(lv.start, lo, l1, (int64_t)re - 10 + 1);      // start + 10 <= re (necessary)
    } else {
        hi = lower_bound_i32(lv.start, l0, l1, (int64_t)re + 10);          // start - 10 < re
        lo = lower_bound_i32(lv.pmax, l0, hi, (int64_t)rs - 10 + 1);       // max(end) + 10 > rs
    }
    if (hi < lo) hi = lo;
}

__global__ void __launch_bounds__(256)
k_join_count(ReadView rv, LocusView lv, int unphased, uint32_t *__restrict__ cand_lo, uint32_t *__restrict__ cand_n,
             uint32_t *__restrict__ bucket_cnt, DevCounters *__restrict__ ctr)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const bool live = r < rv.R;
    int lo = 0, n = 0;
    int32_t rs = 0, re = 0;
    uint32_t h = 0;
    if (live) {
        const int c = rv.contig[r];
        const uint32_t mq = rv.mapq[r];
        h = rv.hp[r];
        if (c >= 0 && c < lv.n_contigs && mq > 10u && (unphased || h != 0xFFu)) {
            rs = rv.rs[r];
            re = rv.re[r];
            int hi;
            candidate_range(unphased != 0, lv, c, rs, re, lo, hi);
            n = hi - lo;
        }
        cand_lo[r] = (uint32_t)lo;
        cand_n[r] = (uint32_t)n;
    }
    uint32_t npass = 0, nvisit = 0;     // bucketed pairs; pairs the reference walks (incl. HP 0)
    bool bad_hp = false;
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    for (int j = 0; j < nmax; ++j) {
        uint32_t bucket = 0xFFFFFFFFu;
        if (j < n) {
            const int l = lo + j;
            if (pair_passes(unphased != 0, rs, re, __ldg(lv.start + l), __ldg(lv.end + l))) {
                ++nvisit;                                       // call.rs:357 runs before the bucket lookup
                if (unphased) bucket = 2u * (uint32_t)l;
                else if (h > 2u) bad_hp = true;                 // call.rs:358 unwrap on None
                else if (h != 0u) bucket = 2u * (uint32_t)l + (h - 1u);  // HP 0 lands in the ignored bucket
            }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
        if (bucket != 0xFFFFFFFFu) {
            ++npass;
            if ((int)lane_id() == __ffs(peers) - 1) atomicAdd(bucket_cnt + bucket, (uint32_t)__popc(peers));
        }
    }
    // statistics (one atomic per warp per counter)
    uint64_t words = 0, nw = 0;
    if (live && nvisit) nw = rv.cig_off[r + 1] - rv.cig_off[r];
    if (npass) words = nw;
    const uint32_t cand_w = warp_sum((uint32_t)n);
    const uint32_t pass_w = warp_sum(npass);
    const uint32_t join_w = warp_sum((uint32_t)(npass != 0));
    const uint64_t words_w = warp_sum(words);
    const uint64_t visits_w = warp_sum(nw * nvisit);
    const uint32_t bad_w = __any_sync(0xffffffffu, bad_hp);
    if (lane_id() == 0) {
        if (cand_w) atomicAdd(&ctr->n_candidates, (unsigned long long)cand_w);
        if (pass_w) atomicAdd(&ctr->n_pairs, (unsigned long long)pass_w);
        if (join_w) atomicAdd(&ctr->n_reads_joined, (unsigned long long)join_w);
        if (words_w) atomicAdd(&ctr->n_words_joined, (unsigned long long)words_w);
        if (visits_w) atomicAdd(&ctr->op_visits, (unsigned long long)visits_w);
        if (bad_w) atomicOr(&ctr->flags, kFlagBadHp);
    }
}

// ----------------------------------------------------------------------------------------------
// decoupled look-back over tile descriptors (status in the top 2 bits, 62-bit value).
// `stop_at_prefix`: walk back until a tile in kDescPrefix state is found, summing values on the way.
// Called by one full warp. Returns the exclusive carry for tile t.
__device__ __forceinline__ uint64_t lookback(const uint64_t *__restrict__ desc, int64_t t)
{
    uint64_t acc = 0;
    int64_t base = t - 1;
    while (true) {
        const int64_t idx = base - (int64_t)lane_id();
        uint64_t d = kDescPrefix;                       // tiles before 0: prefix 0
        if (idx >= 0) {
            do { d = ld_relaxed_u64(desc + idx); } while ((d >> 62) == 0);
        }
        const uint32_t is_prefix = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        const int first = is_prefix ? (__ffs(is_prefix) - 1) : 32;
        const uint64_t contrib = ((int)lane_id() <= first) ? (d & kDescValueMask) : 0ull;
        acc += warp_sum(contrib);
        if (first < 32) break;
        base -= 32;
    }
    return acc;
}

// ----------------------------------------------------------------------------------------------
// K2: segmented CIGAR scan. Persistent CTAs pull 16 KB tiles of the flat packed-CIGAR stream
// through a 3-stage shared-memory ring filled by TMA 1-D bulk copies, compute the running
// reference consumption (warp-shuffle scans, carried across tiles by decoupled look-back and
// reset at read boundaries), and compact every I/D/S op longer than minlen into an ordered event
// list {1-based anchor position, signed length, soft-clip bit}. Each CIGAR word is read once.
struct ScanParams {
    const uint64_t *cig_off;      // R+1
    const int32_t *rs;            // ref_start
    const uint4 *tile_meta;       // ntiles: {first read starting in tile, #reads starting, last start offset, carried pos1}
    uint64_t *desc_ev;            // ntiles, zeroed
    uint64_t *desc_pos;           // ntiles, zeroed
    uint2 *events;
    uint32_t *ev_off;             // R+1
    DevCounters *ctr;
    uint64_t R;
    uint64_t ev_cap;
    uint32_t ntiles;
    uint32_t minlen;
    uint32_t debug;               // timing experiments only (INQ_SCAN_DEBUG): results are wrong when != 0
};

// per-tile metadata, one thread per tile (k_tile_index ran before): everything the scan kernel would
// otherwise have to fetch through dependent global loads on its critical path.
//   x = rA: first read whose CIGAR starts inside the tile      y = number of such reads
//   z = tile-local word offset of the last such read's start   w = ref_start(rA-1) + 1 (carried-in read)
__global__ void k_tile_meta(const uint32_t *__restrict__ tile_first, const uint64_t *__restrict__ cig_off,
                            const int32_t *__restrict__ rs, uint32_t ntiles, uint4 *__restrict__ meta)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t rA = tile_first[t], rB = tile_first[t + 1];
    uint4 m;
    m.x = rA;
    m.y = rB - rA;
    m.z = (rB > rA) ? (uint32_t)min(cig_off[rB - 1] - (uint64_t)t * kTileWords, (uint64_t)kTileWords) : 0u;
    m.w = (rA > 0 ? (uint32_t)rs[rA - 1] : 0u) + 1u;
    meta[t] = m;
}

constexpr int kMaxStarts = 256;             // read starts per tile staged in shared memory
constexpr int kLaneWords = kTileWords / kScanThreads;   // 16 consecutive words per compute lane
constexpr int kCtaThreads = kScanThreads + 32;          // 8 compute warps + 1 control warp
constexpr uint32_t kOpLut = 0x18Du | (0x16u << 16);     // bit op: consumes reference; bit 16+op: I/D/S

struct TileTables {
    uint32_t lpref[kScanThreads];           // warp-local exclusive ref-consumption prefix of each lane block
    uint16_t lev[kScanThreads];             // warp-local exclusive event count of each lane block
    uint32_t wsum[kWarpsPerScanCta];        // per-warp totals (phase A)
    uint32_t wev[kWarpsPerScanCta];
    uint32_t wbase[kWarpsPerScanCta];       // exclusive per-warp bases (publish)
    uint32_t webase[kWarpsPerScanCta];
    uint32_t tot_cons, tot_ev;
};

// read starts of one tile, staged by the control warp
struct Staging {
    uint32_t pos1[kMaxStarts];              // ref_start + 1 - S(start word): add S(word) for the op's anchor
    uint16_t off[kMaxStarts];               // tile-local word index where the read's CIGAR starts
    uint16_t ev[kMaxStarts];                // events in the tile before that word
    uint16_t owner[kScanThreads];           // per lane block: (#starts before the block << 1) | block holds a start
    uint64_t ev_base;                       // events before the tile (look-back)
    uint32_t carry_pos1;                    // carried-in read: ref_start + 1 + bases consumed before the tile
    uint32_t pad;
};

struct ScanSmem {
    alignas(1024) uint32_t stage[kScanStages][kTileWords];   // 128B-swizzled by the TMA tensor map
    alignas(16) uint4 meta[kScanStages];
    TileTables tab[2];
    Staging stg[2];
    alignas(8) uint64_t full[kScanStages];  // TMA landed                     (tx bytes)
    uint64_t freeb[kScanStages];            // compute warps done with stage  (8 arrivals)
    uint64_t bar_a[2];                      // phase A of a tile done         (8 arrivals)
    uint64_t ready[2];                      // control data of a tile ready   (1 arrival)
    uint32_t vid;
};
constexpr size_t kScanSmemBytes = sizeof(ScanSmem) + 1024;   // slack to align the swizzled stages to 1 KB

// word index inside a tile -> word index in the 128B-swizzled stage buffer
// (16-byte chunk index bits [2:4] ^= 128-byte row index bits [5:7])
__device__ __forceinline__ uint32_t swz(uint32_t idx) { return idx ^ (((idx >> 5) & 7u) << 2); }

__device__ __forceinline__ void tma_load_tile(void *dst_smem, const CUtensorMap *tmap, uint32_t row, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t tile_S(const TileTables &tb, const uint32_t *stage, uint32_t b)
{
    // reference bases consumed by tile words [0,b)
    if (b >= (uint32_t)kTileWords) return tb.tot_cons;
    const uint32_t blk = b / kLaneWords;
    uint32_t s = tb.wbase[blk >> 5] + tb.lpref[blk];
    for (uint32_t i = blk * kLaneWords; i < b; ++i) s += cig_consume(stage[swz(i)]);
    return s;
}
__device__ __forceinline__ uint32_t tile_E(const TileTables &tb, const uint32_t *stage, uint32_t b, uint32_t minlen)
{
    // events among tile words [0,b)
    if (b >= (uint32_t)kTileWords) return tb.tot_ev;
    const uint32_t blk = b / kLaneWords;
    uint32_t s = tb.webase[blk >> 5] + tb.lev[blk];
    for (uint32_t i = blk * kLaneWords; i < b; ++i) s += cig_is_event(stage[swz(i)], minlen) ? 1u : 0u;
    return s;
}

// Fused look-back over the two descriptor arrays of the CIGAR scan (one L2 round trip per window
// of 32 tiles for both):
//   *carry_pos = reference bases the carried-in read consumed before tile t (sum back to the nearest
//                tile that holds a read start),   *ev_base = events before tile t.
__device__ __forceinline__ void lookback2(const uint64_t *__restrict__ desc_pos, const uint64_t *__restrict__ desc_ev,
                                          int64_t t, uint64_t *carry_pos, uint64_t *ev_base)
{
    uint64_t acc_p = 0, acc_e = 0;
    bool done_p = false, done_e = false;
    int64_t base = t - 1;
    while (true) {
        const int64_t idx = base - (int64_t)lane_id();
        uint64_t dp = kDescPrefix, de = kDescPrefix;    // tiles before 0: prefix 0
        if (idx >= 0) {
            do {
                if (!done_p) dp = ld_relaxed_u64(desc_pos + idx);
                if (!done_e) de = ld_relaxed_u64(desc_ev + idx);
            } while ((dp >> 62) == 0 || (de >> 62) == 0);
        }
        if (!done_p) {
            const uint32_t m = __ballot_sync(0xffffffffu, (dp >> 62) == 2);
            const int first = m ? (__ffs(m) - 1) : 32;
            acc_p += warp_sum(((int)lane_id() <= first) ? (dp & kDescValueMask) : 0ull);
            done_p = first < 32;
        }
        if (!done_e) {
            const uint32_t m = __ballot_sync(0xffffffffu, (de >> 62) == 2);
            const int first = m ? (__ffs(m) - 1) : 32;
            acc_e += warp_sum(((int)lane_id() <= first) ? (de & kDescValueMask) : 0ull);
            done_e = first < 32;
        }
        if (done_p && done_e) break;
        base -= 32;
    }
    *carry_pos = acc_p;
    *ev_base = acc_e;
}

// Warp-specialised persistent kernel, 288 threads:
//   warps 0..7 (compute): phase A = one pass over 16 consecutive words per lane + two warp scans;
//                         phase D = emit the lane's events. They never block on global memory.
//   warp 8 (control)    : TMA issue, publishing tile aggregates, decoupled look-back, staging of the
//                         read starts (ref_start, first-event index) and the ev_off[] writes.
// The roles meet only through shared-memory mbarriers (full / bar_a / ready / freeb).
__global__ void __launch_bounds__(kCtaThreads, 2)
k_cigar_scan(const __grid_constant__ CUtensorMap tmap, ScanParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // keep the pointer in the shared address space (LDS/STS, not generic loads)
    ScanSmem &sm = *reinterpret_cast<ScanSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t kTileBytes = kTileWords * 4;
    constexpr uint32_t kRowsPerTile = kTileWords / 32;          // 128-byte rows
    const bool is_control = warp == kWarpsPerScanCta;

    if (tid == 0) {
        for (int s = 0; s < kScanStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.freeb[s], kWarpsPerScanCta); }
        for (int b = 0; b < 2; ++b) { mbar_init(&sm.bar_a[b], kWarpsPerScanCta); mbar_init(&sm.ready[b], 1); }
        fence_mbar_init();
        // CTA id in scheduling order: look-back only ever waits on CTAs that are already running
        sm.vid = atomicAdd(&p.ctr->tile_counter, 1u);
    }
    __syncthreads();
    const uint64_t vid = sm.vid, stride = gridDim.x;
    // tiles vid, vid+G, vid+2G, ...: neighbouring tiles are processed by different CTAs at the same time
    auto tile_of = [&](uint32_t itx) -> uint64_t { return vid + (uint64_t)itx * stride; };

    if (!is_control) {
        // =============================== compute warps ===============================
        const uint32_t thr = (p.minlen << 4) | 15u;             // (w >> 4) > minlen  <=>  w > thr
        // this thread's 16 consecutive words: 128-byte row tid/2, chunks 4*(tid&1)+j, swizzled
        const uint32_t rowq = (tid >> 1) * 8, x0 = ((tid & 1u) << 2) ^ ((tid >> 1) & 7u);

        // evmask: bit i <-> word i of the lane block is an event; clast: bases consumed inside the
        // block before its LAST event
        auto phase_a = [&](uint32_t itx, uint32_t &evmask, uint32_t &clast) {
            const uint32_t s = itx % kScanStages;
            mbar_wait(&sm.full[s], (itx / kScanStages) & 1u);
            const uint4 *st4 = reinterpret_cast<const uint4 *>(sm.stage[s]);
            TileTables &tb = sm.tab[itx & 1u];
            uint32_t c = 0;
            evmask = 0;
            clast = 0;
            if (!(p.debug & 8u))
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 v = st4[rowq + (x0 ^ (uint32_t)j)];
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t w = w4[k], lut = kOpLut >> (w & 15u);
                    const bool ev = ((lut & 0x10000u) != 0u) & (w > thr);
                    evmask = ev ? (evmask | (1u << (j * 4 + k))) : evmask;
                    clast = ev ? c : clast;
                    c += (lut & 1u) ? (w >> 4) : 0u;
                }
            }
            const uint32_t ne = __popc(evmask);
            const uint32_t incl_c = warp_incl_scan(c), incl_e = warp_incl_scan(ne);
            tb.lpref[tid] = incl_c - c;
            tb.lev[tid] = (uint16_t)(incl_e - ne);
            if (lane == 31) { tb.wsum[warp] = incl_c; tb.wev[warp] = incl_e; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.bar_a[itx & 1u]);
        };

        uint32_t evmask_n = 0, clast_n = 0;
        if (tile_of(0) < p.ntiles) phase_a(0, evmask_n, clast_n);
        for (uint32_t it = 0;; ++it) {
            if (tile_of(it) >= p.ntiles) break;
            const uint32_t s = it % kScanStages;
            uint32_t evmask = evmask_n;
            const uint32_t clast = clast_n;
            if (tile_of(it + 1) < p.ntiles) phase_a(it + 1, evmask_n, clast_n);

            // ---- phase D: emit this lane's events (about 1% of the words), last event of the block first
            {
                // always taken (normally already complete): it also keeps this warp from overwriting
                // tab[it&1] in its next phase A while the control warp still reads it
                mbar_wait_parked(&sm.ready[it & 1u], (it >> 1) & 1u, 2000u);
                if (evmask && !(p.debug & 1u)) {
                    const uint32_t *stage = sm.stage[s];
                    const TileTables &tb = sm.tab[it & 1u];
                    const Staging &sg = sm.stg[it & 1u];
                    const uint4 meta = sm.meta[s];
                    const uint32_t rA = meta.x, nrs = meta.y, nst = min(nrs, (uint32_t)kMaxStarts);
                    const uint64_t g0 = (tile_of(it)) * kTileWords;
                    const uint32_t s_blk = tb.wbase[warp] + tb.lpref[tid], e_blk = tb.webase[warp] + tb.lev[tid];
                    const uint32_t own = sg.owner[tid];
                    const uint64_t ev_base = sg.ev_base;
                    bool first = true;
                    do {
                        const uint32_t bit = 31u - (uint32_t)__clz(evmask);
                        evmask ^= 1u << bit;
                        const uint32_t idx = tid * kLaneWords + bit;
                        const uint32_t w = stage[swz(idx)];
                        uint32_t s_in = clast;                           // captured in phase A for the last event
                        if (!first) {
                            s_in = 0;
                            for (uint32_t i = idx - bit; i < idx; ++i) s_in += cig_consume(stage[swz(i)]);
                        }
                        first = false;
                        const uint32_t s_here = s_blk + s_in, e_here = e_blk + __popc(evmask);  // lower bits remain
                        // owning read = last read start at or before this word
                        uint32_t lo = own >> 1;
                        if (own & 1u) {                                  // a read starts inside this block
                            uint32_t hi = nst;
                            lo = 0;
                            while (lo < hi) {
                                const uint32_t mid = (lo + hi) >> 1;
                                if (sg.off[mid] <= idx) lo = mid + 1; else hi = mid;
                            }
                        }
                        uint32_t pos1;
                        if (lo == 0) {
                            pos1 = sg.carry_pos1 + s_here;               // read carried in from an earlier tile
                        } else if (lo == nst && nrs > nst) {
                            // more read starts than the staging area holds: search the tail in global memory
                            const uint64_t g = g0 + idx;
                            uint32_t a = rA + nst - 1, b = rA + nrs;
                            while (a < b) {
                                const uint32_t mid = a + ((b - a) >> 1);
                                if (p.cig_off[mid] <= g) a = mid + 1; else b = mid;
                            }
                            const uint32_t r = a - 1;
                            pos1 = (uint32_t)p.rs[r] + 1u + s_here - tile_S(tb, stage, (uint32_t)(p.cig_off[r] - g0));
                        } else {
                            pos1 = sg.pos1[lo - 1] + s_here;             // call.rs:380 cursor at this op
                        }
                        const uint32_t len = w >> 4, op = w & 15u;
                        const int32_t val = (int32_t)(((op == 2u) ? (0u - len) : len) << 1) | (int32_t)(op == 4u);
                        const uint64_t slot = ev_base + e_here;
                        if (slot < p.ev_cap) p.events[slot] = make_uint2(pos1, (uint32_t)val);
                        else atomicOr(&p.ctr->flags, kFlagEventOverflow);
                    } while (evmask);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.freeb[s]);        // this warp is done with stage s and tab/stg[it&1]
        }
    } else {
        // =============================== control warp ===============================
        auto issue = [&](uint32_t s, uint64_t t) {
            mbar_expect_tx(&sm.full[s], kTileBytes + 16u);
            tma_load_tile(sm.stage[s], &tmap, (uint32_t)t * kRowsPerTile, &sm.full[s]);
            bulk_copy_g2s(&sm.meta[s], p.tile_meta + t, 16u, &sm.full[s]);
        };
        if (lane == 0) {
            for (int s = 0; s < kScanStages; ++s)
                if (tile_of(s) < p.ntiles) issue(s, tile_of(s));
        }

        // after phase A of slot itx: per-warp bases + publish the tile's aggregates. A tile that holds a
        // read start resets the position carry, so its position descriptor is final at once.
        auto publish = [&](uint32_t itx) {
            const uint32_t t = (uint32_t)tile_of(itx);
            TileTables &tb = sm.tab[itx & 1u];
            mbar_wait(&sm.full[itx % kScanStages], (itx / kScanStages) & 1u);   // already complete: acquire the TMA data
            const uint32_t a = lane < kWarpsPerScanCta ? tb.wsum[lane] : 0u, b = lane < kWarpsPerScanCta ? tb.wev[lane] : 0u;
            const uint32_t ia = warp_incl_scan(a), ib = warp_incl_scan(b);
            if (lane < kWarpsPerScanCta) { tb.wbase[lane] = ia - a; tb.webase[lane] = ib - b; }
            const uint32_t tot_cons = __shfl_sync(0xffffffffu, ia, 31), tot_ev = __shfl_sync(0xffffffffu, ib, 31);
            if (lane == 0) { tb.tot_cons = tot_cons; tb.tot_ev = tot_ev; }
            __syncwarp();
            if (lane == 0) {
                const uint4 m = sm.meta[itx % kScanStages];
                const uint32_t trailing = m.y ? tot_cons - tile_S(tb, sm.stage[itx % kScanStages], m.z) : tot_cons;
                st_relaxed_u64(p.desc_pos + t, (m.y ? kDescPrefix : kDescAggregate) | (uint64_t)trailing);
                st_relaxed_u64(p.desc_ev + t, (t == 0 ? kDescPrefix : kDescAggregate) | (uint64_t)tot_ev);
            }
        };

        // stage the read starts of slot itx (needs publish(itx)): offsets, event ranks, position bases, and
        // the per-lane-block owner table that replaces a binary search for most events
        // (pre_off, pre_rs): cig_off / ref_start of read start `lane` of the slot, loaded early by the caller
        auto staging = [&](uint32_t itx, uint64_t pre_off, int32_t pre_rs) {
            const uint32_t s = itx % kScanStages;
            const TileTables &tb = sm.tab[itx & 1u];
            Staging &sg = sm.stg[itx & 1u];
            const uint32_t *stage = sm.stage[s];
            const uint4 meta = sm.meta[s];
            const uint32_t rA = meta.x, nst = min(meta.y, (uint32_t)kMaxStarts);
            const uint64_t g0 = tile_of(itx) * kTileWords;
            uint32_t *own32 = reinterpret_cast<uint32_t *>(sg.owner);
#pragma unroll
            for (int k = 0; k < kScanThreads / 64; ++k) own32[lane + 32 * k] = 0u;
            __syncwarp();
            for (uint32_t i = lane; i < nst; i += 32) {
                const uint32_t r = rA + i;
                if (i >= 32) { pre_off = p.cig_off[r]; pre_rs = p.rs[r]; }
                const uint32_t b = (uint32_t)min(pre_off - g0, (uint64_t)kTileWords);
                sg.off[i] = (uint16_t)b;
                sg.ev[i] = (uint16_t)tile_E(tb, stage, b, p.minlen);
                sg.pos1[i] = (uint32_t)pre_rs + 1u - tile_S(tb, stage, b);
                if (b < (uint32_t)kTileWords) {
                    const uint32_t blk = b / kLaneWords;
                    atomicAdd(&own32[blk >> 1], (blk & 1u) ? 0x10000u : 1u);      // u16 histogram, counts <= kMaxStarts
                }
            }
            __syncwarp();
            // exclusive scan of the histogram: lane owns kScanThreads/32 consecutive blocks
            constexpr int kPer = kScanThreads / 32;
            uint32_t cnt[kPer], sum = 0;
#pragma unroll
            for (int k = 0; k < kPer; ++k) { cnt[k] = sg.owner[lane * kPer + k]; sum += cnt[k]; }
            uint32_t run = warp_incl_scan(sum) - sum;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                sg.owner[lane * kPer + k] = (uint16_t)((run << 1) | (cnt[k] ? 1u : 0u));
                run += cnt[k];
            }
        };
        // early loads for the staging of slot itx (its tile and metadata must have landed)
        auto prefetch_starts = [&](uint32_t itx, uint64_t &pre_off, int32_t &pre_rs) {
            const uint32_t s = itx % kScanStages;
            mbar_wait_parked(&sm.full[s], (itx / kScanStages) & 1u, 1000u);
            const uint4 meta = sm.meta[s];
            pre_off = 0;
            pre_rs = 0;
            if (lane < meta.y) { pre_off = p.cig_off[meta.x + lane]; pre_rs = p.rs[meta.x + lane]; }
        };

        if (tile_of(0) < p.ntiles) {
            uint64_t po; int32_t pr;
            prefetch_starts(0, po, pr);
            mbar_wait_parked(&sm.bar_a[0], 0, 1000u);
            publish(0);
            staging(0, po, pr);
        }
        for (uint32_t it = 0;; ++it) {
            const uint64_t t64 = tile_of(it);
            if (t64 >= p.ntiles) break;
            const uint32_t t = (uint32_t)t64;
            const uint32_t s = it % kScanStages;
            const TileTables &tb = sm.tab[it & 1u];
            Staging &sg = sm.stg[it & 1u];
            const uint4 meta = sm.meta[s];
            const uint32_t rA = meta.x, nrs = meta.y, nst = min(nrs, (uint32_t)kMaxStarts);
            const uint32_t tot_cons = tb.tot_cons, tot_ev = tb.tot_ev;
            const bool have_next = tile_of(it + 1) < p.ntiles;
            uint64_t pre_off = 0;
            int32_t pre_rs = 0;
            if (have_next && !(p.debug & 4u)) prefetch_starts(it + 1, pre_off, pre_rs);   // latency overlaps the look-back below

            // look back (predecessors published one iteration ago), finish this tile's control data
            uint64_t ev_base = 0, carry_pos = 0;
            if (t > 0 && !(p.debug & 2u)) {
                lookback2(p.desc_pos, p.desc_ev, (int64_t)t, &carry_pos, &ev_base);
                if (lane == 0) {
                    st_relaxed_u64(p.desc_ev + t, kDescPrefix | ((ev_base + tot_ev) & kDescValueMask));
                    if (nrs == 0)
                        st_relaxed_u64(p.desc_pos + t, kDescPrefix | ((carry_pos + tot_cons) & 0xFFFFFFFFull));
                }
            }
            if (lane == 0) {
                sg.carry_pos1 = meta.w + (uint32_t)carry_pos;
                sg.ev_base = ev_base;
            }
            // first-event index of every read whose CIGAR starts in this tile
            if (!(p.debug & 16u))
            for (uint32_t i = lane; i < nrs; i += 32) {
                uint32_t e;
                if (i < nst) e = sg.ev[i];
                else e = tile_E(tb, sm.stage[s], (uint32_t)min(p.cig_off[rA + i] - (uint64_t)t * kTileWords, (uint64_t)kTileWords), p.minlen);
                p.ev_off[rA + i] = (uint32_t)(ev_base + e);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.ready[it & 1u]);
                if (t == p.ntiles - 1) {
                    p.ev_off[p.R] = (uint32_t)(ev_base + tot_ev);
                    p.ctr->n_events = ev_base + tot_ev;
                    if (ev_base + tot_ev > 0xFFFFFFFFull) atomicOr(&p.ctr->flags, kFlagCountOverflow);
                }
            }

            // next tile: aggregates out as early as possible, then its staging
            if (have_next) {
                mbar_wait_parked(&sm.bar_a[(it + 1) & 1u], ((it + 1) >> 1) & 1u, 1000u);
                if (!(p.debug & 16u)) publish(it + 1);
                if (!(p.debug & 4u)) staging(it + 1, pre_off, pre_rs);
            }

            // refill stage s once the compute warps are done with it
            mbar_wait_parked(&sm.freeb[s], (it / kScanStages) & 1u, 1000u);
            if (lane == 0) {
                const uint64_t t2 = tile_of(it + kScanStages);
                if (t2 < p.ntiles) {
                    fence_proxy_async();
                    issue(s, t2);
                }
            }
            __syncwarp();
        }
    }
}

// ----------------------------------------------------------------------------------------------
// exclusive scan of u32 counts -> u32 offsets (n+1 outputs), single pass, decoupled look-back
constexpr int kXsThreads = 256;
constexpr int kXsItems = 8;
constexpr int kXsTile = kXsThreads * kXsItems;

__global__ void __launch_bounds__(kXsThreads)
k_exclusive_scan(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n, uint32_t ntiles,
                 uint64_t *__restrict__ desc, DevCounters *__restrict__ ctr)
{
    __shared__ uint32_t wsum[kXsThreads / 32];
    __shared__ uint32_t tile_s;
    __shared__ uint64_t base_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) tile_s = atomicAdd(&ctr->scan_counter, 1u);
        __syncthreads();
        const uint32_t t = tile_s;
        if (t >= ntiles) break;
        const uint64_t i0 = (uint64_t)t * kXsTile + (uint64_t)tid * kXsItems;
        uint32_t v[kXsItems], sum = 0;
#pragma unroll
        for (int k = 0; k < kXsItems; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
            sum += v[k];
        }
        const uint32_t incl = warp_incl_scan(sum);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kXsThreads / 32; ++w) {
            if (w < (int)warp) wbase += wsum[w];
            total += wsum[w];
        }
        if (warp == 0) {
            if (lane == 0) st_relaxed_u64(desc + t, (t == 0 ? kDescPrefix : kDescAggregate) | (uint64_t)total);
            uint64_t base = 0;
            if (t > 0) {
                base = lookback(desc, (int64_t)t);
                if (lane == 0) st_relaxed_u64(desc + t, kDescPrefix | ((base + total) & kDescValueMask));
            }
            if (lane == 0) {
                base_s = base;
                if (t == ntiles - 1) {
                    out[n] = (uint32_t)(base + total);
                    if (base + total > 0xFFFFFFFFull) atomicOr(&ctr->flags, kFlagCountOverflow);
                }
            }
        }
        __syncthreads();
        uint32_t run = (uint32_t)base_s + wbase + incl - sum;
#pragma unroll
        for (int k = 0; k < kXsItems; ++k) {
            if (i0 + k < n) out[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------
// K2b: per (read, locus) pair, sum the read's events anchored inside the locus window
// (call.rs:388,394,400: start < P && P < end) and scatter the packed call into its bucket.
__global__ void __launch_bounds__(256)
k_pair_eval(ReadView rv, LocusView lv, int unphased, const uint32_t *__restrict__ cand_lo,
            const uint32_t *__restrict__ cand_n, const uint2 *__restrict__ events,
            const uint32_t *__restrict__ ev_off, uint64_t ev_cap, const uint32_t *__restrict__ bucket_off,
            uint32_t *__restrict__ bucket_cnt, uint64_t *__restrict__ vals, uint64_t vals_cap,
            DevCounters *__restrict__ ctr)
{
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    int lo = 0, n = 0;
    int32_t rs = 0, re = 0;
    uint32_t h = 0, e0 = 0, e1 = 0;
    bool is2d = false;
    if (r < rv.R) {
        n = (int)cand_n[r];
        if (n) {
            lo = (int)cand_lo[r];
            rs = rv.rs[r];
            re = rv.re[r];
            h = rv.hp[r];
            is2d = (rv.flags[r] & 1u) != 0;
            // if the speculative event list overflowed the run is repeated; stay inside the allocation
            e0 = (uint32_t)min((uint64_t)ev_off[r], ev_cap);
            e1 = (uint32_t)min((uint64_t)ev_off[r + 1], ev_cap);
        }
    }
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    for (int j = 0; j < nmax; ++j) {
        uint32_t bucket = 0xFFFFFFFFu;
        uint64_t key = 0;
        if (j < n) {
            const int l = lo + j;
            const int32_t ls = __ldg(lv.start + l), le = __ldg(lv.end + l);
            if (pair_passes(unphased != 0, rs, re, ls, le)) {
                if (unphased) bucket = 2u * (uint32_t)l;
                else if (h == 1u || h == 2u) bucket = 2u * (uint32_t)l + (h - 1u);
            }
            if (bucket != 0xFFFFFFFFu) {
                const uint32_t start_ext = (uint32_t)ls - 10u, end_ext = (uint32_t)le + 10u;
                // first event with pos1 > start_ext
                uint32_t a = e0, b = e1;
                while (a < b) {
                    const uint32_t m = a + ((b - a) >> 1);
                    if (events[m].x > start_ext) b = m; else a = m + 1;
                }
                int64_t call = 0;
                uint32_t clip = 0;
                for (uint32_t e = a; e < e1; ++e) {
                    const uint2 ev = events[e];
                    if (!(ev.x < end_ext)) break;
                    const int32_t v = (int32_t)ev.y;
                    const uint32_t is_s = (uint32_t)v & 1u;
                    if (is_s && is2d) continue;              // call.rs:394 !is_accidental_2d(&r)
                    call += (int64_t)(v >> 1);
                    clip |= is_s;
                }
                key = ((uint64_t)(call + kCallBias) << 1) | clip;
            }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
        if (bucket != 0xFFFFFFFFu) {
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if ((int)lane_id() == leader) old = atomicSub(bucket_cnt + bucket, (uint32_t)__popc(peers));
            old = __shfl_sync(peers, old, leader);
            const uint32_t rank = __popc(peers & lanemask_lt());
            const uint64_t slot = (uint64_t)bucket_off[bucket] + (old - 1u - rank);
            if (slot < vals_cap) vals[slot] = key;
            else atomicOr(&ctr->flags, kFlagValsOverflow);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// K3: per-locus sort / split / support filter / median (call.rs:308-321,365-369,497-522).
// Keys are 63-bit ((call + bias) << 1 | clip): ascending key order == (value, Span before Clip).
// Phased loci set bit 63 on H2 keys so one sort orders [H1 | H2]; unphased loci split the sorted
// run at n/2 (call.rs:314).

// all-ascending bitonic network over K striped registers per lane (element i = k*32 + lane)
template <int K>
__device__ __forceinline__ void warp_sort(uint64_t (&key)[K])
{
    constexpr int N = K * 32;
#pragma unroll
    for (int blk = 2; blk <= N; blk <<= 1) {
        // mirror step: partner = i ^ (blk - 1)
        {
            const int lane_x = (blk - 1) & 31, reg_x = (blk - 1) >> 5;
            uint64_t other[K];
#pragma unroll
            for (int k = 0; k < K; ++k) other[k] = __shfl_xor_sync(0xffffffffu, key[k ^ reg_x], lane_x);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = k * 32 + (int)lane_id();
                const bool lower = (i & (blk >> 1)) == 0;      // lower half of the mirrored block keeps the min
                const uint64_t a = key[k], b = other[k];
                key[k] = lower ? (a < b ? a : b) : (a > b ? a : b);
            }
        }
#pragma unroll
        for (int d = blk >> 2; d >= 1; d >>= 1) {
            const int lane_x = d & 31, reg_x = d >> 5;
            uint64_t other[K];
#pragma unroll
            for (int k = 0; k < K; ++k) other[k] = __shfl_xor_sync(0xffffffffu, key[k ^ reg_x], lane_x);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = k * 32 + (int)lane_id();
                const bool lower = (i & d) == 0;
                const uint64_t a = key[k], b = other[k];
                key[k] = lower ? (a < b ? a : b) : (a > b ? a : b);
            }
        }
    }
}

__device__ __forceinline__ int64_t key_call(uint64_t key)
{
    return (int64_t)((key & ~kKeyHapBit) >> 1) - kCallBias;
}

// median of the sorted sub-range [a, a+n) held striped across the warp. Returns validity;
// *twice = 2 x median. call.rs:497-522.
template <int K>
__device__ __forceinline__ bool warp_median_part(const uint64_t (&key)[K], uint32_t a, uint32_t n, uint32_t support,
                                                 int64_t *twice, bool *panicked)
{
    *twice = 0;
    if (n < support) return false;                                   // call.rs:498-500
    uint32_t span_b[K], clip_b[K];
    uint32_t s = 0, c = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = k * 32 + lane_id();
        const bool in = (i >= a) && (i < a + n);
        span_b[k] = __ballot_sync(0xffffffffu, in && !(key[k] & 1ull));
        clip_b[k] = __ballot_sync(0xffffffffu, in && (key[k] & 1ull));
        s += __popc(span_b[k]);
        c += __popc(clip_b[k]);
    }
    const uint32_t topk = (s <= support) ? (support - s) : 0u;       // call.rs:509-513
    const uint32_t m = s + topk;
    if (m == 0) { *panicked = true; return false; }                  // call.rs:516 on an empty vector
    const uint32_t t2 = m >> 1, t1 = (m & 1u) ? t2 : t2 - 1u;        // call.rs:515-521
    uint32_t clip_before = 0, sel_before = 0;
    int64_t contrib = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t me = 1u << lane_id();
        const bool is_span = (span_b[k] & me) != 0, is_clip = (clip_b[k] & me) != 0;
        const uint32_t clip_idx = clip_before + __popc(clip_b[k] & lanemask_lt());
        const bool sel = is_span || (is_clip && clip_idx + topk >= c);   // the topk largest clips
        const uint32_t sel_b = __ballot_sync(0xffffffffu, sel);
        const uint32_t rank = sel_before + __popc(sel_b & lanemask_lt());
        if (sel) {
            const int64_t v = key_call(key[k]);
            if (rank == t1) contrib += v;
            if (rank == t2) contrib += v;
        }
        clip_before += __popc(clip_b[k]);
        sel_before += __popc(sel_b);
    }
    *twice = warp_sum(contrib);
    return true;
}

template <int K>
__device__ __forceinline__ void warp_locus(const uint64_t *__restrict__ vals, uint32_t b0, uint32_t n1, uint32_t ntot,
                                           bool phased, uint32_t support, int64_t *t1, int64_t *t2, uint32_t *valid,
                                           bool *panicked)
{
    uint64_t key[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = k * 32 + lane_id();
        uint64_t v = kKeyInf;
        if (i < ntot) {
            v = vals[b0 + i];
            if (phased && i >= n1) v |= kKeyHapBit;
        }
        key[k] = v;
    }
    warp_sort<K>(key);
    const bool v1 = warp_median_part<K>(key, 0, n1, support, t1, panicked);
    const bool v2 = warp_median_part<K>(key, n1, ntot - n1, support, t2, panicked);
    *valid = (v1 ? 1u : 0u) | (v2 ? 2u : 0u);
}

constexpr int kMedianWarpMax = 128;         // loci with more calls go to the CTA kernel

__global__ void __launch_bounds__(256)
k_locus_median(uint32_t L, int unphased, uint32_t support, const uint32_t *__restrict__ bucket_off,
               const uint64_t *__restrict__ vals, int64_t *__restrict__ twice_h1, int64_t *__restrict__ twice_h2,
               uint8_t *__restrict__ valid, uint32_t *__restrict__ big_list, DevCounters *__restrict__ ctr)
{
    const uint32_t l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (l >= L) return;
    const uint32_t b0 = bucket_off[2 * l], b1 = bucket_off[2 * l + 1], b2 = bucket_off[2 * l + 2];
    const uint32_t ntot = b2 - b0;
    const uint32_t n1 = unphased ? (ntot >> 1) : (b1 - b0);        // call.rs:314 split_at(len/2)
    int64_t t1 = 0, t2 = 0;
    uint32_t vm = 0;
    bool panicked = false;
    if (ntot <= 32) warp_locus<1>(vals, b0, n1, ntot, !unphased, support, &t1, &t2, &vm, &panicked);
    else if (ntot <= 64) warp_locus<2>(vals, b0, n1, ntot, !unphased, support, &t1, &t2, &vm, &panicked);
    else if (ntot <= kMedianWarpMax) warp_locus<4>(vals, b0, n1, ntot, !unphased, support, &t1, &t2, &vm, &panicked);
    else {
        if (lane_id() == 0) big_list[atomicAdd(&ctr->big_count, 1u)] = l;
        return;
    }
    if (lane_id() == 0) {
        twice_h1[l] = t1;
        twice_h2[l] = t2;
        valid[l] = (uint8_t)vm;
        if (panicked) atomicOr(&ctr->flags, kFlagMedianEmpty);
    }
}

// CTA-wide path for loci with more than kMedianWarpMax calls: bitonic sort in shared memory when
// the run fits, otherwise in place in the (scratch) vals array; then a block-wide selection pass.
constexpr int kBigThreads = 256;
constexpr int kBigSmemKeys = 4096;

__device__ __forceinline__ void block_part_median(volatile uint64_t *a, uint32_t lo, uint32_t n, uint32_t support,
                                                  int64_t *twice, bool *ok, bool *panicked, uint32_t *sh_cnt,
                                                  unsigned long long *sh_acc)
{
    // sh_cnt[0..1]: spans, clips ; sh_cnt[2]: running selected count ; sh_acc: sum of the median elements
    const uint32_t tid = threadIdx.x;
    *twice = 0;
    *ok = false;
    if (n < support) return;
    if (tid == 0) { sh_cnt[0] = 0; sh_cnt[1] = 0; sh_cnt[2] = 0; sh_cnt[3] = 0; *sh_acc = 0ull; }
    __syncthreads();
    uint32_t my_s = 0, my_c = 0;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        if (a[lo + i] & 1ull) ++my_c; else ++my_s;
    }
    my_s = warp_sum(my_s);
    my_c = warp_sum(my_c);
    if (lane_id() == 0) { atomicAdd(&sh_cnt[0], my_s); atomicAdd(&sh_cnt[1], my_c); }
    __syncthreads();
    const uint32_t s = sh_cnt[0], c = sh_cnt[1];
    const uint32_t topk = (s <= support) ? (support - s) : 0u;
    const uint32_t m = s + topk;
    if (m == 0) { *panicked = true; return; }
    const uint32_t t2 = m >> 1, t1 = (m & 1u) ? t2 : t2 - 1u;
    __shared__ uint32_t wtot[2][kBigThreads / 32];
    // chunked pass in sorted order; running counts of clips (sh_cnt[3]) and selected (sh_cnt[2])
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + tid;
        const bool in = i < n;
        const uint64_t k = in ? a[lo + i] : 0ull;
        const bool is_clip = in && (k & 1ull), is_span = in && !(k & 1ull);
        const uint32_t cb = __ballot_sync(0xffffffffu, is_clip);
        if (lane_id() == 0) wtot[0][tid >> 5] = __popc(cb);
        __syncthreads();
        uint32_t clip_idx = sh_cnt[3] + __popc(cb & lanemask_lt());
        for (uint32_t w = 0; w < (tid >> 5); ++w) clip_idx += wtot[0][w];
        const bool sel = is_span || (is_clip && clip_idx + topk >= c);
        const uint32_t sb = __ballot_sync(0xffffffffu, sel);
        if (lane_id() == 0) wtot[1][tid >> 5] = __popc(sb);
        __syncthreads();
        uint32_t rank = sh_cnt[2] + __popc(sb & lanemask_lt());
        for (uint32_t w = 0; w < (tid >> 5); ++w) rank += wtot[1][w];
        if (sel) {
            const int64_t v = key_call(k);
            if (rank == t1) atomicAdd(sh_acc, (unsigned long long)v);
            if (rank == t2) atomicAdd(sh_acc, (unsigned long long)v);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t ct = 0, st = 0;
            for (uint32_t w = 0; w < blockDim.x / 32; ++w) { ct += wtot[0][w]; st += wtot[1][w]; }
            sh_cnt[3] += ct;
            sh_cnt[2] += st;
        }
        __syncthreads();
    }
    *twice = (int64_t)*sh_acc;
    *ok = true;
}

__global__ void __launch_bounds__(kBigThreads)
k_locus_median_big(int unphased, uint32_t support, const uint32_t *__restrict__ bucket_off, uint64_t *__restrict__ vals,
                   int64_t *__restrict__ twice_h1, int64_t *__restrict__ twice_h2, uint8_t *__restrict__ valid,
                   const uint32_t *__restrict__ big_list, DevCounters *__restrict__ ctr)
{
    __shared__ uint64_t skeys[kBigSmemKeys];
    __shared__ uint32_t item_s;
    __shared__ uint32_t sh_cnt[4];
    __shared__ unsigned long long sh_acc;
    const uint32_t tid = threadIdx.x;
    while (true) {
        if (tid == 0) item_s = atomicAdd(&ctr->big_cursor, 1u);
        __syncthreads();
        const uint32_t item = item_s;
        if (item >= ctr->big_count) break;
        const uint32_t l = big_list[item];
        const uint32_t b0 = bucket_off[2 * l], b1 = bucket_off[2 * l + 1], b2 = bucket_off[2 * l + 2];
        const uint32_t ntot = b2 - b0;
        const uint32_t n1 = unphased ? (ntot >> 1) : (b1 - b0);
        volatile uint64_t *a;
        if (ntot <= (uint32_t)kBigSmemKeys) {
            for (uint32_t i = tid; i < ntot; i += blockDim.x) {
                uint64_t v = vals[b0 + i];
                if (!unphased && i >= n1) v |= kKeyHapBit;
                skeys[i] = v;
            }
            a = skeys;
        } else {
            if (!unphased)
                for (uint32_t i = n1 + tid; i < ntot; i += blockDim.x) vals[b0 + i] |= kKeyHapBit;
            a = vals + b0;
        }
        __syncthreads();
        // all-ascending bitonic network with virtual +inf padding (comparators touching i >= ntot are no-ops)
        uint32_t npow = 1;
        while (npow < ntot) npow <<= 1;
        for (uint32_t blk = 2; blk <= npow; blk <<= 1) {
            for (uint32_t i = tid; i < npow; i += blockDim.x) {
                const uint32_t j = i ^ (blk - 1);
                if (i < j && j < ntot) {
                    const uint64_t x = a[i], y = a[j];
                    if (x > y) { a[i] = y; a[j] = x; }
                }
            }
            __syncthreads();
            for (uint32_t d = blk >> 2; d >= 1; d >>= 1) {
                for (uint32_t i = tid; i < npow; i += blockDim.x) {
                    const uint32_t j = i ^ d;
                    if (i < j && j < ntot) {
                        const uint64_t x = a[i], y = a[j];
                        if (x > y) { a[i] = y; a[j] = x; }
                    }
                }
                __syncthreads();
            }
        }
        int64_t t1, t2;
        bool ok1, ok2, panicked = false;
        block_part_median(a, 0, n1, support, &t1, &ok1, &panicked, sh_cnt, &sh_acc);
        __syncthreads();
        block_part_median(a, n1, ntot - n1, support, &t2, &ok2, &panicked, sh_cnt, &sh_acc);
        if (tid == 0) {
            twice_h1[l] = t1;
            twice_h2[l] = t2;
            valid[l] = (uint8_t)((ok1 ? 1u : 0u) | (ok2 ? 2u : 0u));
            if (panicked) atomicOr(&ctr->flags, kFlagMedianEmpty);
        }
        __syncthreads();
    }
}

}  // namespace inq
