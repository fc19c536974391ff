// inq_device.cuh -- sm_100a kernels of the `inquiSTR call` hot path.
//
// Reference semantics being reproduced (files under /root/reference/src, v0.13.0):
//   K1  k_join_ranges     read x locus overlap join (ranges)   call.rs:288,338
//   K2  k_cigar_scan      segmented CIGAR scan -> event list   call.rs:377-413 (position cursor + op tests)
//   K2b k_pair_eval       filter, window sum, bucket scatter   call.rs:297-301,349-355,388-403,304,358
//   K3  k_locus_median*   sort / split / support / median      call.rs:308-321,365-369,497-522
//
// Layout in HBM (all structure-of-arrays):
//   reads : contig/ref_start/ref_end int32[R], mapq/hp/flags u8[R], cig_off u64[R+1], cigar u32[C]
//   loci  : start/end/pmax_end int32[L] sorted by (contig,start), contig_off int64[n_contigs+1]
//   events: uint2{pos1 (1-based anchor, u32), val = (signed len << 1) | is_softclip}[E], ev_off u32[R+1]
//   calls : one segment per locus sized by its candidate count (seg_off u32[L+1]); H1 from the front,
//           H2 from the back (cursor u64[L] = front count | back count << 32); vals u64[#candidates]
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace inq {

// ----------------------------------------------------------------------------------------------
// constants
constexpr int kTileWords = 8192;            // padding granularity of the device CIGAR stream (32 KB)

constexpr uint32_t kFlagBadHp = 1u << 0;
constexpr uint32_t kFlagMedianEmpty = 1u << 1;
constexpr uint32_t kFlagEventOverflow = 1u << 2;
constexpr uint32_t kFlagValsOverflow = 1u << 3;
constexpr uint32_t kFlagCountOverflow = 1u << 4;
constexpr uint32_t kFlagBadSa = 1u << 5;

constexpr uint64_t kDescAggregate = 1ull << 62;
constexpr uint64_t kDescPrefix = 2ull << 62;
constexpr uint64_t kDescValueMask = (1ull << 62) - 1;

constexpr int64_t kCallBias = 1ll << 61;    // keys are (call + bias) << 1 | clip, 63 bits
constexpr uint64_t kKeyHapBit = 1ull << 63;
constexpr uint64_t kKeyInf = ~0ull;

// device-side counters, one struct per ctx
// statistics are accumulated per block into one of kStatSlots copies (a single address would
// serialise hundreds of thousands of atomics in L2); the host adds the copies up
constexpr int kStatSlots = 64;
enum { ST_CANDIDATES = 0, ST_PAIRS, ST_READS_JOINED, ST_WORDS_JOINED, ST_OP_VISITS, ST_COUNT = 8 };

constexpr int kMaxRanges = 16;              // the CIGAR stream is scanned in up to this many ranges of warp tiles (see inq_genotype)
constexpr int kMaxMedianChunks = 48;        // the medians run in catalog chunks so that the result copy of one overlaps the next
struct DevCounters {
    unsigned long long stat[kStatSlots][ST_COUNT];
    unsigned long long n_events;
    unsigned long long ev_alloc;      // event slots handed out to warps (chunks)
    unsigned long long wt_carry[kMaxRanges + 1][2];   // warp-tile prefix {consumption, events} at the start of every scanned range
    unsigned int flags;
    unsigned int tile_counter;
    unsigned int scan_counter[4];
    unsigned int wt_scan_counter[kMaxRanges];  // dynamic tile ids of k_exclusive_scan2, one per range
    unsigned int big_count[kMaxMedianChunks];  // per chunk of the catalog (see inq_genotype): loci for the CTA path
    unsigned int big_cursor[kMaxMedianChunks];
    unsigned int bad_hp_value;        // diagnostics: HP value and read index of one offending read
    unsigned long long bad_hp_read;
    unsigned long long bad_sa_read;
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

// mbarrier helpers for the TMA tensor loads of k_cigar_scan
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// poll with a short sleep in between so that waiting warps do not burn issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
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

// ---- the same two lower bounds, found by the warp together (k_join_ranges runs under the HBM-saturating CIGAR scan,
// where a dependent L2 round trip costs ~1 us: 2 x 17 of them per read are what the kernel's time is made of).
// Reads arrive coordinate-sorted, so the 32 answers of a warp lie within a few catalog entries of each other:
//   1. one 33-ary search by all lanes for the smallest key of the warp (4-5 round trips instead of 17),
//   2. the 32 entries from there on are loaded once, one per lane, and every lane finds its own bound among them
//      with shuffles (no memory),
//   3. the second bound lies within a 32-entry window of the first one; same trick.
// A lane whose answer falls outside a window (unsorted reads, very long reads) finishes with the plain binary search
// over the rest, and a warp that straddles contigs takes the scalar path: the results are the unique lower bounds
// either way.
__device__ __forceinline__ int warp_lower_bound(const int32_t *__restrict__ a, int lo, int hi, int64_t key)
{
    const int lane = (int)lane_id();
    while (hi - lo > 32) {
        const int n = hi - lo;
        const int p = lo + (int)(((int64_t)(lane + 1) * n) / 33);            // lo < p < hi, increasing with the lane
        const uint32_t m = __ballot_sync(0xffffffffu, (int64_t)__ldg(a + p) >= key);
        const int f = m ? __ffs(m) - 1 : 32;                                 // first probe that is >= key
        const int new_hi = f < 32 ? lo + (int)(((int64_t)(f + 1) * n) / 33) : hi;
        const int new_lo = f > 0 ? lo + (int)(((int64_t)f * n) / 33) + 1 : lo;
        lo = new_lo;
        hi = new_hi;
    }
    const int p = lo + lane;
    const uint32_t m = __ballot_sync(0xffffffffu, p < hi && (int64_t)__ldg(a + p) >= key);
    return m ? lo + __ffs(m) - 1 : hi;
}

// number of entries of the window a[win, win + wn) (wn <= 32, entry j held by lane j in `v`) that are < key
__device__ __forceinline__ int window_count_less(int32_t v, int wn, int64_t key)
{
    int cnt = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int32_t probe = __shfl_sync(0xffffffffu, v, (cnt + step - 1) & 31);
        if (cnt + step - 1 < wn && (int64_t)probe < key) cnt += step;
    }
    const int32_t last = __shfl_sync(0xffffffffu, v, 31);
    if (cnt == 31 && wn == 32 && (int64_t)last < key) cnt = 32;
    return cnt;
}

// all 32 lanes call this; `act` lanes carry a read on contig c (warp-uniform among the active lanes)
__device__ __forceinline__ void candidate_range_warp(bool unphased, const LocusView &lv, int c, bool act, int32_t rs, int32_t re,
                                                     int &lo, int &hi)
{
    const int lane = (int)lane_id();
    const int l0 = (int)lv.contig_off[c], l1 = (int)lv.contig_off[c + 1];
    // first bound: over start[l0, l1)
    const int64_t key1 = unphased ? (int64_t)rs + 10 : (int64_t)re + 10;
    const int64_t big = 0x7fffffffffffffffll;
    int64_t kmin = act ? key1 : big;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { const int64_t o = __shfl_xor_sync(0xffffffffu, kmin, d); kmin = o < kmin ? o : kmin; }
    const int p1 = warp_lower_bound(lv.start, l0, l1, kmin);
    const int wn1 = min(32, l1 - p1);
    const int32_t v1 = lane < wn1 ? __ldg(lv.start + p1 + lane) : 0;
    const int c1 = window_count_less(v1, wn1, key1);
    int b1 = p1 + c1;
    if (act && c1 == 32 && p1 + 32 < l1) b1 = lower_bound_i32(lv.start, p1 + 32, l1, key1);
    if (unphased) {
        lo = b1;                                                             // start - 10 >= rs
        // hi = lower_bound(start[lo, l1), re - 9): at or after the warp's smallest lo
        const int64_t key2 = (int64_t)re - 10 + 1;
        const int w2 = __reduce_min_sync(0xffffffffu, act ? lo : 0x7fffffff);
        const bool any = w2 != 0x7fffffff;
        const int wn2 = any ? min(32, l1 - w2) : 0;
        const int32_t v2 = lane < wn2 ? __ldg(lv.start + w2 + lane) : 0;
        const int c2 = window_count_less(v2, wn2, key2);
        int b2 = w2 + c2;
        if (act && c2 == 32 && w2 + 32 < l1) b2 = lower_bound_i32(lv.start, w2 + 32, l1, key2);
        hi = b2 > lo ? b2 : lo;
    } else {
        hi = b1;                                                             // start - 10 < re
        // lo = lower_bound(pmax[l0, hi), rs - 9): at most the warp's largest hi, usually within 32 entries below it
        const int64_t key2 = (int64_t)rs - 10 + 1;
        const int hmax = __reduce_max_sync(0xffffffffu, act ? hi : -1);
        const int w2 = hmax - 32 > l0 ? hmax - 32 : l0;
        const int wn2 = hmax > w2 ? min(32, hmax - w2) : 0;
        const int32_t v2 = lane < wn2 ? __ldg(lv.pmax + w2 + lane) : 0;
        const int c2 = window_count_less(v2, wn2, key2);
        int b2 = w2 + c2;
        if (act && c2 == 0 && w2 > l0) b2 = lower_bound_i32(lv.pmax, l0, w2, key2);
        lo = b2 < hi ? b2 : hi;
    }
    if (hi < lo) hi = lo;
}

// K1: per read, the contiguous range [lo, lo+n) of catalog loci whose window it can pair with
// (two lower bounds over the sorted, L2-resident catalog). Instead of visiting the candidates,
// the kernel adds +1/-1 to a difference array; its prefix sum is the number of candidate reads of
// every locus, which sizes that locus' segment of the call buffer (an upper bound on its pairs).
__global__ void __launch_bounds__(256)
k_join_ranges(ReadView rv, LocusView lv, int unphased, int coop, uint32_t *__restrict__ cand_lo, uint32_t *__restrict__ cand_n,
              uint32_t *__restrict__ delta /* L+1, zeroed */, DevCounters *__restrict__ ctr, unsigned long long *__restrict__ trace)
{
    // (trace, experiments only: globaltimer at the start and the end of every CTA, and the SM it ran on)
    if (trace && threadIdx.x == 0) {
        unsigned long long t; unsigned int sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        trace[3ull * blockIdx.x] = t;
        trace[3ull * blockIdx.x + 2] = sm;
    }
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    int lo = 0, n = 0;
    int c = -1;
    int32_t rs = 0, re = 0;
    bool act = false;
    if (r < rv.R) {
        c = rv.contig[r];
        const uint32_t mq = rv.mapq[r], h = rv.hp[r];
        rs = rv.rs[r];
        re = rv.re[r];
        act = c >= 0 && c < lv.n_contigs && mq > 10u && (unphased || h != 0xFFu);
    }
    // one contig for the whole warp? (all warps but the few that straddle a contig boundary)
    const int cmin = __reduce_min_sync(0xffffffffu, act ? c : 0x7fffffff), cmax = __reduce_max_sync(0xffffffffu, act ? c : -1);
    int hi = 0;
    if (coop && cmin == cmax) {
        candidate_range_warp(unphased != 0, lv, cmin, act, rs, re, lo, hi);
    } else if (act) {
        candidate_range(unphased != 0, lv, c, rs, re, lo, hi);
    }
    if (r < rv.R) {
        if (act) n = hi - lo;
        else lo = 0;
        cand_lo[r] = (uint32_t)lo;
        cand_n[r] = (uint32_t)n;
    }
    // +1 at lo, -1 at hi: neighbouring reads share their bounds, so the lanes that hit the same entry send one atomic
    // (11.6 M same-address-heavy atomics per pass were what kept this kernel busy for the whole CIGAR scan)
    {
        const bool has = n > 0;
        const uint32_t mlo = __match_any_sync(0xffffffffu, has ? (uint32_t)lo : 0xFFFFFFFFu);
        const uint32_t mhi = __match_any_sync(0xffffffffu, has ? (uint32_t)hi : 0xFFFFFFFFu);
        if (has) {
            const uint32_t lane = lane_id();
            if ((uint32_t)__ffs(mlo) - 1u == lane) atomicAdd(delta + lo, (uint32_t)__popc(mlo));
            if ((uint32_t)__ffs(mhi) - 1u == lane) atomicAdd(delta + hi, 0u - (uint32_t)__popc(mhi));   // -count (mod 2^32)
        }
    }
    const uint32_t cand_w = __reduce_add_sync(0xffffffffu, (uint32_t)n);
    if (lane_id() == 0 && cand_w)
        atomicAdd(&ctr->stat[(blockIdx.x * 8u + (threadIdx.x >> 5)) % kStatSlots][ST_CANDIDATES], (unsigned long long)cand_w);
    if (trace) {
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            trace[3ull * blockIdx.x + 1] = t;
        }
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
// K2: CIGAR scan. Persistent warps stream 4 KB "warp tiles" (1024 words) of the flat packed-CIGAR array
// through per-warp 2-stage shared-memory rings filled by TMA (2-D tensor map, 128B swizzle); each lane
// owns 32 consecutive words. One pass over the words gives the lane's reference consumption
// (call.rs:384-392,404) and its event mask (I/D/S ops longer than minlen, call.rs:388,394,400); two warp
// scans turn that into warp-local prefixes. Nothing in this kernel depends on another CTA or warp:
// it writes
//   rd_pre : warp-tile-local exclusive prefixes {consumption, events} at the first word of every read
//            that starts inside the tile (tile_first, built when reads are pushed, names those reads)
//   wt     : totals per warp tile (prefix-summed afterwards by k_exclusive_scan2)
//   evraw  : {bases consumed inside the warp tile before the op, (signed len << 1) | is_S}
//            for every event, stored per warp tile in chunks handed out by one atomic
// and k_read_fixup turns these into per-read event lists with absolute anchors.
// Each CIGAR word is read from HBM exactly once.
struct ScanParams {
    const uint32_t *tile_first;   // [n_wt + 1] first read (index into cig_off, sentinel R included) starting in warp tile t or later
    const uint64_t *cig_off;      // [R + 1] first CIGAR word of every read (+ sentinel)
    uint2 *rd_pre;                // [R + 1] {consumption, events} inside the read's warp tile before its first word
    uint2 *wt;                    // [n_wt] {consumption, events} totals per warp tile (prefix-summed by k_exclusive_scan2 into a second array)
    uint32_t *wt_sbase;           // [n_wt] storage slot of the warp tile's first event
    uint2 *evraw;
    DevCounters *ctr;
    uint64_t raw_cap;             // capacity of evraw (slots)
    uint64_t wt_begin;            // this launch scans warp tiles [wt_begin, n_wt)
    uint64_t n_wt;
    uint32_t thr;                 // (min(minlen, 2^28 - 1) << 4) | 15: an op is longer than minlen iff its word > thr
    uint32_t neg1;                // 0xFFFFFFFF as a run-time value (keeps thr - w a multiply-add on the FMA pipe)
    uint32_t evict_first;         // 1: the CIGAR words are loaded with the L2 evict-first policy
    uint32_t debug;               // timing experiments only, compiled in with -DINQ_TIMING_EXPERIMENTS: results are wrong when != 0
};
#ifdef INQ_TIMING_EXPERIMENTS
#define INQ_DBG(x) (x)
#else
#define INQ_DBG(x) 0u             // the shipped library cannot be switched into a wrong-answer mode
#endif

#ifndef INQ_LANE_WORDS
#define INQ_LANE_WORDS 32
#endif
constexpr int kLaneWords = INQ_LANE_WORDS;               // consecutive CIGAR words per lane (16 or 32)
constexpr int kWarpTileWords = 32 * kLaneWords;         // 1024 words = 4 KB = one TMA box
#ifndef INQ_SCAN_WARPS
#define INQ_SCAN_WARPS 24
#endif
#ifndef INQ_WARP_STAGES
#define INQ_WARP_STAGES 2
#endif
#ifndef INQ_SCAN_MIN_CTAS
#define INQ_SCAN_MIN_CTAS 1
#endif
constexpr int kScanWarps = INQ_SCAN_WARPS;              // warps per CTA, each fully autonomous
constexpr int kCtaThreads = kScanWarps * 32;
constexpr int kWarpStages = INQ_WARP_STAGES;             // per-warp ring of TMA boxes
// 64-bit LUT read through one wrap-mode funnel shift by 2*op: result bit 0 = op consumes the reference
// (M,D,N,=,X: lo bit 2*op), result bit 31 = op is I, D or S (hi bit 2*op-1)
constexpr uint32_t kOpLutLo = (1u << 0) | (1u << 4) | (1u << 6) | (1u << 14) | (1u << 16);
constexpr uint32_t kOpLutHi = (1u << 1) | (1u << 3) | (1u << 7);
constexpr uint32_t kEvChunk = 1024;                     // event slots a warp takes per atomic
// cig_consume through the LUT: branch-free, the multiply issues on the FMA pipe
__device__ __forceinline__ uint32_t cig_consume_lut(uint32_t w)
{
    return (w >> 4) * (__funnelshift_r(kOpLutLo, kOpLutHi, w + w) & 1u);
}

struct ScanSmem {
    alignas(1024) uint32_t stage[kScanWarps][kWarpStages][kWarpTileWords];   // 128B-swizzled by the TMA tensor map
    uint32_t snap[kScanWarps][kLaneWords / 4][32];                            // per-quad cursor snapshots of every lane
    uint16_t qlist[kScanWarps][32];                                           // word positions of one round of events
    alignas(8) uint64_t full[kScanWarps][kWarpStages];                        // TMA landed (tx bytes)
};
constexpr size_t kScanSmemBytes = sizeof(ScanSmem) + 1024;   // slack to align the swizzled stages to 1 KB

// word index inside a warp tile -> word index in the 128B-swizzled stage buffer
// (16-byte chunk index bits [2:4] ^= 128-byte row index bits [5:7])
__device__ __forceinline__ uint32_t swz(uint32_t idx) { return idx ^ (((idx >> 5) & 7u) << 2); }

// `policy`: L2 cache policy of the load, 0 = none. The stream is read exactly once, so its lines are marked evict-first:
// otherwise 10 GB of them flush the catalog and the per-read arrays out of the 126 MB L2, and every probe of the join
// chain that runs next to the scan (and of the pair kernel after it) goes to the saturated HBM instead.
__device__ __forceinline__ void tma_load_tile(void *dst_smem, const CUtensorMap *tmap, uint32_t row, uint64_t *bar, uint64_t policy)
{
    if (policy) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                     ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row), "r"(smem_u32(bar)), "l"(policy)
                     : "memory");
    } else {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(0), "r"(row), "r"(smem_u32(bar))
                     : "memory");
    }
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Every warp is an independent pipeline: it owns warp tiles gwid, gwid + W, gwid + 2W, ... (W = warps
// in the grid), a kWarpStages-deep ring of 4 KB shared-memory boxes (one warp tile each) that it fills itself with TMA, and the
// mbarriers of that ring. There is no block-level synchronisation and no producer warp.
template <bool kThrHigh>
__global__ void __launch_bounds__(kCtaThreads, INQ_SCAN_MIN_CTAS)
k_cigar_scan(const __grid_constant__ CUtensorMap tmap, ScanParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // keep the pointer in the shared address space (LDS/STS, not generic loads)
    ScanSmem &sm = *reinterpret_cast<ScanSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t kBoxBytes = kWarpTileWords * 4;
    constexpr uint32_t kRowsPerBox = kWarpTileWords / 32;       // 128-byte rows
    const uint64_t gwid = p.wt_begin + (uint64_t)blockIdx.x * kScanWarps + warp, stride = (uint64_t)gridDim.x * kScanWarps;
    uint64_t *full = sm.full[warp];
    const uint64_t l2_policy = p.evict_first ? l2_evict_first_policy() : 0ull;

    if (lane == 0) {
        for (int s = 0; s < kWarpStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        for (int s = 0; s < kWarpStages; ++s) {
            const uint64_t gw = gwid + (uint64_t)s * stride;
            if (gw < p.n_wt) {
                mbar_expect_tx(&full[s], kBoxBytes);
                tma_load_tile(sm.stage[warp][s], &tmap, (uint32_t)gw * kRowsPerBox, &full[s], l2_policy);
            }
        }
    }
    __syncwarp();

    const uint32_t thr = p.thr;                                 // (w >> 4) > minlen  <=>  w > thr
    const uint32_t neg1 = p.neg1;
    // this lane's kLaneWords consecutive words: quads q0..q0+kLaneWords/4-1 of the box; quad q lives in
    // 128-byte row q/8 at chunk (q%8) ^ (row%8) (128B swizzle) -> conflict-free LDS.128 across the warp
    constexpr uint32_t kQuads = kLaneWords / 4;
    const uint32_t q0 = lane * kQuads;
    uint64_t chunk_cur = 0, chunk_end = 0;                      // this warp's private range of event slots
    // reads that start inside the current warp tile: [tf_lo, tf_hi), lane i holds the first word of read
    // tf_lo + i. Both are fetched one iteration ahead so that their latency never shows.
    uint32_t tf_lo = 0, tf_hi = 0;
    uint64_t g_cur = 0;
    if (gwid < p.n_wt) {
        tf_lo = __ldg(p.tile_first + gwid);
        tf_hi = __ldg(p.tile_first + gwid + 1);
        if (tf_lo + lane < tf_hi) g_cur = __ldg(p.cig_off + tf_lo + lane);
    }

    for (uint32_t it = 0;; ++it) {
        const uint64_t gw = gwid + (uint64_t)it * stride;       // global warp-tile index
        if (gw >= p.n_wt) break;
        const uint32_t s = it % kWarpStages;
        uint32_t nf_lo = 0, nf_hi = 0;
        if (gw + stride < p.n_wt) {
            nf_lo = __ldg(p.tile_first + gw + stride);
            nf_hi = __ldg(p.tile_first + gw + stride + 1);
        }
        mbar_wait_backoff(&full[s], (it / kWarpStages) & 1u, 32u);
        const uint32_t *stage = sm.stage[warp][s];
        const uint4 *st4 = reinterpret_cast<const uint4 *>(stage);

        // ---- one pass over the lane's words. evrev: bit (31 - i) <-> word i is an event;
        //      snap[j]: bases consumed inside the lane's words before quad j (parked in shared memory below)
        uint32_t c = 0, evrev = 0;
        uint32_t snap[kQuads];
        if (!(INQ_DBG(p.debug) & 8u))
#pragma unroll
        for (int j = 0; j < (int)kQuads; ++j) {
            const uint32_t q = q0 + (uint32_t)j, row = q >> 3;
            const uint4 v = st4[row * 8 + ((q & 7u) ^ (row & 7u))];
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            snap[j] = c;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // one funnel shift in wrap mode indexes the LUT by 2*op without masking the op first
                const uint32_t w = w4[k], lut = __funnelshift_r(kOpLutLo, kOpLutHi, w + w);
                // borrow of thr - w (<=> w > thr) lands in the sign bit of d | w (thr < 2^31) or d & w
                // (thr >= 2^31); the subtraction is a multiply-add by a run-time -1 so that it issues on
                // the FMA pipe. One LOP3 ands it with the op class, one funnel shift appends the sign bit.
                uint32_t d;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(neg1), "r"(thr));
                const uint32_t t = kThrHigh ? (lut & (d & w)) : (lut & (d | w));
                evrev = __funnelshift_l(t, evrev, 1);
                // c += consumes ? len : 0 as one multiply-add (FMA pipe) instead of select + add (ALU pipe)
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c) : "r"(w >> 4), "r"(lut & 1u));
            }
        }
        else {
#pragma unroll
            for (int j = 0; j < (int)kQuads; ++j) snap[j] = 0;
        }
        const uint32_t ne = __popc(evrev);
        const uint32_t incl_c = warp_incl_scan(c), incl_e = warp_incl_scan(ne);
        const uint32_t excl_c = incl_c - c, excl_e = incl_e - ne;
        const uint32_t tot_e = __shfl_sync(0xffffffffu, incl_e, 31);
        uint64_t g_nxt = 0;
        if (nf_lo + lane < nf_hi) g_nxt = __ldg(p.cig_off + nf_lo + lane);

        // ---- event slots: contiguous per warp tile, taken from a warp-private chunk
        uint64_t sbase = chunk_cur;
        if (tot_e) {
            if (chunk_end - chunk_cur < tot_e) {
                unsigned long long base = 0;
                const uint32_t n = max(kEvChunk, tot_e);
                if (lane == 0) {
                    base = atomicAdd(&p.ctr->ev_alloc, (unsigned long long)n);
                    // wt_sbase holds 32-bit slots: more than 2^32 slots handed out (stranded chunk remainders included) is an error
                    if (base + n > 0xFFFFFFFFull) atomicOr(&p.ctr->flags, kFlagCountOverflow);
                }
                chunk_cur = __shfl_sync(0xffffffffu, base, 0);
                chunk_end = chunk_cur + n;
                sbase = chunk_cur;
            }
            chunk_cur += tot_e;
        }
        if (lane == 31 && !(INQ_DBG(p.debug) & 16u)) {
            p.wt[gw] = make_uint2(incl_c, incl_e);
            p.wt_sbase[gw] = (uint32_t)sbase;
        }

        // ---- position queries. Two things need the tile-local prefix at a word position: every event
        //      (about 1 % of the words -> {bases consumed before the op, value}) and the first word of every
        //      read that starts in the tile (-> rd_pre). They are dealt to the lanes 32 at a time (one round
        //      for a typical tile): the lane that owns the word supplies its scan values by shuffle and its
        //      per-quad snapshot through shared memory; at most 3 words are re-read from the stage.
#pragma unroll
        for (int j = 0; j < (int)kQuads; ++j) sm.snap[warp][j][lane] = snap[j];
        const uint32_t n_starts = tf_hi - tf_lo;
        const uint32_t n_q = tot_e + n_starts;                  // queries [0, tot_e) are the events in word order
        const uint32_t off_cur = (uint32_t)(g_cur - gw * kWarpTileWords);    // lane i: first word of read tf_lo + i
        uint32_t ev_left = evrev, ev_idx = excl_e;
        if (!(INQ_DBG(p.debug) & 1u))
        for (uint32_t base = 0; base < n_q; base += 32) {
            // owners list their events with index in [base, base + 32): word order = descending bits of evrev
            while (ev_left && ev_idx < base + 32u) {
                const uint32_t bit = 31u - (uint32_t)__clz(ev_left);
                ev_left ^= 1u << bit;
                sm.qlist[warp][ev_idx - base] = (uint16_t)(lane * kLaneWords + (kLaneWords - 1u - bit));
                ++ev_idx;
            }
            __syncwarp();
            const uint32_t qi = base + lane;
            const bool is_ev = qi < tot_e, act = qi < n_q;
            const uint32_t si = qi - tot_e;                     // index of the read start (when !is_ev)
            const uint32_t spos = __shfl_sync(0xffffffffu, off_cur, si & 31u);
            uint32_t pos = 0;
            if (is_ev) pos = sm.qlist[warp][lane];
            else if (act) pos = (si < 32u) ? spos : (uint32_t)(__ldg(p.cig_off + tf_lo + si) - gw * kWarpTileWords);
            const uint32_t owner = pos / kLaneWords, k = pos % kLaneWords;
            const uint32_t xc = __shfl_sync(0xffffffffu, excl_c, owner), xe = __shfl_sync(0xffffffffu, excl_e, owner);
            const uint32_t er = __shfl_sync(0xffffffffu, evrev, owner);
            if (act) {
                const uint32_t j = k >> 2, kk = k & 3u, q = owner * kQuads + j, row = q >> 3;
                const uint4 v = st4[row * 8 + ((q & 7u) ^ (row & 7u))];
                // a zero word (M, length 0) consumes nothing: blank the words at and after the position
                const uint32_t c_in = xc + sm.snap[warp][j][owner] + cig_consume_lut(kk > 0u ? v.x : 0u) +
                                      cig_consume_lut(kk > 1u ? v.y : 0u) + cig_consume_lut(kk > 2u ? v.z : 0u);
                if (is_ev) {
                    const uint32_t w = kk == 0u ? v.x : kk == 1u ? v.y : kk == 2u ? v.z : v.w;
                    const uint32_t len = w >> 4, op = w & 15u;
                    const int32_t val = (int32_t)(((op == 2u) ? (0u - len) : len) << 1) | (int32_t)(op == 4u);
                    const uint64_t slot = sbase + qi;
                    if (slot < p.raw_cap) p.evraw[slot] = make_uint2(c_in, (uint32_t)val);
                    else atomicOr(&p.ctr->flags, kFlagEventOverflow);
                } else if (pos) {
                    const uint32_t e_in = k ? __popc(er >> (kLaneWords - k)) : 0u;     // word i <-> bit kLaneWords-1-i
                    p.rd_pre[tf_lo + si] = make_uint2(c_in, xe + e_in);
                }
            }
            __syncwarp();                                       // qlist is rewritten in the next round
        }
        tf_lo = nf_lo;
        tf_hi = nf_hi;
        g_cur = g_nxt;
        __syncwarp();                                           // every lane is done with stage s
        if (lane == 0) {
            const uint64_t nxt = gwid + (uint64_t)(it + kWarpStages) * stride;
            if (nxt < p.n_wt) {
                fence_proxy_async();
                mbar_expect_tx(&full[s], kBoxBytes);
                tma_load_tile(sm.stage[warp][s], &tmap, (uint32_t)nxt * kRowsPerBox, &full[s], l2_policy);
            }
        }
    }
}

// tile_first[t] = first entry of cig_off[0, n_entries) that is >= t * kWarpTileWords (n_entries if none), for
// warp tiles t0..t1: the reads that start inside warp tile t are [tile_first[t], tile_first[t+1]).
// Runs when reads are pushed, not per genotyping pass.
__global__ void __launch_bounds__(256)
k_tile_first(const uint64_t *__restrict__ cig_off, uint64_t n_entries, uint64_t t0, uint64_t t1, uint32_t *__restrict__ tile_first)
{
    const uint64_t t = t0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t > t1) return;
    const uint64_t target = t * kWarpTileWords;
    uint64_t lo = 0, hi = n_entries;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(cig_off + mid) < target) lo = mid + 1; else hi = mid;
    }
    tile_first[t] = (uint32_t)lo;
}

// Per read: (1) the index of its first event in CIGAR order and the stream-wide consumption prefix
// at its first word = warp-tile prefix + the tile-local prefix k_cigar_scan left in rd_pre;
// (2) its events moved from warp-tile storage into CIGAR order with absolute anchors:
// pos1 = ref_start + 1 + bases consumed by the read before the op (the u32 cursor of call.rs:380).
// A warp owns 31 reads; lane 31 only supplies the first-event index of the next read.
// `wt` holds EXCLUSIVE prefixes here (k_exclusive_scan2 ran in between).
__global__ void __launch_bounds__(256)
k_read_fixup(const uint64_t *__restrict__ cig_off, const int32_t *__restrict__ rs, uint64_t R,
             const uint2 *__restrict__ wt, const uint2 *__restrict__ rd_pre,
             const uint32_t *__restrict__ wt_sbase, const uint2 *__restrict__ evraw, uint64_t raw_cap,
             uint2 *__restrict__ events, uint64_t ev_cap, uint32_t *__restrict__ ev_off, DevCounters *__restrict__ ctr)
{
    const uint32_t lane = lane_id();
    const uint64_t warp_global = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t r = warp_global * 31 + lane;
    uint32_t e = 0, c = 0;
    uint64_t g = 0;
    if (r <= R) {
        g = cig_off[r];
        const uint2 w0 = wt[g / kWarpTileWords];
        c = w0.x;
        e = w0.y;
        if (g % kWarpTileWords) {
            const uint2 pre = rd_pre[r];
            c += pre.x;
            e += pre.y;
        }
    }
    const uint32_t e_next = __shfl_down_sync(0xffffffffu, e, 1);
    if (lane == 31 || r > R) return;
    ev_off[r] = e;
    if (r == R) { ctr->n_events = e; return; }
    if (e == e_next) return;
    const uint32_t base = (uint32_t)rs[r] + 1u - c;
    uint64_t gw = g / kWarpTileWords;
    uint2 pre = wt[gw];
    uint32_t hi = wt[gw + 1].y;
    for (uint32_t k = e; k < e_next; ++k) {
        while (k >= hi) { ++gw; pre = wt[gw]; hi = wt[gw + 1].y; }
        const uint64_t slot = (uint64_t)wt_sbase[gw] + (k - pre.y);
        if (k < ev_cap && slot < raw_cap) {
            const uint2 raw = evraw[slot];
            events[k] = make_uint2(base + pre.x + raw.x, raw.y);
        } else {
            atomicOr(&ctr->flags, kFlagEventOverflow);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// exclusive scan of u32 counts -> u32 offsets (n+1 outputs), single pass, decoupled look-back
constexpr int kXsThreads = 256;
constexpr int kXsItems = 8;
constexpr int kXsTile = kXsThreads * kXsItems;

__global__ void __launch_bounds__(kXsThreads)
k_exclusive_scan(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t ntiles,
                 uint64_t *__restrict__ desc, unsigned int *__restrict__ tile_counter, unsigned int *__restrict__ overflow_flags)
{
    __shared__ uint32_t wsum[kXsThreads / 32];
    __shared__ uint32_t tile_s;
    __shared__ uint64_t base_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) tile_s = atomicAdd(tile_counter, 1u);
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
                    if (overflow_flags && base + total > 0xFFFFFFFFull) atomicOr(overflow_flags, kFlagCountOverflow);
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

// same for pairs of u32 (both components wrap mod 2^32; `y` is checked against 2^32): totals `in[0, n)` ->
// exclusive prefixes `out[0, n]`. The CIGAR stream is scanned in ranges of warp tiles (inq_genotype); every
// range continues from the prefix the previous one left in carry_in and leaves its own end in carry_out
// (out[n], the end sentinel of this range, is the value the next range writes to the same slot again).
__global__ void __launch_bounds__(kXsThreads)
k_exclusive_scan2(const uint2 *__restrict__ in, uint2 *__restrict__ out, uint64_t n, uint32_t ntiles,
                  uint64_t *__restrict__ desc_x, uint64_t *__restrict__ desc_y, unsigned int *__restrict__ tile_counter,
                  const unsigned long long *__restrict__ carry_in, unsigned long long *__restrict__ carry_out,
                  unsigned int *__restrict__ overflow_flags)
{
    __shared__ uint32_t wsx[kXsThreads / 32], wsy[kXsThreads / 32];
    __shared__ uint32_t tile_s;
    __shared__ uint64_t bx_s, by_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t cx = carry_in[0], cy = carry_in[1];
    while (true) {
        if (tid == 0) tile_s = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const uint32_t t = tile_s;
        if (t >= ntiles) break;
        const uint64_t i0 = (uint64_t)t * kXsTile + (uint64_t)tid * kXsItems;
        uint2 v[kXsItems];
        uint32_t sx = 0, sy = 0;
#pragma unroll
        for (int k = 0; k < kXsItems; ++k) {
            v[k] = (i0 + k < n) ? in[i0 + k] : make_uint2(0u, 0u);
            sx += v[k].x;
            sy += v[k].y;
        }
        const uint32_t ix = warp_incl_scan(sx), iy = warp_incl_scan(sy);
        if (lane == 31) { wsx[warp] = ix; wsy[warp] = iy; }
        __syncthreads();
        uint32_t wbx = 0, wby = 0, tx = 0, ty = 0;
#pragma unroll
        for (int w = 0; w < kXsThreads / 32; ++w) {
            if (w < (int)warp) { wbx += wsx[w]; wby += wsy[w]; }
            tx += wsx[w];
            ty += wsy[w];
        }
        if (warp == 0) {
            if (lane == 0) {
                st_relaxed_u64(desc_x + t, (t == 0 ? kDescPrefix : kDescAggregate) | (uint64_t)tx);
                st_relaxed_u64(desc_y + t, (t == 0 ? kDescPrefix : kDescAggregate) | (uint64_t)ty);
            }
            uint64_t bx = 0, by = 0;
            if (t > 0) {
                bx = lookback(desc_x, (int64_t)t);
                by = lookback(desc_y, (int64_t)t);
                if (lane == 0) {
                    st_relaxed_u64(desc_x + t, kDescPrefix | ((bx + tx) & 0xFFFFFFFFull));
                    st_relaxed_u64(desc_y + t, kDescPrefix | ((by + ty) & kDescValueMask));
                }
            }
            if (lane == 0) {
                bx_s = bx + cx;
                by_s = by + cy;
                if (t == ntiles - 1) {
                    out[n] = make_uint2((uint32_t)(cx + bx + tx), (uint32_t)(cy + by + ty));
                    carry_out[0] = (cx + bx + tx) & 0xFFFFFFFFull;
                    carry_out[1] = cy + by + ty;
                    if (overflow_flags && cy + by + ty > 0xFFFFFFFFull) atomicOr(overflow_flags, kFlagCountOverflow);
                }
            }
        }
        __syncthreads();
        uint32_t rx = (uint32_t)bx_s + wbx + ix - sx, ry = (uint32_t)by_s + wby + iy - sy;
#pragma unroll
        for (int k = 0; k < kXsItems; ++k) {
            if (i0 + k < n) out[i0 + k] = make_uint2(rx, ry);
            rx += v[k].x;
            ry += v[k].y;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------
// K2b: per (read, locus) candidate: the filter (call.rs:297-300 / 350-352), the sum of the read's
// events anchored inside the locus window (call.rs:388,394,400: start < P && P < end) and the
// scatter of the packed call into the locus' segment: H1 (or every unphased call) grows from the
// front of the segment, H2 from its back; one 64-bit atomic per pair hands out the slot.
// The kernel reads the events straight from the scan kernel's warp-tile storage: event k of the stream
// lives in warp tile t with wt[t].y <= k < wt[t+1].y at slot wt_sbase[t] + (k - wt[t].y), and its 1-based
// anchor (the u32 cursor of call.rs:380) is ref_start + 1 + (wt[t].x + raw.x) - (consumption prefix at the
// read's first word).
#ifndef INQ_PAIR_POOL
#define INQ_PAIR_POOL 256
#endif
#ifndef INQ_PAIR_WARPS
#define INQ_PAIR_WARPS 4
#endif
#ifndef INQ_PAIR_MIN_CTAS
#define INQ_PAIR_MIN_CTAS (40 / INQ_PAIR_WARPS)
#endif
constexpr int kPairWarps = INQ_PAIR_WARPS;  // warps per CTA of k_pair_eval (each warp works alone: no block-level sync)
constexpr int kPairEvPool = INQ_PAIR_POOL;  // events per warp (32 consecutive reads) kept in shared memory
constexpr int kPairLociCache = 128;         // catalog entries per warp kept in shared memory
constexpr int kPairTileCache = 64;          // warp-tile prefixes per warp kept in shared memory

struct EventSource {
    const uint2 *wt;              // [n_wt + 1] exclusive prefixes {consumption, events} per warp tile
    const uint2 *rd_pre;          // [R + 1] tile-local prefixes at each read's first word
    const uint32_t *wt_sbase;     // [n_wt] storage slot of the warp tile's first event
    const uint2 *evraw;           // {consumption inside the warp tile before the op, val}
    uint64_t raw_cap;
};

// stream-wide {consumption, events} prefix at CIGAR word g, which is the first word of read r
__device__ __forceinline__ uint2 read_prefix(const EventSource &es, uint64_t g, uint64_t r)
{
    const uint2 pre = es.rd_pre[r];                          // unconditional: not behind the load of g
    uint2 v = es.wt[g / kWarpTileWords];
    if (g % kWarpTileWords) {                               // (rd_pre is only written for reads starting inside a tile)
        v.x += pre.x;
        v.y += pre.y;
    }
    return v;
}

// event k of a read whose words lie in warp tiles [t_lo, t_hi]: {anchor - base, val} via a search in global memory
__device__ __forceinline__ uint2 event_at(const EventSource &es, uint32_t k, uint32_t t_lo, uint32_t t_hi)
{
    while (t_lo < t_hi) {
        const uint32_t mid = (t_lo + t_hi) >> 1;
        if (es.wt[mid + 1].y > k) t_hi = mid; else t_lo = mid + 1;
    }
    const uint2 w = es.wt[t_lo];
    const uint64_t slot = (uint64_t)es.wt_sbase[t_lo] + (k - w.y);
    uint2 raw = make_uint2(0u, 0u);
    if (slot < es.raw_cap) raw = es.evraw[slot];            // (overflowed speculative buffer: the run is repeated)
    return make_uint2(w.x + raw.x, raw.y);
}

struct PairSmem {                           // per CTA of kPairWarps warps (dynamic shared memory)
    uint2 ev[kPairWarps][kPairEvPool];
    uint32_t off[kPairWarps][32];
    uint32_t ty[kPairWarps][kPairTileCache + 1], tx[kPairWarps][kPairTileCache], tsb[kPairWarps][kPairTileCache];
    int32_t ls[kPairWarps][kPairLociCache], le[kPairWarps][kPairLociCache];
    uint32_t seg[kPairWarps][kPairLociCache + 1];
    uint32_t joined[kPairWarps];
};

// reads [r_begin, r_end): the CIGAR stream is scanned in ranges of warp tiles and every range's reads are
// evaluated as soon as their last word has been scanned (inq_genotype)
__global__ void __launch_bounds__(kPairWarps * 32, INQ_PAIR_MIN_CTAS)
k_pair_eval(ReadView rv, uint64_t r_begin, uint64_t r_end, LocusView lv, int unphased, const uint32_t *__restrict__ cand_lo,
            const uint32_t *__restrict__ cand_n, EventSource es, const uint32_t *__restrict__ seg_off,
            unsigned long long *__restrict__ cursor, uint64_t *__restrict__ vals, uint64_t vals_cap,
            DevCounters *__restrict__ ctr, uint32_t debug)
{
    // A warp owns 32 consecutive reads; their candidates are flattened and dealt to the lanes
    // 32 at a time (reads have 0..hundreds of candidates, a per-read loop leaves most lanes idle).
    // Everything the inner loop needs is staged in shared memory first (the reads' events -- one
    // contiguous run of the event stream, rebased from tile-local to stream-wide consumption on the way
    // in: a read's anchors are its abase + that --, the slice of
    // the catalog the warp touches), so that the only global operations left in the loop are the slot
    // atomic and the store of the call -- and the store is deferred by one iteration so that the atomic's
    // round trip overlaps the next candidate.
    extern __shared__ __align__(16) unsigned char pair_smem_raw[];
    PairSmem &sm = *reinterpret_cast<PairSmem *>(pair_smem_raw);
    auto &s_off = sm.off; auto &s_joined = sm.joined; auto &s_ev = sm.ev;
    auto &s_ty = sm.ty; auto &s_tx = sm.tx; auto &s_tsb = sm.tsb; auto &s_ls = sm.ls; auto &s_le = sm.le; auto &s_seg = sm.seg;
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    const uint64_t r = r_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t n = 0, lo = 0, e0 = 0xFFFFFFFFu, e1 = 0xFFFFFFFFu, hf = 0, words = 0, t_lo = 0, t_hi = 0, abase = 0;
    int32_t rs = 0, re = 0;
    if (r < r_end) {
        // independent loads, all in flight together
        n = cand_n[r];
        lo = cand_lo[r];
        rs = rv.rs[r];
        re = rv.re[r];
        hf = (uint32_t)rv.hp[r] | ((uint32_t)(rv.flags[r] & 1u) << 8) | ((uint32_t)(rv.flags[r] & 2u) << 9);   // bit 10: SA panic
        const uint64_t g0 = rv.cig_off[r], g1 = rv.cig_off[r + 1];
        const uint2 p0 = read_prefix(es, g0, r), p1 = read_prefix(es, g1, r + 1);
        e0 = p0.y;
        e1 = p1.y;
        abase = (uint32_t)rs + 1u - p0.x;                       // anchor = abase + stream-wide consumption before the op
        t_lo = (uint32_t)(g0 / kWarpTileWords);
        t_hi = g1 > g0 ? (uint32_t)((g1 - 1) / kWarpTileWords) : t_lo;
        words = (uint32_t)min(g1 - g0, (uint64_t)0xFFFFFFFFu);
    }
    // the events of the warp's joined reads are one run [ev_lo, ev_hi) of the event stream, stored in the
    // warp tiles [tile_lo, tile_hi]; the tile range only depends on cig_off, so its prefixes are fetched
    // while the per-read prefixes are still in flight
    const uint32_t tile_lo = __reduce_min_sync(0xffffffffu, n ? t_lo : 0xFFFFFFFFu);
    const uint32_t tile_hi = __reduce_max_sync(0xffffffffu, n ? t_hi : 0u);
    const uint32_t nt = tile_hi - tile_lo + 1u;
    const bool tiles_cached = tile_lo <= tile_hi && nt <= (uint32_t)kPairTileCache;
    if (tiles_cached) {
        for (uint32_t i = lane; i <= nt; i += 32) {
            const uint2 w = es.wt[tile_lo + i];
            s_ty[wid][i] = w.y;
            if (i < nt) {
                s_tx[wid][i] = w.x;
                s_tsb[wid][i] = es.wt_sbase[tile_lo + i];
            }
        }
    }
    const uint32_t ev_lo = __reduce_min_sync(0xffffffffu, n ? e0 : 0xFFFFFFFFu);
    const uint32_t ev_hi = __reduce_max_sync(0xffffffffu, n ? e1 : 0u);
    __syncwarp();
    if (ev_lo < ev_hi) {
        const uint32_t cnt = min(ev_hi - ev_lo, (uint32_t)kPairEvPool);
        // pass 1: where each event lives (storage slot) and the consumption prefix of its tile -- parked in s_ev
        // (enumerating the events tile by tile instead was measured: fewer instructions, but 14 dependent shared-memory
        // round trips in a row per warp -- 0.57 ms instead of 0.52 for the kernel)
        uint32_t ta = 0;                                        // cached tiles: this lane's k only grows, so does its tile
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint32_t k = ev_lo + i;
            uint32_t slot, rel;                                 // rel: stream-wide consumption before the tile
            if (tiles_cached) {
                // first tile t with ty[t + 1] > k: 32 events further on is typically 2-3 tiles further on
                while (ta + 1u < nt && s_ty[wid][ta + 1u] <= k) ++ta;
                slot = s_tsb[wid][ta] + (k - s_ty[wid][ta]);
                rel = s_tx[wid][ta];
            } else {
                uint32_t a = tile_lo, b = tile_hi;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    if (es.wt[mid + 1].y > k) b = mid; else a = mid + 1;
                }
                const uint2 w = es.wt[a];
                slot = es.wt_sbase[a] + (k - w.y);
                rel = w.x;
            }
            s_ev[wid][i] = make_uint2(slot, rel);
        }
        // pass 2: independent loads from the event storage (several in flight per lane)
#pragma unroll 4
        for (uint32_t i = lane; i < cnt; i += 32) {
            const uint2 sr = s_ev[wid][i];
            uint2 raw = make_uint2(0u, 0u);
            if ((uint64_t)sr.x < es.raw_cap) raw = es.evraw[sr.x];   // (overflowed speculative buffer: the run is repeated)
            s_ev[wid][i] = make_uint2(sr.y + raw.x, raw.y);
        }
    }
    if (n && e1 - ev_lo > (uint32_t)kPairEvPool) hf |= 1u << 9;      // not (entirely) staged: searched in global memory
    const uint32_t eb = e0 - ev_lo;                                   // first staged event of the read
    const uint32_t incl = warp_incl_scan(n);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    s_off[wid][lane] = incl - n;
    if (lane == 0) s_joined[wid] = 0u;
    // slice of the catalog this warp touches
    const uint32_t lo_min = __reduce_min_sync(0xffffffffu, n ? lo : 0xFFFFFFFFu);
    const uint32_t hi_max = __reduce_max_sync(0xffffffffu, n ? lo + n : 0u);
    const bool cached = total && (hi_max - lo_min <= (uint32_t)kPairLociCache);
    if (cached) {
        const uint32_t span = hi_max - lo_min;
        for (uint32_t i = lane; i < span; i += 32) {
            s_ls[wid][i] = __ldg(lv.start + lo_min + i);
            s_le[wid][i] = __ldg(lv.end + lo_min + i);
        }
        for (uint32_t i = lane; i <= span; i += 32) s_seg[wid][i] = __ldg(seg_off + lo_min + i);
    }
    __syncwarp();

    uint32_t npass = 0;                 // pairs that land in a bucket
    uint64_t visits = 0;                // CIGAR words the reference walks: every passing pair, incl. HP 0
    bool bad_hp = false, bad_sa = false;
    // deferred store of the previous candidate
    bool pend = false;
    unsigned long long pend_old = 0;
    uint64_t pend_key = 0;
    uint32_t pend_seg = 0, pend_cap = 0;
    bool pend_back = false;
    auto flush_pending = [&]() {
        if (!pend) return;
        const uint32_t slot_k = pend_back ? (uint32_t)(pend_old >> 32) : (uint32_t)pend_old;
        const uint64_t slot = pend_back ? (uint64_t)pend_seg + (pend_cap - 1u - slot_k) : (uint64_t)pend_seg + slot_k;
        if (INQ_DBG(debug) & 2u) {
        } else if (slot_k < pend_cap && slot < vals_cap) vals[slot] = pend_key;
        else atomicOr(&ctr->flags, kFlagValsOverflow);
        pend = false;
    };

    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t k = base + lane;
        const bool active = k < total;
        // owning read: the last j with s_off[j] <= k
        uint32_t j = 0;
#pragma unroll
        for (uint32_t step = 16; step >= 1; step >>= 1)
            if (active && s_off[wid][j + step] <= k) j += step;
        const uint32_t idx = k - s_off[wid][j];
        const int32_t rs_j = __shfl_sync(0xffffffffu, rs, j), re_j = __shfl_sync(0xffffffffu, re, j);
        const uint32_t hf_j = __shfl_sync(0xffffffffu, hf, j), lo_j = __shfl_sync(0xffffffffu, lo, j);
        const uint32_t e0_j = __shfl_sync(0xffffffffu, e0, j), e1_j = __shfl_sync(0xffffffffu, e1, j);
        const uint32_t eb_j = __shfl_sync(0xffffffffu, eb, j);
        const uint32_t tlo_j = __shfl_sync(0xffffffffu, t_lo, j), thi_j = __shfl_sync(0xffffffffu, t_hi, j);
        const uint32_t abase_j = __shfl_sync(0xffffffffu, abase, j);
        const uint32_t words_j = __shfl_sync(0xffffffffu, words, j);
        bool emit = false;
        uint32_t l = 0, h = 0, seg = 0, cap = 0;
        int32_t ls = 0, le = 0;
        if (active) {
            l = lo_j + idx;
            h = hf_j & 0xFFu;
            if (cached) {
                ls = s_ls[wid][l - lo_min];
                le = s_le[wid][l - lo_min];
                seg = s_seg[wid][l - lo_min];
                cap = s_seg[wid][l - lo_min + 1] - seg;
            } else {
                ls = __ldg(lv.start + l);
                le = __ldg(lv.end + l);
                seg = __ldg(seg_off + l);
                cap = __ldg(seg_off + l + 1) - seg;
            }
            if (pair_passes(unphased != 0, rs_j, re_j, ls, le)) {
                visits += words_j;                          // call.rs:357 runs before the bucket lookup
                emit = true;
                if (hf_j & (1u << 10)) {                    // call.rs:394 -> 431: the walk panics inside is_accidental_2d
                    bad_sa = true;
                    ctr->bad_sa_read = r - lane + j;
                    emit = false;
                } else if (!unphased) {
                    if (h > 2u) {                           // call.rs:358 unwrap on None
                        bad_hp = true;
                        ctr->bad_hp_value = h;
                        ctr->bad_hp_read = r - lane + j;
                        emit = false;
                    } else if (h == 0u) {
                        emit = false;                       // HP 0 lands in the ignored bucket
                    }
                }
            }
        }
        unsigned long long old = 0;
        uint64_t key = 0;
        const bool back = !unphased && h == 2u;
        if (emit) {
            ++npass;
            atomicOr(&s_joined[wid], 1u << j);
            // slot first: the atomic's round trip overlaps the rest of this iteration and the next one
            old = (INQ_DBG(debug) & 1u) ? 0ull : atomicAdd(cursor + l, back ? (1ull << 32) : 1ull);
            const uint32_t start_ext = (uint32_t)ls - 10u, end_ext = (uint32_t)le + 10u;
            const bool is2d = ((hf_j >> 8) & 1u) != 0u;
            int64_t call = 0;
            uint32_t clip = 0;
            if (INQ_DBG(debug) & 4u) {
            } else if (!(hf_j & (1u << 9))) {
                // the read's events sit in shared memory, sorted by anchor: binary search, then a short scan
                const uint32_t ne = e1_j - e0_j;
                const uint2 *sev = &s_ev[wid][eb_j];
                uint32_t a = 0, b = ne;
                while (a < b) {
                    const uint32_t m = (a + b) >> 1;
                    if (abase_j + sev[m].x > start_ext) b = m; else a = m + 1;
                }
                for (uint32_t q = a; q < ne; ++q) {
                    const uint2 ev = sev[q];
                    if (!(abase_j + ev.x < end_ext)) break; // start_ext < P && P < end_ext (call.rs:388,394,400)
                    const int32_t v = (int32_t)ev.y;
                    const uint32_t is_s = (uint32_t)v & 1u;
                    if (is_s && is2d) continue;             // call.rs:394 !is_accidental_2d(&r)
                    call += (int64_t)(v >> 1);
                    clip |= is_s;
                }
            } else {
                // long event list, read from the warp-tile storage: anchors increase along the read, so a
                // binary search finds the first event with pos1 > start_ext; then a short scan
                uint32_t a = e0_j, b = e1_j;
                while (a < b) {
                    const uint32_t m = (a + b) >> 1;
                    if (abase_j + event_at(es, m, tlo_j, thi_j).x > start_ext) b = m; else a = m + 1;
                }
                for (uint32_t e = a; e < e1_j; ++e) {
                    const uint2 ev = event_at(es, e, tlo_j, thi_j);
                    if (!(abase_j + ev.x < end_ext)) break;
                    const int32_t v = (int32_t)ev.y;
                    const uint32_t is_s = (uint32_t)v & 1u;
                    if (is_s && is2d) continue;             // call.rs:394 !is_accidental_2d(&r)
                    call += (int64_t)(v >> 1);
                    clip |= is_s;
                }
            }
            key = ((uint64_t)(call + kCallBias) << 1) | clip;
        }
        flush_pending();                                    // store of the previous candidate
        if (emit) {
            pend = true;
            pend_old = old;
            pend_key = key;
            pend_seg = seg;
            pend_cap = cap;
            pend_back = back;
        }
    }
    flush_pending();
    __syncwarp();
    // statistics (one atomic per warp per counter)
    const bool joined = ((s_joined[wid] >> lane) & 1u) != 0u;
    const uint32_t pass_w = __reduce_add_sync(0xffffffffu, npass);
    const uint32_t join_w = __popc(__ballot_sync(0xffffffffu, joined));
    const uint64_t words_w = warp_sum(joined ? (uint64_t)words : 0ull);
    const uint64_t visits_w = warp_sum(visits);
    const uint32_t bad_w = __any_sync(0xffffffffu, bad_hp), bad_sa_w = __any_sync(0xffffffffu, bad_sa);
    if (lane == 0) {
        // one of kStatSlots copies per warp: no block barrier, and no single hot address in L2
        unsigned long long *slot = ctr->stat[(blockIdx.x * (uint32_t)kPairWarps + wid) % kStatSlots];
        if (pass_w) atomicAdd(slot + ST_PAIRS, (unsigned long long)pass_w);
        if (join_w) atomicAdd(slot + ST_READS_JOINED, (unsigned long long)join_w);
        if (words_w) atomicAdd(slot + ST_WORDS_JOINED, (unsigned long long)words_w);
        if (visits_w) atomicAdd(slot + ST_OP_VISITS, (unsigned long long)visits_w);
        if (bad_w) atomicOr(&ctr->flags, kFlagBadHp);
        if (bad_sa_w) atomicOr(&ctr->flags, kFlagBadSa);
    }
}

// ----------------------------------------------------------------------------------------------
// K3: per-locus sort / split / support filter / median (call.rs:308-321,365-369,497-522).
// Keys are 63-bit ((call + bias) << 1 | clip): ascending key order == (value, Span before Clip).
// Phased loci set bit 63 on H2 keys so one sort orders [H1 | H2]; unphased loci split the sorted
// run at n/2 (call.rs:314).

// Key traits: 64-bit keys are ((call + 2^61) << 1) | clip with the H2 marker in bit 63; when every
// call of a locus fits 30 bits the same layout is used in 32 bits (bias 2^29, H2 marker bit 31),
// which halves the shuffles and compares of the sorting network.
template <typename T> struct KeyTraits;
template <> struct KeyTraits<uint64_t> {
    static constexpr uint64_t hap = kKeyHapBit, inf = kKeyInf;
    static __device__ __forceinline__ int64_t call(uint64_t k) { return (int64_t)((k & ~hap) >> 1) - kCallBias; }
};
template <> struct KeyTraits<uint32_t> {
    static constexpr uint32_t hap = 1u << 31, inf = 0xFFFFFFFFu;
    static constexpr int32_t bias = 1 << 29;
    static __device__ __forceinline__ int64_t call(uint32_t k) { return (int64_t)((int32_t)((k & ~hap) >> 1) - bias); }
};

// all-ascending bitonic network over K striped registers per lane (element i = k*32 + lane)
// N < K * 32 sorts every aligned run of N elements on its own
template <int K, typename T, int N = K * 32>
__device__ __forceinline__ void warp_sort(T (&key)[K])
{
#pragma unroll
    for (int blk = 2; blk <= N; blk <<= 1) {
        // mirror step: partner = i ^ (blk - 1)
        {
            const int lane_x = (blk - 1) & 31, reg_x = (blk - 1) >> 5;
            T other[K];
#pragma unroll
            for (int k = 0; k < K; ++k) other[k] = __shfl_xor_sync(0xffffffffu, key[k ^ reg_x], lane_x);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = k * 32 + (int)lane_id();
                const bool lower = (i & (blk >> 1)) == 0;      // lower half of the mirrored block keeps the min
                const T a = key[k], b = other[k];
                key[k] = lower ? (a < b ? a : b) : (a > b ? a : b);
            }
        }
#pragma unroll
        for (int d = blk >> 2; d >= 1; d >>= 1) {
            const int lane_x = d & 31, reg_x = d >> 5;
            T other[K];
#pragma unroll
            for (int k = 0; k < K; ++k)
                other[k] = lane_x ? __shfl_xor_sync(0xffffffffu, key[k ^ reg_x], lane_x) : key[k ^ reg_x];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = k * 32 + (int)lane_id();
                const bool lower = (i & d) == 0;
                const T a = key[k], b = other[k];
                key[k] = lower ? (a < b ? a : b) : (a > b ? a : b);
            }
        }
    }
}

__device__ __forceinline__ int64_t key_call(uint64_t key) { return KeyTraits<uint64_t>::call(key); }

// median of the sorted sub-range [a, a+n) held striped across the warp. Returns validity;
// *twice = 2 x median. call.rs:497-522.
template <int K, typename T>
__device__ __forceinline__ bool warp_median_part(const T (&key)[K], uint32_t a, uint32_t n, uint32_t support,
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
        span_b[k] = __ballot_sync(0xffffffffu, in && !(key[k] & 1u));
        clip_b[k] = __ballot_sync(0xffffffffu, in && (key[k] & 1u));
        s += __popc(span_b[k]);
        c += __popc(clip_b[k]);
    }
    const uint32_t topk = (s <= support) ? (support - s) : 0u;       // call.rs:509-513
    const uint32_t m = s + topk;
    if (m == 0) { *panicked = true; return false; }                  // call.rs:516 on an empty vector
    const uint32_t t2 = m >> 1, t1 = (m & 1u) ? t2 : t2 - 1u;        // call.rs:515-521
    uint32_t clip_before = 0, sel_before = 0;
    int64_t sum = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t me = 1u << lane_id();
        const bool is_span = (span_b[k] & me) != 0, is_clip = (clip_b[k] & me) != 0;
        const uint32_t clip_idx = clip_before + __popc(clip_b[k] & lanemask_lt());
        const bool sel = is_span || (is_clip && clip_idx + topk >= c);   // the topk largest clips
        const uint32_t sel_b = __ballot_sync(0xffffffffu, sel);
        const uint32_t rank = sel_before + __popc(sel_b & lanemask_lt());
        if (sel) {
            const int64_t v = KeyTraits<T>::call(key[k]);
            if (rank == t1) sum += v;
            if (rank == t2) sum += v;
        }
        clip_before += __popc(clip_b[k]);
        sel_before += __popc(sel_b);
    }
    // (handing the two middle keys out by ballot + shuffle instead was measured: 15 % slower)
    *twice = warp_sum(sum);
    return true;
}

// A locus' segment [seg, seg+cap) holds its H1 calls at the front and its H2 calls at the back
// (unphased: everything at the front). nf/nb = number of front/back entries; n1 = split point of
// the sorted run (phased: nf, unphased: (nf)/2).
// Both haplotypes of a phased locus in one pass (call.rs:497-522 twice): 32-bit keys, H1 sorted in lanes
// 0-15 (nf of them), H2 sorted in lanes 16-31 (nb of them). Every lane works on its own half: the ballots
// are shared, the counts are taken under the half's lane mask.
__device__ __forceinline__ void median_halves32(uint32_t key, uint32_t nf, uint32_t nb, uint32_t support, int64_t *t1,
                                                int64_t *t2, uint32_t *valid, bool *panicked)
{
    const uint32_t lane = lane_id();
    const bool upper = lane >= 16u;
    const uint32_t hm = upper ? 0xFFFF0000u : 0x0000FFFFu;
    const uint32_t n = upper ? nb : nf;
    const bool in = (lane & 15u) < n;
    const uint32_t span_b = __ballot_sync(0xffffffffu, in && !(key & 1u));
    const uint32_t clip_b = __ballot_sync(0xffffffffu, in && (key & 1u));
    const uint32_t s = __popc(span_b & hm), c = __popc(clip_b & hm);
    const bool enough = n >= support;                                // call.rs:498-500
    const uint32_t topk = (s <= support) ? (support - s) : 0u;       // call.rs:509-513
    const uint32_t m = s + topk;
    const bool ok = enough && m != 0u;                               // m == 0: call.rs:516 on an empty vector
    const uint32_t i2 = m >> 1, i1 = (m & 1u) ? i2 : i2 - 1u;        // call.rs:515-521
    const uint32_t below = hm & lanemask_lt();
    const bool is_span = (span_b >> lane) & 1u, is_clip = (clip_b >> lane) & 1u;
    const bool sel = ok && (is_span || (is_clip && __popc(clip_b & below) + topk >= c));   // the topk largest clips
    const uint32_t sel_b = __ballot_sync(0xffffffffu, sel);
    const uint32_t rank = __popc(sel_b & below);
    int32_t sum = 0;                                                 // |call| < 2^29: two of them fit
    if (sel) {
        const int32_t v = (int32_t)KeyTraits<uint32_t>::call(key);
        sum = (rank == i1 ? v : 0) + (rank == i2 ? v : 0);
    }
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const uint32_t ok_b = __ballot_sync(0xffffffffu, ok), bad_b = __ballot_sync(0xffffffffu, enough && m == 0u);
    *t1 = (ok_b & 1u) ? (int64_t)__shfl_sync(0xffffffffu, sum, 0) : 0;
    *t2 = (ok_b & 0x10000u) ? (int64_t)__shfl_sync(0xffffffffu, sum, 16) : 0;
    *valid = ((ok_b & 1u) ? 1u : 0u) | ((ok_b & 0x10000u) ? 2u : 0u);
    if (bad_b & 0x10001u) *panicked = true;
}

// kSplit > 0 (phased only): H1 lives at positions [0, nf) and H2 at [kSplit, kSplit + nb) and the two
// runs of kSplit elements are sorted independently (a shorter network than one sort over both).
template <int K, int kSplit = 0>
__device__ __forceinline__ void warp_locus(const uint64_t *__restrict__ vals, uint32_t seg, uint32_t cap, uint32_t nf,
                                           uint32_t nb, uint32_t n1, bool phased, uint32_t support, int64_t *t1,
                                           int64_t *t2, uint32_t *valid, bool *panicked)
{
    const uint32_t ntot = nf + nb;
    constexpr int kSortN = kSplit ? kSplit : K * 32;
    const uint32_t a2 = kSplit ? (uint32_t)kSplit : n1;              // first position of the second part
    uint64_t key[K];
    bool have_k[K];
    bool small = true;                      // every call fits the 32-bit key layout
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t i = k * 32 + lane_id();
        uint64_t v = kKeyInf;
        bool have;
        if constexpr (kSplit > 0) {
            have = (i < nf) || (i >= (uint32_t)kSplit && i - (uint32_t)kSplit < nb);
            if (i < nf) v = vals[seg + i];
            else if (have) v = vals[seg + cap - nb + (i - (uint32_t)kSplit)];
        } else {
            have = i < ntot;
            if (i < nf) v = vals[seg + i];
            else if (have) v = vals[seg + cap - nb + (i - nf)] | (phased ? kKeyHapBit : 0ull);
        }
        if (have) {
            const int64_t c = key_call(v);
            small = small && c >= -(int64_t)KeyTraits<uint32_t>::bias && c < (int64_t)KeyTraits<uint32_t>::bias;
        }
        key[k] = v;
        have_k[k] = have;
    }
    bool v1, v2;
    if (__all_sync(0xffffffffu, small)) {
        uint32_t k32[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint64_t v = key[k];
            const uint32_t body = (uint32_t)((key_call(v) + KeyTraits<uint32_t>::bias) << 1) | (uint32_t)(v & 1ull);
            k32[k] = have_k[k] ? (body | ((v & kKeyHapBit) ? KeyTraits<uint32_t>::hap : 0u)) : KeyTraits<uint32_t>::inf;
        }
        warp_sort<K, uint32_t, kSortN>(k32);
        if constexpr (K == 1 && kSplit == 16) {
            median_halves32(k32[0], nf, nb, support, t1, t2, valid, panicked);
            return;
        }
        v1 = warp_median_part<K, uint32_t>(k32, 0, n1, support, t1, panicked);
        v2 = warp_median_part<K, uint32_t>(k32, a2, ntot - n1, support, t2, panicked);
    } else {
        warp_sort<K, uint64_t, kSortN>(key);
        v1 = warp_median_part<K, uint64_t>(key, 0, n1, support, t1, panicked);
        v2 = warp_median_part<K, uint64_t>(key, a2, ntot - n1, support, t2, panicked);
    }
    *valid = (v1 ? 1u : 0u) | (v2 ? 2u : 0u);
}

constexpr int kMedianWarpMax = 128;         // loci with more calls go to the CTA kernel

#ifndef INQ_MEDIAN_LOCI_PER_WARP
#define INQ_MEDIAN_LOCI_PER_WARP 1     // measured: 1 -> 0.357 ms, 2 -> 0.369-0.389, 3 -> 0.473 (the kernel is issue-bound, not latency-bound)
#endif
#ifndef INQ_MEDIAN_MIN_CTAS
#define INQ_MEDIAN_MIN_CTAS (INQ_MEDIAN_LOCI_PER_WARP == 1 ? 8 : 6)
#endif
constexpr int kMedianLociPerWarp = INQ_MEDIAN_LOCI_PER_WARP;

// header of one locus' segment
struct LocusSeg {
    uint32_t seg, cap, nf, nb;
    bool ok;
};
__device__ __forceinline__ LocusSeg locus_seg(uint32_t seg, uint32_t cap, unsigned long long cur, uint64_t vals_cap)
{
    LocusSeg h;
    h.seg = seg;
    h.cap = cap;
    h.ok = (uint64_t)seg + cap <= vals_cap;
    h.nf = min((uint32_t)cur, cap);
    h.nb = min((uint32_t)(cur >> 32), cap - h.nf);
    return h;
}
// the common phased case (both haplotypes <= 16 calls): lane i < 16 holds H1's call i, lane 16 + i H2's call i
__device__ __forceinline__ uint64_t fast16_load(const uint64_t *__restrict__ vals, const LocusSeg &h)
{
    const uint32_t i = lane_id();
    if (i < h.nf) return vals[h.seg + i];
    if (i >= 16u && i - 16u < h.nb) return vals[h.seg + h.cap - h.nb + (i - 16u)];
    return kKeyInf;
}

// one locus: everything after the header. `pre`: the keys of the fast path, already loaded (kKeyInf lanes are empty).
__device__ __forceinline__ void median_one(uint32_t l, uint32_t l0, int chunk, int unphased, uint32_t support, const LocusSeg &h, bool fast, uint64_t pre,
                                           const uint64_t *__restrict__ vals, int64_t *__restrict__ twice_h1, int64_t *__restrict__ twice_h2,
                                           uint8_t *__restrict__ valid, uint32_t *__restrict__ big_list, DevCounters *__restrict__ ctr)
{
    if (!h.ok) {                                        // speculatively sized call buffer too small: the run is repeated
        if (lane_id() == 0) atomicOr(&ctr->flags, kFlagValsOverflow);
        return;
    }
    const uint32_t nf = h.nf, nb = h.nb, seg = h.seg, cap = h.cap;
    const uint32_t ntot = nf + nb;
    const uint32_t n1 = unphased ? (ntot >> 1) : nf;               // call.rs:314 split_at(len/2)
    int64_t t1 = 0, t2 = 0;
    uint32_t vm = 0;
    bool panicked = false;
    bool done = false;
    if (fast) {
        // keys are here already: 32-bit layout when every call fits 30 bits (practically always)
        const bool have = pre != kKeyInf;
        const int64_t c = key_call(pre);
        const bool small = !have || (c >= -(int64_t)KeyTraits<uint32_t>::bias && c < (int64_t)KeyTraits<uint32_t>::bias);
        if (__all_sync(0xffffffffu, small)) {
            uint32_t k32[1];
            k32[0] = have ? ((uint32_t)((c + KeyTraits<uint32_t>::bias) << 1) | (uint32_t)(pre & 1ull)) : KeyTraits<uint32_t>::inf;
            warp_sort<1, uint32_t, 16>(k32);
            median_halves32(k32[0], nf, nb, support, &t1, &t2, &vm, &panicked);
            done = true;
        }
    }
    if (!done) {
        if (!unphased && nf <= 16 && nb <= 16) warp_locus<1, 16>(vals, seg, cap, nf, nb, n1, true, support, &t1, &t2, &vm, &panicked);
        else if (ntot <= 32) warp_locus<1>(vals, seg, cap, nf, nb, n1, !unphased, support, &t1, &t2, &vm, &panicked);
        else if (!unphased && nf <= 32 && nb <= 32) warp_locus<2, 32>(vals, seg, cap, nf, nb, n1, true, support, &t1, &t2, &vm, &panicked);
        else if (ntot <= 64) warp_locus<2>(vals, seg, cap, nf, nb, n1, !unphased, support, &t1, &t2, &vm, &panicked);
        else if (ntot <= kMedianWarpMax) warp_locus<4>(vals, seg, cap, nf, nb, n1, !unphased, support, &t1, &t2, &vm, &panicked);
        else {
            if (lane_id() == 0) big_list[l0 + atomicAdd(&ctr->big_count[chunk], 1u)] = l;
            return;
        }
    }
    if (lane_id() == 0) {
        twice_h1[l] = t1;
        twice_h2[l] = t2;
        valid[l] = (uint8_t)vm;
        if (panicked) atomicOr(&ctr->flags, kFlagMedianEmpty);
    }
}

// A warp owns kMedianLociPerWarp consecutive loci (default 1). Each locus costs two dependent round trips (segment header,
// then its calls); with several loci per warp all headers and then all key loads are in flight together before any locus is
// reduced -- measured, it does not pay: 68 % of the issue slots are busy already and the extra registers cost occupancy.
__global__ void __launch_bounds__(256, INQ_MEDIAN_MIN_CTAS)
k_locus_median(uint32_t l0, uint32_t l1, int chunk, int unphased, uint32_t support, const uint32_t *__restrict__ seg_off,
               const unsigned long long *__restrict__ cursor, const uint64_t *__restrict__ vals, uint64_t vals_cap,
               int64_t *__restrict__ twice_h1, int64_t *__restrict__ twice_h2, uint8_t *__restrict__ valid,
               uint32_t *__restrict__ big_list, DevCounters *__restrict__ ctr)
{
    // loci [l0, l1) of the catalog (chunk < kMaxMedianChunks); the chunk's CTA-path loci are listed in big_list[l0 ...]
    const uint32_t lw = l0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (uint32_t)kMedianLociPerWarp;
    if (lw >= l1) return;
    LocusSeg h[kMedianLociPerWarp];
    bool live[kMedianLociPerWarp], fast[kMedianLociPerWarp];
    uint64_t pre[kMedianLociPerWarp];
    // round trip 1: the headers of all the warp's loci
    {
        uint32_t so[kMedianLociPerWarp + 1];
        unsigned long long cur[kMedianLociPerWarp];
#pragma unroll
        for (int i = 0; i <= kMedianLociPerWarp; ++i) so[i] = (lw + i <= l1) ? seg_off[lw + i] : 0u;
#pragma unroll
        for (int i = 0; i < kMedianLociPerWarp; ++i) {
            live[i] = lw + i < l1;
            cur[i] = live[i] ? cursor[lw + i] : 0ull;
        }
#pragma unroll
        for (int i = 0; i < kMedianLociPerWarp; ++i) h[i] = locus_seg(so[i], live[i] ? so[i + 1] - so[i] : 0u, cur[i], vals_cap);
    }
    // round trip 2: the calls of the loci on the common path
#pragma unroll
    for (int i = 0; i < kMedianLociPerWarp; ++i) {
        fast[i] = live[i] && h[i].ok && !unphased && h[i].nf <= 16u && h[i].nb <= 16u;
        pre[i] = fast[i] ? fast16_load(vals, h[i]) : kKeyInf;
    }
#pragma unroll
    for (int i = 0; i < kMedianLociPerWarp; ++i)
        if (live[i]) median_one(lw + i, l0, chunk, unphased, support, h[i], fast[i], pre[i], vals, twice_h1, twice_h2, valid, big_list, ctr);
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
k_locus_median_big(uint32_t l0, int chunk, int unphased, uint32_t support, const uint32_t *__restrict__ seg_off,
                   const unsigned long long *__restrict__ cursor, uint64_t *__restrict__ vals, uint64_t vals_cap,
                   int64_t *__restrict__ twice_h1, int64_t *__restrict__ twice_h2, uint8_t *__restrict__ valid,
                   const uint32_t *__restrict__ big_list, DevCounters *__restrict__ ctr)
{
    __shared__ uint64_t skeys[kBigSmemKeys];
    __shared__ uint32_t item_s;
    __shared__ uint32_t sh_cnt[4];
    __shared__ unsigned long long sh_acc;
    const uint32_t tid = threadIdx.x;
    while (true) {
        if (tid == 0) item_s = atomicAdd(&ctr->big_cursor[chunk], 1u);
        __syncthreads();
        const uint32_t item = item_s;
        if (item >= ctr->big_count[chunk]) break;
        const uint32_t l = big_list[l0 + item];
        const uint32_t seg = seg_off[l], cap = seg_off[l + 1] - seg;
        if ((uint64_t)seg + cap > vals_cap) { __syncthreads(); continue; }     // see k_locus_median
        const unsigned long long cur = cursor[l];
        const uint32_t nf = min((uint32_t)cur, cap), nb = min((uint32_t)(cur >> 32), cap - nf);
        uint32_t ntot = nf + nb;
        const uint32_t n1 = unphased ? (ntot >> 1) : nf;
        const uint32_t n2 = ntot - n1;
        volatile uint64_t *a;
        if (ntot <= (uint32_t)kBigSmemKeys) {
            for (uint32_t i = tid; i < ntot; i += blockDim.x) {
                uint64_t v;
                if (i < nf) v = vals[seg + i];
                else v = vals[seg + cap - nb + (i - nf)] | (unphased ? 0ull : kKeyHapBit);
                skeys[i] = v;
            }
            a = skeys;
        } else {
            // sort the whole segment in place; the unused gap between front and back becomes +inf keys
            for (uint32_t i = nf + tid; i < cap; i += blockDim.x) {
                if (i < cap - nb) vals[seg + i] = kKeyInf;
                else if (!unphased) vals[seg + i] |= kKeyHapBit;
            }
            a = vals + seg;
            ntot = cap;
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
        block_part_median(a, n1, n2, support, &t2, &ok2, &panicked, sh_cnt, &sh_acc);
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
