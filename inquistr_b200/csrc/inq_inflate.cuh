// inq_inflate.cuh -- prototype: BGZF (raw DEFLATE, RFC 1951) inflate on the GPU, sm_100a.
// SURVEY 8f rank 1 names "GPU inflate later" as the step after the host decoder (csrc/host/inflate_fast.hpp): the
// reference's wall time lives in htslib's per-locus BGZF inflate (call.rs:288,294,338) and the host side of this
// repository is inflate-bound too. This kernel is the device half of that step, measured on its own
// (tools/bench_gpu_inflate.py); the product CLI does not use it yet (records would have to be parsed on the device).
//
// One warp per BGZF block (<= 64 KB of output). Huffman decoding is serial per block, so the warp runs it in
// lock-step -- every lane holds the same bit buffer and takes the same branches, nothing diverges -- and uses its
// width where DEFLATE allows it:
//   * input: the lanes hold 256 bytes of the compressed stream in registers (8 bytes each, next chunk prefetched);
//     a refill is one shuffle, no memory latency on the decode chain;
//   * tables: built by all lanes (7.3 KB of shared memory per warp: 10-bit literal/length and 9-bit distance
//     primary tables with fixed-size second-level tables, 16-bit entries), looked up by a broadcast LDS;
//   * matches: copied by all lanes, byte k of the match from out[pos - dist + k % dist] (the source never overlaps
//     what this copy writes); literals are gathered four at a time into one store.
// Anything unexpected (invalid code, output overrun, too many long codes for the second-level tables) sets the
// block's status; the caller re-inflates such blocks on the host.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace inqz {

constexpr int kLitBits = 10, kLitSubBits = 5, kLitSubTables = 32;      // 15-bit codes: 10 + 5
constexpr int kDistBits = 9, kDistSubBits = 6, kDistSubTables = 8;     // 15-bit codes: 9 + 6
constexpr int kLitEntries = (1 << kLitBits) + kLitSubTables * (1 << kLitSubBits);      // 2048
constexpr int kDistEntries = (1 << kDistBits) + kDistSubTables * (1 << kDistSubBits);  // 1024
constexpr int kWarpsPerCta = 4;                     // 29 KB of shared memory per CTA, 7 CTAs = 28 warps per SM

enum : uint32_t { kTypeLit = 0, kTypeBase = 1, kTypeEob = 2, kTypeSub = 3 };
enum : uint32_t { kOk = 0, kErrCode = 1, kErrOverrun = 2, kErrTables = 3, kErrHeader = 4, kErrSize = 5 };

struct BlockDesc {
    uint64_t in_off;        // first byte of the raw deflate payload inside `comp`
    uint64_t out_off;
    uint32_t in_len, out_len;
};

// Table entry (16 bits, 0 = invalid): bits 0-3 code bits to consume at this level, 4-5 type, 6-15 payload = literal byte,
// length / distance SYMBOL (base and extra bits come from the constant tables, a uniform constant-cache read), or the
// number of the second-level table. Half the shared memory of 32-bit entries: 7.3 KB per warp, 28 warps per SM.
typedef uint16_t Entry;
struct WarpSmem {
    Entry lit[kLitEntries];
    Entry dist[kDistEntries];
    uint16_t code[320];     // canonical code of every symbol (scratch of the table build)
    uint8_t lens[320];
    Entry cl[128];          // code-length code table (7 bits)
};

__device__ __constant__ uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ __constant__ uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ __constant__ uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ __constant__ uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ __constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ Entry mk(uint32_t type, uint32_t bits, uint32_t payload)
{
    return (Entry)(bits | (type << 4) | (payload << 6));
}
__device__ __forceinline__ uint32_t e_bits(uint32_t e) { return e & 15u; }
__device__ __forceinline__ uint32_t e_type(uint32_t e) { return (e >> 4) & 3u; }
__device__ __forceinline__ uint32_t e_val(uint32_t e) { return e >> 6; }

// The compressed stream as the warp sees it: 32-bit units, unit u lives in lane (u / 2) % 32 of chunk u / 64.
struct BitReader {
    const uint64_t *base;       // 8-byte aligned start
    uint64_t n_words;           // readable 64-bit words (zero beyond)
    uint64_t cur, nxt;          // this lane's word of the current / next 256-byte chunk
    uint64_t bitbuf;
    int bitcnt;
    uint32_t unit;              // next 32-bit unit to append
    uint64_t consumed_limit_bits, consumed_bits;

    __device__ __forceinline__ uint64_t load_word(uint64_t w) const { return w < n_words ? __ldg(base + w) : 0ull; }
    __device__ __forceinline__ void init(const uint8_t *p, uint32_t len, uint32_t lane)
    {
        const uint64_t addr = reinterpret_cast<uint64_t>(p);
        base = reinterpret_cast<const uint64_t *>(addr & ~7ull);
        const uint32_t skip = (uint32_t)(addr & 7ull);
        n_words = ((uint64_t)skip + len + 7) / 8;
        cur = load_word(lane);
        nxt = load_word(32 + lane);
        bitbuf = 0;
        bitcnt = 0;
        unit = 0;
        consumed_bits = 0;
        consumed_limit_bits = (uint64_t)len * 8;
        refill();
        refill();
        // drop the bytes in front of the payload
        bitbuf >>= skip * 8;
        bitcnt -= (int)skip * 8;
        refill();
    }
    // append 32 bits while there is room for them
    __device__ __forceinline__ void refill()
    {
        if (bitcnt <= 32) {
            const uint32_t mine = (unit & 1u) ? (uint32_t)(cur >> 32) : (uint32_t)cur;
            const uint32_t v = __shfl_sync(0xffffffffu, mine, (unit >> 1) & 31u);
            bitbuf |= (uint64_t)v << bitcnt;
            bitcnt += 32;
            ++unit;
            if ((unit & 63u) == 0u) {                            // chunk exhausted: rotate, prefetch the one after the next
                cur = nxt;
                nxt = load_word((uint64_t)(unit >> 6) * 32 + 32 + (threadIdx.x & 31));
            }
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bitbuf & ((1u << n) - 1u); }
    __device__ __forceinline__ void consume(int n) { bitbuf >>= n; bitcnt -= n; consumed_bits += (uint64_t)n; }
    __device__ __forceinline__ uint32_t take(int n) { const uint32_t v = peek(n); consume(n); return v; }
};

// Canonical Huffman lengths (in sm.lens[0, n)) -> two-level table. All lanes call this; returns false (uniformly) when
// the code is over-subscribed or needs more second-level tables than there is room for.
__device__ bool build_table(WarpSmem &sm, int n, Entry *table, int P, int sub_bits, int max_sub, bool is_dist, uint32_t lane)
{
    // histogram of the lengths (lane-serial over 16 counters is cheap enough: n <= 320)
    uint32_t my_cnt = 0;                                         // lane l (1..15) counts length l
    if (lane >= 1 && lane <= 15)
        for (int s = 0; s < n; ++s) my_cnt += sm.lens[s] == lane;
    uint32_t next = 0, left = 1;
    bool over = false;
    uint32_t my_first = 0;                                       // first code of length `lane`
    for (int l = 1; l <= 15; ++l) {
        const uint32_t c = __shfl_sync(0xffffffffu, my_cnt, l);
        left <<= 1;
        if (c > left) over = true;
        left -= c;
        if ((int)lane == l) my_first = next;
        next = (next + c) << 1;
    }
    if (over) return false;
    // canonical code of every symbol: lane l walks the symbols of length l in order
    if (lane >= 1 && lane <= 15) {
        uint32_t code = my_first;
        for (int s = 0; s < n; ++s)
            if (sm.lens[s] == lane) sm.code[s] = (uint16_t)code++;
    }
    const int psize = 1 << P;
    for (int i = lane; i < psize; i += 32) table[i] = 0;
    __syncwarp();
    // second-level tables: claimed serially by lane 0 for every prefix that has a long code below it
    uint32_t n_sub = 0;
    bool ok = true;
    if (lane == 0) {
        for (int s = 0; s < n; ++s) {
            const int l = sm.lens[s];
            if (l <= P) continue;
            const uint32_t rev = __brev((uint32_t)sm.code[s]) >> (32 - l);
            const uint32_t prefix = rev & (uint32_t)(psize - 1);
            if (table[prefix] == 0) {
                if ((int)n_sub == max_sub) { ok = false; break; }
                table[prefix] = mk(kTypeSub, (uint32_t)P, n_sub);
                ++n_sub;
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
    n_sub = __shfl_sync(0xffffffffu, n_sub, 0);
    if (!ok) return false;
    for (uint32_t i = lane; i < n_sub * (1u << sub_bits); i += 32) table[psize + i] = 0;
    __syncwarp();
    // fill: every lane takes symbols lane, lane + 32, ...
    for (int s = lane; s < n; s += 32) {
        const int l = sm.lens[s];
        if (!l) continue;
        uint32_t type, value;
        if (is_dist) {
            if (s >= 30) continue;
            type = kTypeBase; value = (uint32_t)s;
        } else if (s < 256) { type = kTypeLit; value = (uint32_t)s; }
        else if (s == 256) { type = kTypeEob; value = 0; }
        else {
            if (s > 285) continue;
            type = kTypeBase; value = (uint32_t)(s - 257);
        }
        const uint32_t rev = __brev((uint32_t)sm.code[s]) >> (32 - l);
        if (l <= P) {
            const Entry e = mk(type, (uint32_t)l, value);
            for (uint32_t i = rev; i < (uint32_t)psize; i += 1u << l) table[i] = e;
        } else {
            const uint32_t off = (uint32_t)psize + e_val(table[rev & (uint32_t)(psize - 1)]) * (1u << sub_bits);
            const Entry e = mk(type, (uint32_t)(l - P), value);
            for (uint32_t i = rev >> P; i < (1u << sub_bits); i += 1u << (l - P)) table[off + i] = e;
        }
    }
    __syncwarp();
    return true;
}

__device__ __forceinline__ uint32_t lookup(const Entry *table, BitReader &br, int P, int sub_bits)
{
    uint32_t e = table[br.peek(P)];
    if (e_type(e) == kTypeSub) {
        br.consume(P);
        e = table[(1u << P) + e_val(e) * (1u << sub_bits) + br.peek(sub_bits)];
    }
    return e;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_bgzf_inflate(const uint8_t *__restrict__ comp, const BlockDesc *__restrict__ blocks, uint32_t n_blocks, uint8_t *__restrict__ out,
               uint32_t *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char zsmem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem &sm = reinterpret_cast<WarpSmem *>(zsmem)[warp];
    for (uint32_t b = blockIdx.x * kWarpsPerCta + warp; b < n_blocks; b += gridDim.x * kWarpsPerCta) {
        const BlockDesc d = blocks[b];
        uint8_t *const o = out + d.out_off;
        uint32_t op = 0;                                          // bytes written (uniform)
        uint32_t err = kOk;
        BitReader br;
        br.init(comp + d.in_off, d.in_len, lane);
        // literals are gathered in `lit_acc` and stored four at a time once the output position is 4-byte aligned
        // (out_off of a BGZF block is arbitrary, so alignment is relative to the real address)
        uint32_t lit_acc = 0, lit_n = 0;
        const uint32_t misalign = (uint32_t)(reinterpret_cast<uint64_t>(o) & 3ull);
        auto flush_lits = [&]() {
            if (lit_n && lane == 0)
                for (uint32_t k = 0; k < lit_n; ++k) o[op - lit_n + k] = (uint8_t)(lit_acc >> (8 * k));
            lit_n = 0;
            lit_acc = 0;
        };
        bool last = false;
        while (!last && err == kOk) {
            br.refill();
            last = br.take(1) != 0;
            const uint32_t btype = br.take(2);
            if (btype == 0) {
                // stored block: to the byte boundary, LEN / NLEN, raw bytes
                flush_lits();
                br.consume(br.bitcnt & 7);
                br.refill();
                const uint32_t len = br.take(16);
                br.refill();
                const uint32_t nlen = br.take(16);
                if ((len ^ 0xFFFFu) != nlen) { err = kErrHeader; break; }
                if (op + len > d.out_len) { err = kErrOverrun; break; }
                for (uint32_t k = 0; k < len; ++k) {              // byte-serial through the bit reader (rare: level-0 blocks)
                    br.refill();
                    const uint32_t v = br.take(8);
                    if (lane == 0) o[op + k] = (uint8_t)v;
                }
                op += len;
                __syncwarp();
                continue;
            }
            if (btype == 3) { err = kErrHeader; break; }
            int hlit = 288, hdist = 30;
            if (btype == 1) {
                for (int s = lane; s < 288; s += 32) sm.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
                __syncwarp();
                if (!build_table(sm, 288, sm.lit, kLitBits, kLitSubBits, kLitSubTables, false, lane)) { err = kErrTables; break; }
                for (int s = lane; s < 32; s += 32) sm.lens[s] = 5;
                __syncwarp();
                if (!build_table(sm, 32, sm.dist, kDistBits, kDistSubBits, kDistSubTables, true, lane)) { err = kErrTables; break; }
            } else {
                br.refill();
                hlit = (int)br.take(5) + 257;
                hdist = (int)br.take(5) + 1;
                const int hclen = (int)br.take(4) + 4;
                if (hlit > 286 || hdist > 30) { err = kErrHeader; break; }
                if (lane < 19) sm.lens[lane] = 0;
                __syncwarp();
                for (int i = 0; i < hclen; ++i) {
                    br.refill();
                    const uint32_t v = br.take(3);
                    if (lane == 0) sm.lens[kClOrder[i]] = (uint8_t)v;
                }
                __syncwarp();
                if (!build_table(sm, 19, sm.cl, 7, 0, 0, false, lane)) { err = kErrTables; break; }
                // the code lengths themselves (serial; every lane decodes, lane 0 stores)
                const int total = hlit + hdist;
                int i = 0;
                uint32_t prev = 0;
                while (i < total) {
                    br.refill();
                    const uint32_t e = sm.cl[br.peek(7)];
                    if (!e) { err = kErrCode; break; }
                    br.consume((int)e_bits(e));
                    const uint32_t sym = e_val(e);
                    if (sym < 16) {
                        if (lane == 0) sm.lens[i] = (uint8_t)sym;
                        prev = sym;
                        ++i;
                        continue;
                    }
                    uint32_t rep, v = 0;
                    if (sym == 16) { if (i == 0) { err = kErrHeader; break; } v = prev; rep = 3 + br.take(2); }
                    else if (sym == 17) rep = 3 + br.take(3);
                    else rep = 11 + br.take(7);
                    if (i + (int)rep > total) { err = kErrHeader; break; }
                    for (uint32_t k = lane; k < rep; k += 32) sm.lens[i + k] = (uint8_t)v;
                    prev = v;
                    i += (int)rep;
                }
                if (err != kOk) break;
                __syncwarp();
                if (sm.lens[256] == 0) { err = kErrHeader; break; }
                // the distance lengths follow the literal/length ones: build the distance table first (from a copy at the
                // front would disturb lit), so move them behind a gap
                uint8_t dl = lane < (uint32_t)hdist ? sm.lens[hlit + lane] : 0;
                __syncwarp();
                if (!build_table(sm, hlit, sm.lit, kLitBits, kLitSubBits, kLitSubTables, false, lane)) { err = kErrTables; break; }
                if (lane < 32) sm.lens[lane] = dl;
                __syncwarp();
                if (!build_table(sm, hdist, sm.dist, kDistBits, kDistSubBits, kDistSubTables, true, lane)) { err = kErrTables; break; }
            }
            // ---- symbols
            for (;;) {
                br.refill();                                      // >= 32 bits: a literal/length code and its extra bits (<= 20)
                uint32_t e = lookup(sm.lit, br, kLitBits, kLitSubBits);
                if (!e) { err = kErrCode; break; }
                br.consume((int)e_bits(e));
                const uint32_t type = e_type(e);
                if (type == kTypeLit) {
                    if (op >= d.out_len) { err = kErrOverrun; break; }
                    lit_acc |= e_val(e) << (8 * lit_n);
                    ++lit_n;
                    ++op;
                    if (lit_n == 4 || ((misalign + op) & 3u) == 0u) {
                        if (lit_n == 4 && ((misalign + op) & 3u) == 0u) {
                            if (lane == 0) *reinterpret_cast<uint32_t *>(o + op - 4) = lit_acc;
                            lit_n = 0;
                            lit_acc = 0;
                        } else {
                            flush_lits();
                        }
                    }
                    continue;
                }
                if (type == kTypeEob) break;
                const uint32_t ls = e_val(e);
                const uint32_t len = kLenBase[ls] + br.take((int)kLenExtra[ls]);
                br.refill();                                      // a distance code and its extra bits (<= 28)
                const uint32_t de = lookup(sm.dist, br, kDistBits, kDistSubBits);
                if (!de || e_type(de) != kTypeBase) { err = kErrCode; break; }
                br.consume((int)e_bits(de));
                const uint32_t ds = e_val(de);
                const uint32_t dist = kDistBase[ds] + br.take((int)kDistExtra[ds]);
                if (dist > op || op + len > d.out_len) { err = kErrOverrun; break; }
                flush_lits();
                __syncwarp();                                     // everything written so far is visible to all lanes
                const uint8_t *src = o + op - dist;
                if (dist >= len) {
                    for (uint32_t k = lane; k < len; k += 32) o[op + k] = src[k];
                } else {
                    for (uint32_t k = lane; k < len; k += 32) o[op + k] = src[k % dist];
                }
                op += len;
                __syncwarp();
            }
        }
        flush_lits();
        if (err == kOk && op != d.out_len) err = kErrSize;
        if (err == kOk && br.consumed_bits > br.consumed_limit_bits + 7) err = kErrOverrun;
        if (lane == 0) status[b] = err;
        __syncwarp();
    }
}

}  // namespace inqz
