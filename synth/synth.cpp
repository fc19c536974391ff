// synth.cpp -- seeded synthetic ONT-like reads + STR locus catalogs written straight
// into the structure-of-arrays layout the C ABI (include/inqcall.h) and the oracle consume.
// Benchmark / test infrastructure: not part of the product path, never linked into libinqcall.so.
//
// Deterministic for a given (seed, parameters): every read draws from its own counter-based
// stream (splitmix64 keyed by seed and read index), so the two passes (count, fill) and any
// thread count produce identical bytes.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint64_t below(uint64_t n) { return n ? (uint64_t)(((__uint128_t)next() * n) >> 64) : 0; }
    double normal()
    {
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    double gamma(double k)   // Marsaglia-Tsang, k >= 1
    {
        const double d = k - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = normal(), v = 1.0 + c * x;
            if (v <= 0) continue;
            v = v * v * v;
            double u = uniform();
            if (std::log(u < 1e-300 ? 1e-300 : u) < 0.5 * x * x + d - d * v + d * std::log(v)) return d * v;
        }
    }
};

inline uint64_t mix(uint64_t a, uint64_t b)
{
    Rng r(a * 0xD6E8FEB86659FD93ull ^ (b + 0x9E3779B97F4A7C15ull) * 0xCA5A826395121157ull);
    r.next();
    return r.next();
}

// geometric (support 1,2,...) by 16-bit inverse-CDF table
struct GeomLut {
    uint16_t v[65536];
    void init(double mean)
    {
        const double p = 1.0 / mean, lq = std::log(1.0 - p);
        for (int k = 0; k < 65536; ++k) {
            double u = (k + 0.5) / 65536.0;
            double g = 1.0 + std::floor(std::log(1.0 - u) / lq);
            v[k] = (uint16_t)std::min(g, 65535.0);
        }
    }
};

GeomLut g_mrun, g_indel, g_clip;
double g_mrun_mean = -1, g_indel_mean = -1, g_clip_mean = -1;

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_S = 4 };
inline uint32_t W(uint32_t len, uint32_t op) { return (len << 4) | op; }

}  // namespace

extern "C" {

struct synth_cfg {
    uint64_t seed;
    int32_t n_contigs;
    int32_t threads;
    const int64_t *contig_len;
    // catalog (sorted by contig,start)
    const int64_t *contig_locus_off;   // n_contigs+1
    const int32_t *lstart, *lend;
    const int32_t *delta_h1, *delta_h2; // allele at the locus start: +n = nI, -n = nD, 0 = none
    // read placement regions (sorted, disjoint): reads start uniformly inside [reg_start, reg_end)
    int64_t n_regions;
    const int32_t *reg_contig;
    const int64_t *reg_start, *reg_end;
    double depth;
    double gamma_shape, gamma_scale;   // read length ~ Gamma(shape) * scale, clipped to [min_len,max_len]
    int32_t min_len, max_len;
    double mrun_mean, indel_mean, clip_mean;
    double p_ins;                      // P(indel is an insertion)
    double p_clip;                     // P(leading soft clip) = P(trailing soft clip)
    double p_tagged;                   // P(read carries HP 1/2); 0 => all untagged
    double p_supp, p_2d_given_supp;    // SA-tagged reads / accidental-2D among them
    int32_t big_delta;                 // |delta| >= big_delta marks an expansion allele
    double p_big_trunc;                // P(read ends at an expansion with a matching trailing clip)
    // optional shard selection: only reads [sel_lo[g], sel_hi[g]) of each region are materialised
    // (same bytes as in the unsharded set: the RNG is keyed by the global read index)
    const uint64_t *sel_lo, *sel_hi;
};

// ---- loci ------------------------------------------------------------------------------------
// mode 0: genome-wide STR catalog (length 10+Geom(25) capped 1000, 60% hom-ref per haplotype,
//         otherwise +-k*motif with k~Geom(mean 3), motif 2..6, capped at 60)
// mode 1: expansion panel: loci evenly spaced, H1 small allele, H2 carries a 1-10 kb insertion
int synth_loci(uint64_t seed, int32_t n_contigs, const int64_t *contig_len, int64_t n_loci, int mode,
               int64_t *contig_locus_off, int32_t *lstart, int32_t *lend, int32_t *d1, int32_t *d2)
{
    long double total = 0;
    for (int c = 0; c < n_contigs; ++c) total += (long double)contig_len[c];
    std::vector<int64_t> per(n_contigs, 0);
    int64_t assigned = 0;
    for (int c = 0; c < n_contigs; ++c) {
        per[c] = (int64_t)std::floor((long double)n_loci * contig_len[c] / total);
        assigned += per[c];
    }
    for (int c = 0; assigned < n_loci; c = (c + 1) % n_contigs) { per[c]++; assigned++; }
    contig_locus_off[0] = 0;
    for (int c = 0; c < n_contigs; ++c) contig_locus_off[c + 1] = contig_locus_off[c] + per[c];
    GeomLut len_lut, k_lut;
    len_lut.init(25.0);
    k_lut.init(3.0);
    for (int c = 0; c < n_contigs; ++c) {
        const int64_t n = per[c], base = contig_locus_off[c];
        const int64_t lo = 10, span = contig_len[c] - 1200 - lo;
        if (n > 0 && span <= 0) return -1;
        for (int64_t j = 0; j < n; ++j) {
            Rng r(mix(seed ^ 0x10C1ull, (uint64_t)(base + j)));
            const double u = r.uniform();
            int64_t s = lo + (int64_t)(((double)j + u) * (double)span / (double)n);
            int32_t len = 10 + (int32_t)len_lut.v[r.next() & 0xFFFF];
            if (len > 1000) len = 1000;
            lstart[base + j] = (int32_t)s;
            lend[base + j] = (int32_t)(s + len);
            int32_t dd[2];
            for (int h = 0; h < 2; ++h) {
                if (mode == 1) {
                    dd[h] = (h == 0) ? (int32_t)(r.below(3) * 3) : (int32_t)(1000 + r.below(9001));
                } else if (r.uniform() < 0.6) {
                    dd[h] = 0;
                } else {
                    int32_t motif = 2 + (int32_t)r.below(5), k = (int32_t)k_lut.v[r.next() & 0xFFFF];
                    int32_t m = std::min(motif * k, 60);
                    if (r.next() & 1) m = -std::min(m, len);
                    dd[h] = m;
                }
            }
            d1[base + j] = dd[0];
            d2[base + j] = dd[1];
        }
    }
    return 0;
}

// ---- reads -----------------------------------------------------------------------------------
static double mean_len(const synth_cfg *c) { return c->gamma_shape * c->gamma_scale; }

static void ensure_luts(const synth_cfg *c)
{
    if (g_mrun_mean != c->mrun_mean) { g_mrun.init(c->mrun_mean); g_mrun_mean = c->mrun_mean; }
    if (g_indel_mean != c->indel_mean) { g_indel.init(c->indel_mean); g_indel_mean = c->indel_mean; }
    if (g_clip_mean != c->clip_mean) { g_clip.init(c->clip_mean); g_clip_mean = c->clip_mean; }
}

// reads per region; returns total
uint64_t synth_plan(const synth_cfg *c, uint64_t *reads_per_region)
{
    uint64_t tot = 0;
    for (int64_t g = 0; g < c->n_regions; ++g) {
        const double span = (double)(c->reg_end[g] - c->reg_start[g]);
        uint64_t n = (uint64_t)std::llround(c->depth * span / mean_len(c));
        reads_per_region[g] = n;
        tot += n;
    }
    return tot;
}

// number of reads the shard selection materialises
uint64_t synth_selected(const synth_cfg *c)
{
    std::vector<uint64_t> per(c->n_regions);
    synth_plan(c, per.data());
    uint64_t tot = 0;
    for (int64_t g = 0; g < c->n_regions; ++g) {
        const uint64_t a = c->sel_lo ? std::min(c->sel_lo[g], per[g]) : 0;
        const uint64_t b = c->sel_hi ? std::min(c->sel_hi[g], per[g]) : per[g];
        if (b > a) tot += b - a;
    }
    return tot;
}

}  // extern "C"

namespace {

struct ReadOut {
    int32_t contig, rs, re;
    uint8_t mapq, hp, flags;
    uint64_t n_cigar;
};

// Generates read i (global index), j-th of n in region g. If `out` is non-null the packed CIGAR is
// written there. Returns the header.
ReadOut gen_read(const synth_cfg *c, int64_t g, uint64_t j, uint64_t n, uint64_t i, uint32_t *out)
{
    Rng r(mix(c->seed, i));
    ReadOut h;
    const int32_t contig = c->reg_contig[g];
    const int64_t clen = c->contig_len[contig];
    const double span = (double)(c->reg_end[g] - c->reg_start[g]);
    int64_t start = c->reg_start[g] + (int64_t)(((double)j + r.uniform()) * span / (double)n);
    if (start > clen - 2) start = clen - 2;
    double len_d = r.gamma(c->gamma_shape) * c->gamma_scale;
    int64_t len = (int64_t)std::min(std::max(len_d, (double)c->min_len), (double)c->max_len);
    if (start + len > clen - 1) len = clen - 1 - start;
    if (len < 1) len = 1;
    const int64_t end_target = start + len;

    const double um = r.uniform();
    h.mapq = um < 0.92 ? 60 : (um < 0.97 ? (uint8_t)(11 + r.below(49)) : (uint8_t)r.below(11));
    const bool tagged = r.uniform() < c->p_tagged;
    const int hap = (int)(r.next() & 1);            // haplotype the read was sampled from
    h.hp = tagged ? (uint8_t)(1 + hap) : 0xFF;
    const bool supp = r.uniform() < c->p_supp;
    h.flags = (supp && r.uniform() < c->p_2d_given_supp) ? 1 : 0;
    h.contig = contig;
    h.rs = (int32_t)start;

    uint64_t nc = 0;
    auto emit = [&](uint32_t len_, uint32_t op) {
        if (out) out[nc] = W(len_, op);
        ++nc;
    };
    if (r.uniform() < c->p_clip) emit(g_clip.v[r.next() & 0xFFFF], OP_S);

    // next locus at or after the read start
    const int64_t l0 = c->contig_locus_off[contig], l1 = c->contig_locus_off[contig + 1];
    int64_t k = std::lower_bound(c->lstart + l0, c->lstart + l1, (int32_t)std::min<int64_t>(start + 1, INT32_MAX)) - c->lstart;
    const int32_t *delta = hap ? c->delta_h2 : c->delta_h1;

    int64_t pos = start;
    bool truncated = false;
    while (pos < end_target) {
        const uint64_t bits = r.next();
        int64_t m = g_mrun.v[bits & 0xFFFF];
        int64_t next_pos = std::min(pos + m, end_target);
        while (k < l1 && c->lstart[k] <= pos) ++k;                  // loci we are already past
        if (k < l1 && c->lstart[k] <= next_pos && c->lstart[k] < end_target) {
            // match up to the locus start, then this haplotype's allele anchored there
            const int64_t ls = c->lstart[k];
            emit((uint32_t)(ls - pos), OP_M);
            pos = ls;
            const int32_t d = delta[k];
            ++k;
            if (d >= c->big_delta && r.uniform() < c->p_big_trunc) {
                // read ends inside the expansion: trailing soft clip of the part it still covers
                emit((uint32_t)(100 + r.below((uint64_t)d - 99)), OP_S);
                truncated = true;
                break;
            }
            if (d > 0) emit((uint32_t)d, OP_I);
            else if (d < 0 && pos + (-d) < end_target) { emit((uint32_t)(-d), OP_D); pos += -d; }
            continue;
        }
        emit((uint32_t)(next_pos - pos), OP_M);
        pos = next_pos;
        if (pos >= end_target) break;
        const uint32_t il = g_indel.v[(bits >> 16) & 0xFFFF];
        const bool ins = (double)((bits >> 32) & 0xFFFF) < c->p_ins * 65536.0;
        if (ins) emit(il, OP_I);
        else if (pos + il < end_target) { emit(il, OP_D); pos += il; }
    }
    if (!truncated && r.uniform() < c->p_clip) emit(g_clip.v[r.next() & 0xFFFF], OP_S);
    h.re = (int32_t)(pos > start ? pos : start + 1);
    h.n_cigar = nc;
    return h;
}

// fn(region, j, n_in_region, global read index, local output index)
template <typename F>
void parallel_regions(const synth_cfg *c, const std::vector<uint64_t> &per, const std::vector<uint64_t> &base, F fn)
{
    // work items: chunks of up to 4096 reads within a region
    struct Item { int64_t g; uint64_t j0, j1, out0; };
    std::vector<Item> items;
    uint64_t out = 0;
    for (int64_t g = 0; g < c->n_regions; ++g) {
        const uint64_t a = c->sel_lo ? std::min(c->sel_lo[g], per[g]) : 0;
        const uint64_t b = c->sel_hi ? std::min(c->sel_hi[g], per[g]) : per[g];
        for (uint64_t j = a; j < b; j += 4096) {
            const uint64_t j1 = std::min<uint64_t>(j + 4096, b);
            items.push_back({g, j, j1, out});
            out += j1 - j;
        }
    }
    std::atomic<size_t> next{0};
    int nt = c->threads > 0 ? c->threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 256));
    auto worker = [&]() {
        for (;;) {
            size_t it = next.fetch_add(1);
            if (it >= items.size()) break;
            const Item &w = items[it];
            for (uint64_t j = w.j0; j < w.j1; ++j) fn(w.g, j, per[w.g], base[w.g] + j, w.out0 + (j - w.j0));
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

// pass 1: headers + per-read CIGAR length. n_cigar has R entries.
int synth_headers(const synth_cfg *c, uint64_t R, int32_t *contig, int32_t *rs, int32_t *re, uint8_t *mapq,
                  uint8_t *hp, uint8_t *flags, uint64_t *cig_off /* R+1, filled with exclusive prefix */)
{
    ensure_luts(c);
    std::vector<uint64_t> per(c->n_regions), base(c->n_regions);
    synth_plan(c, per.data());
    if (synth_selected(c) != R) return -1;
    uint64_t acc = 0;
    for (int64_t g = 0; g < c->n_regions; ++g) { base[g] = acc; acc += per[g]; }
    parallel_regions(c, per, base, [&](int64_t g, uint64_t j, uint64_t n, uint64_t gi, uint64_t i) {
        ReadOut h = gen_read(c, g, j, n, gi, nullptr);
        contig[i] = h.contig; rs[i] = h.rs; re[i] = h.re;
        mapq[i] = h.mapq; hp[i] = h.hp; flags[i] = h.flags;
        cig_off[i + 1] = h.n_cigar;
    });
    cig_off[0] = 0;
    for (uint64_t i = 0; i < R; ++i) cig_off[i + 1] += cig_off[i];
    return 0;
}

// pass 2: packed CIGAR words at the offsets computed by pass 1
int synth_cigars(const synth_cfg *c, uint64_t R, const uint64_t *cig_off, uint32_t *cigar)
{
    ensure_luts(c);
    std::vector<uint64_t> per(c->n_regions), base(c->n_regions);
    synth_plan(c, per.data());
    if (synth_selected(c) != R) return -1;
    uint64_t acc = 0;
    for (int64_t g = 0; g < c->n_regions; ++g) { base[g] = acc; acc += per[g]; }
    std::atomic<int> bad{0};
    parallel_regions(c, per, base, [&](int64_t g, uint64_t j, uint64_t n, uint64_t gi, uint64_t i) {
        ReadOut h = gen_read(c, g, j, n, gi, cigar + cig_off[i]);
        if (h.n_cigar != cig_off[i + 1] - cig_off[i]) bad.store(1);
    });
    return bad.load() ? -2 : 0;
}

// pass 2 for a packed subset: only reads with keep[i] != 0 are written, read i at cigar + dst_off[i]
// (the host packer's pre-filter, see synth.py: the random stream of every read is the same as without it)
int synth_cigars_sel(const synth_cfg *c, uint64_t R, const uint64_t *cig_off, const uint8_t *keep, const uint64_t *dst_off,
                     uint32_t *cigar)
{
    ensure_luts(c);
    std::vector<uint64_t> per(c->n_regions), base(c->n_regions);
    synth_plan(c, per.data());
    if (synth_selected(c) != R) return -1;
    uint64_t acc = 0;
    for (int64_t g = 0; g < c->n_regions; ++g) { base[g] = acc; acc += per[g]; }
    std::atomic<int> bad{0};
    parallel_regions(c, per, base, [&](int64_t g, uint64_t j, uint64_t n, uint64_t gi, uint64_t i) {
        if (!keep[i]) return;
        ReadOut h = gen_read(c, g, j, n, gi, cigar + dst_off[i]);
        if (h.n_cigar != cig_off[i + 1] - cig_off[i]) bad.store(1);
    });
    return bad.load() ? -2 : 0;
}

}  // extern "C"

// ---- BAM writer (benchmark input for the C++ host): multi-threaded BGZF deflate -----------------
#include <zlib.h>
#include <cstdio>
#include <string>

namespace {

struct OutBlock { std::vector<uint8_t> data; };

void put32(std::vector<uint8_t> &v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back((uint8_t)(x >> (8 * i))); }
void put16(std::vector<uint8_t> &v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }

int reg2bin(int64_t beg, int64_t end)
{
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// one BGZF block (<= 65280 payload bytes)
void bgzf_compress(const uint8_t *src, size_t n, int level, std::vector<uint8_t> &out)
{
    out.resize(18 + compressBound((uLong)n) + 8);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = (uInt)n;
    zs.next_out = out.data() + 18;
    zs.avail_out = (uInt)(out.size() - 18 - 8);
    deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    memcpy(out.data(), hdr, 16);
    const uint16_t bsize = (uint16_t)(clen + 25);
    out[16] = (uint8_t)bsize;
    out[17] = (uint8_t)(bsize >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n);
    uint8_t *t = out.data() + 18 + clen;
    for (int i = 0; i < 4; ++i) t[i] = (uint8_t)(crc >> (8 * i));
    for (int i = 0; i < 4; ++i) t[4 + i] = (uint8_t)((uint32_t)n >> (8 * i));
    out.resize(18 + clen + 8);
}

}  // namespace

extern "C" {

// Writes the SoA read set as a coordinate-sorted BAM. with_seq != 0 adds pseudo-random SEQ and constant
// QUAL of the query length implied by the CIGAR (realistic record sizes for ingest benchmarks).
// Reads flagged accidental-2D get an SA tag the host must classify as 2D. HP is written as type C.
// Returns bytes written, or -1.
int64_t synth_write_bam(const char *path, int32_t n_contigs, const char *const *names, const int64_t *lens, uint64_t R,
                        const int32_t *contig, const int32_t *rs, const int32_t *re, const uint8_t *mapq, const uint8_t *hp,
                        const uint8_t *flags, const uint64_t *cig_off, const uint32_t *cigar, int with_seq, int level,
                        int threads)
{
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 128));
    // header
    std::vector<uint8_t> head;
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (int c = 0; c < n_contigs; ++c) text += std::string("@SQ\tSN:") + names[c] + "\tLN:" + std::to_string(lens[c]) + "\n";
    head.insert(head.end(), {'B', 'A', 'M', 1});
    put32(head, (uint32_t)text.size());
    head.insert(head.end(), text.begin(), text.end());
    put32(head, (uint32_t)n_contigs);
    for (int c = 0; c < n_contigs; ++c) {
        const std::string nm = names[c];
        put32(head, (uint32_t)nm.size() + 1);
        head.insert(head.end(), nm.begin(), nm.end());
        head.push_back(0);
        put32(head, (uint32_t)lens[c]);
    }
    int64_t written = 0;
    auto emit_stream = [&](const std::vector<uint8_t> &raw) {
        // compress a raw byte stream as consecutive BGZF blocks, in parallel
        const size_t kPay = 65280;
        const size_t nb = (raw.size() + kPay - 1) / kPay;
        std::vector<std::vector<uint8_t>> comp(nb);
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= nb) break;
                const size_t off = i * kPay, n = std::min(kPay, raw.size() - off);
                bgzf_compress(raw.data() + off, n, level, comp[i]);
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        for (auto &c : comp) { fwrite(c.data(), 1, c.size(), fp); written += (int64_t)c.size(); }
    };
    emit_stream(head);
    // records: chunks of 2048 reads are serialised AND compressed by the worker threads (every chunk is its own run
    // of BGZF blocks); a wave of chunks is written in order before the next wave starts
    auto put_record = [&](std::vector<uint8_t> &raw, uint64_t i) {
        const uint64_t a = cig_off[i], b = cig_off[i + 1];
        const uint32_t ncig = (uint32_t)(b - a);
        uint32_t l_seq = 0;
        if (with_seq)
            for (uint64_t k = a; k < b; ++k) {
                const uint32_t op = cigar[k] & 15u;
                if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) l_seq += cigar[k] >> 4;
            }
        char name[32];
        const int ln = snprintf(name, sizeof(name), "r%llu", (unsigned long long)i) + 1;
        Rng rr(mix(0xBA11ull, i));
        const bool rev = rr.next() & 1;
        std::string sa;
        if (flags[i] & 1) {
            const int64_t mid = ((int64_t)rs[i] + re[i]) / 2 + 1;
            sa = std::string(names[contig[i]]) + "," + std::to_string(mid) + "," + (rev ? "+" : "-") + "," +
                 std::to_string(std::max<int64_t>(re[i] - mid, 1)) + "M10S,60,0;";
        }
        const bool long_cigar = ncig > 65535;
        size_t aux_len = (hp[i] != 0xFF ? 4 : 0) + (sa.empty() ? 0 : 3 + sa.size() + 1) + (long_cigar ? 8 + (size_t)ncig * 4 : 0);
        const uint32_t n_cig_rec = long_cigar ? 2 : ncig;
        const size_t body = 32 + (size_t)ln + (size_t)n_cig_rec * 4 + (l_seq + 1) / 2 + l_seq + aux_len;
        put32(raw, (uint32_t)body);
        put32(raw, (uint32_t)contig[i]);
        put32(raw, (uint32_t)rs[i]);
        raw.push_back((uint8_t)ln);
        raw.push_back(mapq[i]);
        put16(raw, (uint16_t)reg2bin(rs[i], std::max(re[i], rs[i] + 1)));
        put16(raw, (uint16_t)n_cig_rec);
        put16(raw, (uint16_t)(rev ? 0x10 : 0));
        put32(raw, l_seq);
        put32(raw, 0xFFFFFFFFu);
        put32(raw, 0xFFFFFFFFu);
        put32(raw, 0);
        raw.insert(raw.end(), name, name + ln);
        if (long_cigar) {
            put32(raw, (l_seq << 4) | 4u);
            put32(raw, ((uint32_t)(re[i] - rs[i]) << 4) | 3u);
        } else {
            const uint8_t *cp = reinterpret_cast<const uint8_t *>(cigar + a);
            raw.insert(raw.end(), cp, cp + (size_t)ncig * 4);
        }
        {
            // SEQ: pseudo-random A/C/G/T pairs (8 bases per draw) ; QUAL: pseudo-random Phred 10..41 (two-level mix, so that
            // the compressed stream is a mix of literals and matches as in real ONT BAMs, not one long run)
            const size_t ns = (l_seq + 1) / 2, base = raw.size();
            raw.resize(base + ns + l_seq);
            uint8_t *ps = raw.data() + base;
            for (size_t k = 0; k < ns; k += 8) {
                uint64_t r8 = rr.next();
                for (size_t j = k; j < std::min(ns, k + 8); ++j, r8 >>= 8) ps[j] = (uint8_t)((0x1u << (r8 & 3)) | (0x10u << ((r8 >> 2) & 3)));
            }
            uint8_t *pq = ps + ns;
            for (size_t k = 0; k < l_seq; k += 16) {
                uint64_t r8 = rr.next();
                const uint8_t level = (uint8_t)(18 + (r8 & 15));             // a local quality level, 16 bases long
                r8 >>= 4;
                for (size_t j = k; j < std::min<size_t>(l_seq, k + 16); ++j, r8 >>= 3) pq[j] = (uint8_t)(level + (r8 & 7) - 3);
            }
        }
        if (hp[i] != 0xFF) { raw.push_back('H'); raw.push_back('P'); raw.push_back('C'); raw.push_back(hp[i]); }
        if (!sa.empty()) { raw.push_back('S'); raw.push_back('A'); raw.push_back('Z'); raw.insert(raw.end(), sa.begin(), sa.end()); raw.push_back(0); }
        if (long_cigar) {
            raw.push_back('C'); raw.push_back('G'); raw.push_back('B'); raw.push_back('I');
            put32(raw, ncig);
            const uint8_t *cp = reinterpret_cast<const uint8_t *>(cigar + a);
            raw.insert(raw.end(), cp, cp + (size_t)ncig * 4);
        }
    };
    const uint64_t kChunk = 2048;
    const uint64_t n_chunks = (R + kChunk - 1) / kChunk;
    const uint64_t wave = (uint64_t)nt * 4;
    for (uint64_t c0 = 0; c0 < n_chunks; c0 += wave) {
        const uint64_t c1 = std::min(n_chunks, c0 + wave);
        std::vector<std::vector<uint8_t>> comp(c1 - c0);
        std::atomic<uint64_t> next{c0};
        auto work = [&]() {
            std::vector<uint8_t> raw, blk;
            for (;;) {
                const uint64_t c = next.fetch_add(1);
                if (c >= c1) break;
                raw.clear();
                for (uint64_t i = c * kChunk; i < std::min(R, (c + 1) * kChunk); ++i) put_record(raw, i);
                const size_t kPay = 65280;
                std::vector<uint8_t> &out = comp[c - c0];
                for (size_t off = 0; off < raw.size(); off += kPay) {
                    bgzf_compress(raw.data() + off, std::min(kPay, raw.size() - off), level, blk);
                    out.insert(out.end(), blk.begin(), blk.end());
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        for (auto &c : comp) { fwrite(c.data(), 1, c.size(), fp); written += (int64_t)c.size(); }
    }
    static const uint8_t eof_block[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof_block, 1, 28, fp);
    written += 28;
    fclose(fp);
    return written;
}

}  // extern "C"
