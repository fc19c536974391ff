"""Known-answer vectors for the CPU oracle, derived by hand from the cited reference lines
(/root/reference/src/call.rs, v0.13.0). Locus start=1000,end=1050 => window (990,1060).
The reference has no golden `call` output: parity is unpinned, these vectors
pin the oracle to the source text instead. Each case is run through the C oracle and the
independent pure-Python mirror."""
import math

import numpy as np
import pytest

from oracle import oracle as O

S, E, M = 990, 1060, 5


def both_walk(pos0, cigar, is2d=False, minlen=M, s=S, e=E):
    w = O.pack_cigar(cigar)
    a = O.call_from_cigar(pos0, w, minlen, s, e, is2d)
    b = O.py_call_from_cigar(pos0, w, minlen, s, e, is2d)
    assert a == b, (a, b)
    return a


# ---- call.rs:377-413 ----------------------------------------------------------------------
@pytest.mark.parametrize("pos0,cigar,expect", [
    (900, "100M10I200M", (10, False)),          # P=1001 inside, 10 > 5
    (900, "95M6D200M", (-6, False)),            # P=996, deletion subtracts
    (900, "100M5I200M", (0, False)),            # strict >: 5 is not counted (call.rs:400)
    (900, "100M6I200M", (6, False)),
    (889, "100M6I100M", (0, False)),            # P=990: start < P is strict
    (890, "100M6I100M", (6, False)),            # P=991
    (958, "100M6I100M", (6, False)),            # P=1059 < 1060
    (959, "100M6I100M", (0, False)),            # P=1060: P < end is strict
    (884, "100M20D100M", (0, False)),           # anchor 985 outside although it spans in
    (900, "100M12I10M7D100M", (5, False)),      # +12 -7
    (900, "10H50M50N6I10P100M", (6, False)),    # H,P ignored; N advances (call.rs:404-405)
    (900, "50=50X8I100M", (8, False)),          # = and X advance (call.rs:384)
    (995, "200S500M", (200, True)),             # leading clip anchored at P=996 -> Clip
    (500, "520M300S", (300, True)),             # trailing clip anchored at P=1021
    (995, "5S500M", (0, False)),                # clip not longer than minlen -> Span(0)
    (100, "500M", (0, False)),
    (900, "95M6D4M7D200M", (-13, False)),       # second D anchored at 996+6+4=1006
])
def test_walk(pos0, cigar, expect):
    assert both_walk(pos0, cigar) == expect


def test_walk_2d_flag_suppresses_clip_only():
    assert both_walk(995, "200S10M8I400M", is2d=True) == (8, False)
    assert both_walk(995, "200S10M8I400M", is2d=False) == (208, True)


def test_walk_minlen_zero_and_large():
    assert both_walk(900, "100M1I10M1D100M", minlen=0) == (0, False)
    assert both_walk(900, "100M2I10M1D100M", minlen=0) == (1, False)
    assert both_walk(900, "100M2000I100M", minlen=1999) == (2000, False)
    assert both_walk(900, "100M2000I100M", minlen=2000) == (0, False)


def test_walk_empty_cigar():
    assert both_walk(1000, "") == (0, False)


# ---- call.rs:497-522 ----------------------------------------------------------------------
def med(spans, clips=(), support=3):
    vals = list(spans) + list(clips)
    flags = [0] * len(spans) + [1] * len(clips)
    c, panicked = O.median_str_length(vals, flags, support)
    calls = [(v, False) for v in spans] + [(v, True) for v in clips]
    try:
        p = O.py_median_str_length(calls, support)
        assert not panicked
        assert (math.isnan(c) and math.isnan(p)) or c == p
    except IndexError:
        assert panicked
    return c, panicked


@pytest.mark.parametrize("spans,clips,support,expect", [
    ([10, 10, 12], [], 3, 10.0),
    ([10, 11, 12, 13], [], 3, 11.5),
    ([-1, 0, 0, 0], [], 3, 0.0),
    ([-3, -2, 2, 2], [], 3, 0.0),
    ([-1, 0], [], 2, -0.5),
    ([5], [100, 50, 20], 3, 50.0),              # top-up with the 2 largest clips
    ([1, 2, 3, 4], [1000], 3, 2.5),             # more than support spans: clips ignored
    ([1, 2, 3], [1000], 3, 2.0),                # `<=` appends zero clips (call.rs:509-513)
    ([], [7, 9, 8], 3, 8.0),
    ([], [7, 9, 8, 100], 3, 9.0),               # only the top 3 clips
    ([4], [], 1, 4.0),
    ([1, 2], [50], 1, 1.5),                     # 2 spans > support 1
    ([1], [50, 60], 1, 1.0),                    # 1 <= 1: appends 0 clips
])
def test_median(spans, clips, support, expect):
    got, panicked = med(spans, clips, support)
    assert not panicked and got == expect


def test_median_below_support_is_nan():
    got, _ = med([10, 11], [], 3)
    assert math.isnan(got)
    got, _ = med([], [], 3)
    assert math.isnan(got)


def test_median_support_zero_empty_panics():
    _, panicked = med([], [], 0)
    assert panicked
    got, panicked = med([3, 5], [], 0)
    assert not panicked and got == 4.0
    _, panicked = med([], [5], 0)              # no spans, takes 0 clips -> empty -> panic
    assert panicked


# ---- call.rs:279-374 ----------------------------------------------------------------------
def geno(recs, unphased, support=3, minlen=5, start=1000, end=1050, threads=1):
    rd = O.Reads.from_records(recs)
    rc, p1, p2, _ = O.genotype_loci(rd, 1, [0], [start], [end], minlen, support, unphased, threads)
    try:
        q1, q2 = O.py_genotype_locus(rd, 0, start, end, minlen, support, unphased)
        assert rc == 0
        for a, b in ((p1[0], q1), (p2[0], q2)):
            assert (math.isnan(a) and math.isnan(b)) or a == b
    except (KeyError, IndexError, OverflowError):
        assert rc != 0
    return rc, p1[0], p2[0]


def span_read(pos, length, ins=0, **kw):
    left = 1001 - (pos + 1)
    cig = f"{left}M{ins}I{length - left}M" if ins else f"{length}M"
    return dict(pos=pos, cigar=cig, **kw)


def test_unphased_filter_edges():
    # kept iff ref_start <= 990 and ref_end >= 1060 and mapq > 10 (call.rs:297-300)
    keep = [span_read(990, 70, ins=8), span_read(900, 300, ins=8), span_read(100, 2000, ins=8),
            span_read(900, 300, ins=8, mapq=11)]
    drop = [span_read(991, 300, ins=50), span_read(900, 159, ins=50),   # end = 1059
            span_read(900, 300, ins=50, mapq=10), span_read(900, 300, ins=50, mapq=0)]
    rc, p1, p2 = geno(keep + drop, unphased=True, support=2)
    assert rc == 0 and p1 == 8.0 and p2 == 8.0


def test_unphased_split_sizes():
    vals = [0, 0, 0, 30, 30, 33]
    rc, p1, p2 = geno([span_read(900, 300, ins=v) if v else span_read(900, 300) for v in vals], True)
    assert (rc, p1, p2) == (0, 0.0, 30.0)
    rc, p1, p2 = geno([span_read(900, 300, ins=v) for v in (6, 7, 8, 9, 10)], True)
    assert rc == 0 and math.isnan(p1) and p2 == 9.0          # n=5: H1 gets 2 -> NaN


def test_phased_filter_edges():
    recs = [
        span_read(900, 300, ins=10, hp=1), span_read(900, 300, ins=10, hp=1),
        span_read(900, 300, ins=12, hp=1),
        span_read(900, 300, ins=99),                          # no HP -> dropped
        dict(pos=995, cigar="30M40I30M", hp=1),               # [995,1055) strictly inside -> dropped
        dict(pos=995, cigar="6M20I999M", hp=2),               # starts inside, ends beyond -> kept
        dict(pos=100, cigar="900M", hp=2),                    # ends at 1000 > 990 -> kept, call 0
        dict(pos=100, cigar="895M7D5M", hp=2),                # P=996, kept, -7
        span_read(900, 300, ins=50, hp=2, mapq=10),           # mapq <= 10 dropped
        span_read(900, 300, ins=77, hp=0),                    # HP 0: kept but in the ignored bucket
        dict(pos=100, cigar="890M", hp=2),                    # ends at 990: not fetched (endpos > 990 fails)
        dict(pos=1060, cigar="50M", hp=2),                    # pos < 1060 fails: not fetched
    ]
    rc, p1, p2 = geno(recs, unphased=False)
    assert rc == 0 and p1 == 10.0 and p2 == 0.0              # H2 = [20, 0, -7] -> 0


def test_phased_bad_hp_panics():
    rc, _, _ = geno([span_read(900, 300, ins=10, hp=3)], unphased=False)
    assert rc == O.ORC_PANIC_BAD_HP
    # a bad HP on a read that is filtered out never reaches the unwrap (call.rs:350-358)
    rc, _, _ = geno([span_read(900, 300, ins=10, hp=3, mapq=5)], unphased=False)
    assert rc == 0


def test_clip_topup_in_locus():
    recs = [span_read(900, 300, ins=6, hp=1),
            dict(pos=995, cigar="400S800M", hp=1), dict(pos=995, cigar="300S800M", hp=1),
            dict(pos=995, cigar="100S800M", hp=1)]
    rc, p1, p2 = geno(recs, unphased=False)
    assert rc == 0 and p1 == 300.0 and math.isnan(p2)        # [6,300,400] -> 300
    recs[1]["is2d"] = True                                    # 2D read: its clip is not counted -> Span(0)
    rc, p1, _ = geno(recs, unphased=False)
    assert rc == 0 and p1 == 6.0                              # spans [6,0] + top clip 300 -> [0,6,300]


def test_start_below_10_rejected():
    rc, _, _ = geno([span_read(0, 300)], True, start=5, end=50)
    assert rc == O.ORC_PANIC_START_LT_10


def test_threads_same_answer():
    rng = np.random.default_rng(7)
    recs = [span_read(900, 300, ins=int(rng.integers(6, 40)), hp=int(rng.integers(1, 3)))
            for _ in range(40)]
    a = geno(recs, False, threads=1)
    b = geno(recs, False, threads=4)
    assert a == b


# ---- call.rs:415-477 ----------------------------------------------------------------------
def test_cigar_to_rlen():
    assert O.cigar_to_rlen("500M") == 500
    assert O.cigar_to_rlen("10S100M5I20D30N7=3X8H") == 100 + 20 + 30 + 7 + 3


def test_accidental_2d():
    sa = "chr7,1500,-,500M,60,0;"
    assert O.is_accidental_2d(False, sa, 1000, 2000)
    assert not O.is_accidental_2d(True, sa, 1000, 2000)                  # same strand
    assert not O.is_accidental_2d(False, sa + "chr7,9000,-,50M,60,0;", 1000, 2000)  # two entries
    assert O.is_accidental_2d(False, "chr9,1500,-,500M,60,0;", 1000, 2000)          # rname ignored
    assert not O.is_accidental_2d(False, "chr7,2000,-,500M,60,0;", 1000, 2000)      # max(start) == min(end)
    assert O.is_accidental_2d(False, "chr7,1999,-,500M,60,0;", 1000, 2000)
    assert not O.is_accidental_2d(False, "chr7,400,-,100S600M,60,0;", 1000, 2000)   # sa_end = 1000
    assert O.is_accidental_2d(False, "chr7,400,-,100S601M,60,0;", 1000, 2000)
    assert not O.is_accidental_2d(False, None, 1000, 2000)


# ---- call.rs:33-38, 57-65; repeats.rs:96-115 ----------------------------------------------
def test_format():
    assert O.format_f64(12.0) == "12"
    assert O.format_f64(-6.0) == "-6"
    assert O.format_f64(0.0) == "0"
    assert O.format_f64(11.5) == "11.5"
    assert O.format_f64(-0.5) == "-0.5"
    assert O.format_f64(-3.5) == "-3.5"
    assert O.format_f64(0.5) == "0.5"
    assert O.format_f64(float("nan")) == "NaN"
    assert O.format_f64(123456789.5) == "123456789.5"
    assert O.format_row("chr7", 154778571, 154779363, 12.0, float("nan")) == \
        "chr7\t154778571\t154779363\t12\tNaN"


def test_human_order():
    import functools
    names = ["chr10", "chr2", "chrX", "chr1", "chrM", "chr1_KI270706v1_random", "chrY", "chr22",
             "chrEBV", "chrUn_GL000195v1"]
    got = sorted(names, key=functools.cmp_to_key(O.human_compare))
    assert got == ["chr1", "chr1_KI270706v1_random", "chr2", "chr10", "chr22", "chrEBV", "chrM",
                   "chrUn_GL000195v1", "chrX", "chrY"]
    assert O.human_compare("chr7", "chr7") == 0
    assert O.human_compare("chr07", "chr7") < 0   # numeric tie -> exhausted -> plain string compare


def test_validate_interval():
    assert O.validate_interval(100, 200, 1000) == 0
    assert O.validate_interval(200, 100, 1000) != 0       # end < start
    assert O.validate_interval(100, 1000, 1000) != 0      # end must be < chrom_len
    assert O.validate_interval(100, 999, 1000) == 0
    assert O.validate_interval(100, 200, -1) != 0         # contig not in header
    # the only value the reference's own tests assert on this path (call.rs:600-605)
    assert O.validate_interval(154778571, 154779363, 159345973) == 0
