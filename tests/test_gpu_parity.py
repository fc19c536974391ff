"""GPU parity: libinqcall.so (through the C ABI) against the CPU oracle on the same seeded inputs.
Integer/bit-exact: identical medians (as f64 bit patterns incl. NaN) for every locus."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.datagen import expected_events, make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import inquistr_b200 as q
    c = q.Context(0)
    yield c
    c.close()


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def run_case(ctx, case, minlen, support, unphased, chunks=1):
    rd = case["reads"]
    ctx.set_loci(case["contig_off"], case["locus_start"], case["locus_end"])
    ctx.clear_reads()
    if chunks == 1:
        ctx.push(rd)
    else:
        cuts = np.linspace(0, rd.n, chunks + 1).astype(int)
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b <= a:
                continue
            base = rd.cigar_off[a]
            ctx.push_reads(rd.contig[a:b], rd.ref_start[a:b], rd.ref_end[a:b], rd.mapq[a:b], rd.hp[a:b],
                           rd.flags[a:b], rd.cigar_off[a:b + 1] - base,
                           rd.cigar[int(base):int(rd.cigar_off[b])])
    res = ctx.genotype(minlen, support, unphased)
    rc, p1, p2, visits = O.genotype_loci(rd, case["n_contigs"], case["locus_contig"], case["locus_start"],
                                         case["locus_end"], minlen, support, unphased, threads=4)
    assert rc == 0
    bad = np.flatnonzero(~((res.phase1 == p1) | (np.isnan(res.phase1) & np.isnan(p1))) |
                         ~((res.phase2 == p2) | (np.isnan(res.phase2) & np.isnan(p2))))
    assert len(bad) == 0, (bad[:10], res.phase1[bad[:10]], p1[bad[:10]], res.phase2[bad[:10]], p2[bad[:10]])
    assert res.stats["op_visits"] == visits
    return res


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("unphased", [False, True])
def test_random_small(ctx, seed, unphased):
    case = make_case(seed)
    run_case(ctx, case, 5, 3, unphased)


@pytest.mark.parametrize("minlen,support", [(0, 1), (1, 2), (5, 1), (12, 5), (3, 40), (10**9, 3)])
def test_parameters(ctx, minlen, support):
    case = make_case(100 + minlen % 97 + support, n_reads=1500)
    run_case(ctx, case, minlen, support, False)
    run_case(ctx, case, minlen, support, True)


def test_unsorted_reads_and_chunked_push(ctx):
    case = make_case(11, sort_reads=False, n_reads=2000)
    a = run_case(ctx, case, 5, 3, False, chunks=1)
    b = run_case(ctx, case, 5, 3, False, chunks=7)
    assert same(a.phase1, b.phase1) and same(a.phase2, b.phase2)


def test_events_match_cigar_walk(ctx):
    case = make_case(21, n_reads=1200)
    for minlen in (5, 0):
        run_case(ctx, case, minlen, 3, False)
        pos, val, off = ctx.debug_events()
        epos, eval_, eoff = expected_events(case["reads"], minlen)
        assert np.array_equal(off, eoff)
        assert np.array_equal(pos, epos)
        assert np.array_equal(val, eval_)


def test_deep_loci_take_the_cta_path(ctx):
    # > 128 calls per locus: CTA sort in shared memory; > 4096: in-place global sort
    case = make_case(31, n_contigs=1, n_loci=6, n_reads=9000, dense_locus=True, max_read=3000)
    run_case(ctx, case, 5, 3, False)
    run_case(ctx, case, 5, 3, True)
    run_case(ctx, case, 0, 2000, True)


def test_long_reads_span_many_tiles(ctx):
    # reads with > 4096 CIGAR words cross tile boundaries of the scan kernel
    case = make_case(41, n_contigs=1, contig_len=3_000_000, n_loci=400, n_reads=300, max_read=900_000,
                     degenerate=True)
    res = run_case(ctx, case, 5, 3, False)
    assert res.stats["n_tiles"] > 50
    pos, val, off = ctx.debug_events()
    epos, eval_, eoff = expected_events(case["reads"], 5)
    assert np.array_equal(off, eoff) and np.array_equal(pos, epos) and np.array_equal(val, eval_)


def test_empty_inputs(ctx):
    case = make_case(51, n_reads=50)
    ctx.set_loci(case["contig_off"], case["locus_start"], case["locus_end"])
    ctx.clear_reads()
    res = ctx.genotype(5, 3, False)
    assert np.all(res.valid == 0)
    ctx.set_loci(np.zeros(case["n_contigs"] + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32))
    ctx.push(case["reads"])
    res = ctx.genotype(5, 3, False)
    assert len(res.valid) == 0


def test_error_codes(ctx):
    import inquistr_b200 as q
    case = make_case(61, hp_values=(0xFF, 1, 2, 3), hp_probs=(0.1, 0.4, 0.4, 0.1))
    ctx.set_loci(case["contig_off"], case["locus_start"], case["locus_end"])
    ctx.clear_reads()
    ctx.push(case["reads"])
    with pytest.raises(q.InqError) as ei:
        ctx.genotype(5, 3, False)
    assert ei.value.code == -10                       # HP 3 on a read that passes the filter
    rc, *_ = O.genotype_loci(case["reads"], case["n_contigs"], case["locus_contig"], case["locus_start"],
                             case["locus_end"], 5, 3, False)
    assert rc == O.ORC_PANIC_BAD_HP
    ctx.genotype(5, 3, True)                          # unphased ignores HP (call.rs:297-300)
    with pytest.raises(q.InqError) as ei:
        ctx.genotype(5, 0, False)                     # support 0 -> some empty bucket -> panic in the reference
    assert ei.value.code in (-10, -11)
    with pytest.raises(q.InqError) as ei:
        ctx.set_loci([0, 1], [5], [50])
    assert ei.value.code == -12
    with pytest.raises(q.InqError) as ei:
        ctx.set_loci([0, 2], [500, 400], [550, 450])
    assert ei.value.code == -13
    with pytest.raises(q.InqError) as ei:
        ctx.set_loci([0, 1], [500], [450])
    assert ei.value.code == -13
