"""small end-to-end pass for compute-sanitizer (memcheck): all kernels incl. CTA sort paths and regrow"""
import sys
sys.path.insert(0, ".")
import numpy as np
import inquistr_b200 as q
from oracle import oracle as O
from tests.datagen import make_case

ctx = q.Context(0)
for seed, kw, runs in ((0, {}, [(5, 3, False), (5, 3, True), (0, 1, False)]),
                       (31, dict(n_contigs=1, n_loci=6, n_reads=6000, dense_locus=True, max_read=2000), [(5, 3, False), (5, 3, True)]),
                       (41, dict(n_contigs=1, contig_len=400_000, n_loci=60, n_reads=40, max_read=200_000), [(5, 3, False)])):
    case = make_case(seed, **kw)
    rd = case["reads"]
    ctx.set_loci(case["contig_off"], case["locus_start"], case["locus_end"])
    ctx.clear_reads()
    ctx.push(rd)
    for args in runs:
        res = ctx.genotype(*args)
        rc, p1, p2, v = O.genotype_loci(rd, case["n_contigs"], case["locus_contig"], case["locus_start"], case["locus_end"], *args, threads=4)
        assert np.array_equal(res.phase1, p1, equal_nan=True) and np.array_equal(res.phase2, p2, equal_nan=True)
        print("ok", seed, args, res.stats["n_pairs"], res.stats["n_events"])
ctx.close()
