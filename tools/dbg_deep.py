import sys; sys.path.insert(0,'.')
import numpy as np
import inquistr_b200 as q
from tests.datagen import make_case
from oracle import oracle as O
ctx = q.Context(0)
def run(case, *args):
    rd = case["reads"]
    ctx.set_loci(case["contig_off"], case["locus_start"], case["locus_end"])
    ctx.clear_reads(); ctx.push(rd)
    try:
        res = ctx.genotype(*args)
        rc,p1,p2,v = O.genotype_loci(rd, case["n_contigs"], case["locus_contig"], case["locus_start"], case["locus_end"], *args, threads=4)
        print(args, "ok", np.array_equal(res.phase1,p1,equal_nan=True), np.array_equal(res.phase2,p2,equal_nan=True), res.stats["n_pairs"], res.stats["n_candidates"], res.stats["n_events"])
    except Exception as e:
        print(args, "ERR", e)
small = make_case(0)
run(small, 5, 3, False)
c61 = make_case(61, hp_values=(0xFF, 1, 2, 3), hp_probs=(0.1, 0.4, 0.4, 0.1))
deep = make_case(31, n_contigs=1, n_loci=6, n_reads=9000, dense_locus=True, max_read=3000)
run(deep, 5, 3, False)
run(small, 5, 3, False)
run(c61, 5, 3, False)
run(deep, 5, 3, False)
run(deep, 5, 3, False)
