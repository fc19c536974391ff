"""Regenerates tests/golden/*.npz: small seeded inputs + the oracle's outputs for them.

Run from the repo root:  python tests/golden/make_golden.py
The reference itself cannot be run here (Rust, no toolchain; its test BAM is absent), so these
vectors pin the ORACLE (and through it the CUDA path) against regressions; they are not outputs of
the reference binary (parity unpinned, see oracle/oracle.h)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from synth.synth import make_workload  # noqa: E402
from tests.datagen import make_case  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, n_contigs, contig_off, lc, ls, le, rd, runs):
    d = dict(n_contigs=n_contigs, contig_off=contig_off, locus_contig=lc, locus_start=ls, locus_end=le,
             contig=rd.contig, ref_start=rd.ref_start, ref_end=rd.ref_end, mapq=rd.mapq, hp=rd.hp,
             flags=rd.flags, cigar_off=rd.cigar_off, cigar=rd.cigar)
    for (minlen, support, unphased) in runs:
        rc, p1, p2, visits = O.genotype_loci(rd, n_contigs, lc, ls.astype(np.uint32), le.astype(np.uint32), minlen,
                                             support, unphased, threads=2)
        assert rc == 0
        key = f"m{minlen}_s{support}_u{int(unphased)}"
        d[key + "_p1"], d[key + "_p2"], d[key + "_visits"] = p1, p2, np.int64(visits)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k.endswith("_p1")})


RUNS = [(5, 3, False), (5, 3, True), (0, 1, False), (12, 5, True)]

if __name__ == "__main__":
    c = make_case(2024, n_contigs=2, contig_len=40_000, n_loci=80, n_reads=500, max_read=6000)
    save("random_edge_cases", c["n_contigs"], c["contig_off"], c["locus_contig"], c["locus_start"], c["locus_end"],
         c["reads"], RUNS)
    w = make_workload(4, threads=2)          # expansion panel: long insertions, clip top-up, 100x
    # keep the fixture small: first 6 loci and the reads on their contigs
    keep_c = np.unique(w.locus_contig[:6])
    m = np.isin(w.reads.contig, keep_c)
    idx = np.flatnonzero(m)
    off = np.concatenate([[0], np.cumsum((w.reads.cigar_off[1:] - w.reads.cigar_off[:-1])[idx])]).astype(np.uint64)
    cig = np.concatenate([w.reads.cigar[int(w.reads.cigar_off[i]):int(w.reads.cigar_off[i + 1])] for i in idx])
    rd = O.Reads(w.reads.contig[idx], w.reads.ref_start[idx], w.reads.ref_end[idx], w.reads.mapq[idx], w.reads.hp[idx],
                 w.reads.flags[idx], off, cig)
    lm = np.isin(w.locus_contig, keep_c)
    lc = w.locus_contig[lm]
    coff = np.searchsorted(lc, np.arange(w.n_contigs + 1)).astype(np.int64)
    save("expansion_panel", w.n_contigs, coff, lc, w.locus_start[lm], w.locus_end[lm], rd, RUNS[:2])
