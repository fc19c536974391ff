"""Small seeded random read/locus sets with the edge cases the hot path must survive.
Independent of inquistr_b200/synth (the benchmark generator) on purpose."""
from __future__ import annotations

import numpy as np

from oracle.oracle import Reads, cigar_ref_len

OPS = {"M": 0, "I": 1, "D": 2, "N": 3, "S": 4, "H": 5, "P": 6, "=": 7, "X": 8}


def random_cigar(rng, target_ref, *, indel_rate=0.5, big_rate=0.4, clip_rate=0.2, exotic=True):
    """packed CIGAR words consuming about target_ref reference bases"""
    words = []
    if rng.random() < clip_rate:
        if exotic and rng.random() < 0.3:
            words.append((int(rng.integers(1, 50)) << 4) | OPS["H"])
        words.append((int(rng.integers(1, 400)) << 4) | OPS["S"])
    consumed = 0
    while consumed < target_ref:
        run = int(min(target_ref - consumed, rng.geometric(1 / 40.0)))
        mop = "M"
        if exotic:
            u = rng.random()
            mop = "=" if u < 0.05 else "X" if u < 0.08 else "M"
        words.append((run << 4) | OPS[mop])
        consumed += run
        if consumed >= target_ref:
            break
        if rng.random() < indel_rate:
            ln = int(rng.integers(6, 60)) if rng.random() < big_rate else int(rng.integers(1, 7))
            u = rng.random()
            if u < 0.45:
                words.append((ln << 4) | OPS["I"])
            elif u < 0.92:
                words.append((ln << 4) | OPS["D"])
                consumed += ln
            elif exotic and u < 0.96:
                words.append((ln << 4) | OPS["N"])
                consumed += ln
            elif exotic:
                words.append((ln << 4) | OPS["P"])
    if rng.random() < clip_rate:
        words.append((int(rng.integers(1, 400)) << 4) | OPS["S"])
    return np.asarray(words, dtype=np.uint32)


def make_case(seed, n_contigs=3, contig_len=60_000, n_loci=120, n_reads=900, *, phased=True,
              sort_reads=True, hp_values=(0xFF, 0, 1, 2), hp_probs=(0.1, 0.05, 0.45, 0.40),
              max_read=12_000, degenerate=True, exotic=True, dense_locus=False):
    rng = np.random.default_rng(seed)
    # loci: overlapping / nested / duplicated allowed, sorted by (contig, start)
    lc = rng.integers(0, n_contigs, n_loci)
    ls = rng.integers(10, contig_len - 1500, n_loci)
    ll = np.where(rng.random(n_loci) < 0.1, 0, rng.geometric(1 / 40.0, n_loci))
    ll = np.minimum(ll, 900)
    if n_loci >= 8:
        ls[1] = ls[0]; lc[1] = lc[0]                      # duplicate start
        ls[3] = ls[2] + 5; lc[3] = lc[2]; ll[2] = 800     # nested locus
        ls[4] = 10                                        # smallest legal start
    if dense_locus:
        lc[:] = 0
        ls[:] = rng.integers(20_000, 20_400, n_loci)
    le = ls + ll
    order = np.lexsort((ls, lc))
    lc, ls, le = lc[order], ls[order], le[order]
    contig_off = np.searchsorted(lc, np.arange(n_contigs + 1)).astype(np.int64)

    # per-locus, per-HP allele: an I (+) or D (-) placed at the locus start of reads aimed at it
    allele = np.where(rng.random((max(n_loci, 1), 256)) < 0.5, 0, rng.integers(-60, 61, (max(n_loci, 1), 256)))
    contig, rs, re_, mq, hp, fl, words, off = [], [], [], [], [], [], [], [0]
    for i in range(n_reads):
        u = rng.random()
        aimed = -1
        hp_pre = None
        if u < 0.7 and n_loci:
            k = int(rng.integers(0, max(1, int(n_loci * 0.8))))  # aim at a locus (20% stay sparse)
            c = int(lc[k])
            length = int(rng.integers(50, max_read))
            start = int(ls[k]) - int(rng.integers(-30, length))
            aimed = k
        else:
            c = int(rng.integers(0, n_contigs))
            length = int(rng.integers(50, max_read))
            start = int(rng.integers(0, contig_len))
        start = max(0, min(start, contig_len - 2))
        length = max(1, min(length, contig_len - start - 1))
        unmapped = False
        if degenerate and rng.random() < 0.02:
            w = np.zeros(0, np.uint32)                   # placed read without CIGAR
            unmapped = rng.random() < 0.5
        elif degenerate and rng.random() < 0.02:
            w = np.asarray([(int(rng.integers(1, 30)) << 4) | OPS["S"],
                            (int(rng.integers(1, 30)) << 4) | OPS["I"]], np.uint32)  # rlen 0
        else:
            h = int(rng.choice(hp_values, p=hp_probs))
            d = int(allele[aimed, h]) if aimed >= 0 else 0
            left = int(ls[aimed]) - start + 1 if aimed >= 0 else 0
            if d != 0 and 0 < left < length - abs(d) - 1 and rng.random() < 0.8:
                op = OPS["I"] if d > 0 else OPS["D"]
                w = np.concatenate([random_cigar(rng, left, exotic=exotic, clip_rate=0.05),
                                    np.asarray([(abs(d) << 4) | op], np.uint32),
                                    random_cigar(rng, length - left - (abs(d) if d < 0 else 0), exotic=exotic,
                                                 clip_rate=0.05)])
            else:
                w = random_cigar(rng, length, exotic=exotic)
            hp_pre = h
        rlen = 0 if unmapped else cigar_ref_len(w)
        if rlen == 0:
            rlen = 1
        contig.append(c); rs.append(start); re_.append(start + rlen)
        mq.append(int(rng.choice([0, 5, 10, 11, 20, 60], p=[0.03, 0.03, 0.04, 0.05, 0.1, 0.75])))
        hp.append(hp_pre if hp_pre is not None else int(rng.choice(hp_values, p=hp_probs)))
        fl.append(1 if rng.random() < 0.08 else 0)
        words.append(w)
        off.append(off[-1] + len(w))
    contig = np.asarray(contig); rs = np.asarray(rs)
    idx = np.lexsort((rs, contig)) if sort_reads else rng.permutation(n_reads)
    cig = [words[i] for i in idx]
    off = np.concatenate([[0], np.cumsum([len(w) for w in cig])]).astype(np.uint64)
    reads = Reads(contig[idx], rs[idx], np.asarray(re_)[idx], np.asarray(mq)[idx], np.asarray(hp)[idx],
                  np.asarray(fl)[idx], off, np.concatenate(cig) if cig else np.zeros(0, np.uint32))
    return dict(n_contigs=n_contigs, contig_off=contig_off, locus_contig=lc.astype(np.int32),
                locus_start=ls.astype(np.int32), locus_end=le.astype(np.int32), reads=reads)


def expected_events(reads: Reads, minlen: int):
    """per-read event lists the CIGAR scan must produce: (pos1, val) with val = (signed len << 1) | is_S"""
    pos_all, val_all, off = [], [], [0]
    for r in range(reads.n):
        a, b = int(reads.cigar_off[r]), int(reads.cigar_off[r + 1])
        p = (int(reads.ref_start[r]) + 1) & 0xFFFFFFFF
        for w in reads.cigar[a:b]:
            w = int(w); op, ln = w & 15, w >> 4
            if op in (1, 2, 4) and ln > minlen:
                pos_all.append(p)
                val_all.append(((-ln if op == 2 else ln) << 1) | (1 if op == 4 else 0))
            if op in (0, 2, 3, 7, 8):
                p = (p + ln) & 0xFFFFFFFF
        off.append(len(pos_all))
    return np.asarray(pos_all, np.uint32), np.asarray(val_all, np.int32), np.asarray(off, np.uint32)
