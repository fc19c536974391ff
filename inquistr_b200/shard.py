"""Range-sharding of the sorted locus catalog across GPUs: contiguous slices of the
(contig, start)-sorted catalog, every rank gets the reads that can reach its slice (reads at a cut are
duplicated), results are concatenated in rank order. No collective on the data path."""
from __future__ import annotations

import numpy as np


def split_catalog(n_loci: int, world: int) -> list[tuple[int, int]]:
    """[lo, hi) of the sorted catalog owned by each rank (balanced by locus count)."""
    return [((n_loci * r) // world, (n_loci * (r + 1)) // world) for r in range(world)]


def shard_catalog(contig_locus_off, start, end, lo: int, hi: int):
    """Catalog view of loci [lo, hi): per-contig offsets rebuilt for the slice."""
    off = np.asarray(contig_locus_off, dtype=np.int64)
    s_off = np.clip(off, lo, hi) - lo
    return s_off.astype(np.int64), np.ascontiguousarray(start[lo:hi]), np.ascontiguousarray(end[lo:hi])


def reads_for_shard(contig, ref_start, ref_end, s_off, s_start, s_end) -> np.ndarray:
    """Mask of the reads htslib's fetch could return for any locus of the shard:
    pos < max(end)+10 and endpos > min(start)-10 on the shard's contigs (call.rs:285-288)."""
    n_contigs = len(s_off) - 1
    lo = np.full(n_contigs, np.iinfo(np.int64).max, np.int64)
    hi = np.full(n_contigs, np.iinfo(np.int64).min, np.int64)
    for c in range(n_contigs):
        a, b = int(s_off[c]), int(s_off[c + 1])
        if b > a:
            lo[c] = int(s_start[a:b].min()) - 10
            hi[c] = int(s_end[a:b].max()) + 10
    contig = np.asarray(contig)
    ok = (contig >= 0) & (contig < n_contigs)
    cc = np.where(ok, contig, 0)
    return ok & (np.asarray(ref_start, np.int64) < hi[cc]) & (np.asarray(ref_end, np.int64) > lo[cc])


def take_reads(reads, mask):
    """Subset of an SoA read set (any object with the SoA attributes) as a dict of contiguous arrays."""
    idx = np.flatnonzero(mask)
    n_cig = (reads.cigar_off[1:] - reads.cigar_off[:-1]).astype(np.int64)
    off = np.zeros(len(idx) + 1, np.uint64)
    off[1:] = np.cumsum(n_cig[idx])
    starts = reads.cigar_off[:-1].astype(np.int64)[idx]
    if len(idx):
        # gather the CIGAR words of the kept reads
        rep = np.repeat(starts - off[:-1].astype(np.int64), n_cig[idx])
        cigar = reads.cigar[np.arange(int(off[-1]), dtype=np.int64) + rep]
    else:
        cigar = np.zeros(0, np.uint32)
    return dict(contig=reads.contig[idx], ref_start=reads.ref_start[idx], ref_end=reads.ref_end[idx],
                mapq=reads.mapq[idx], hp=reads.hp[idx], flags=reads.flags[idx], cigar_off=off,
                cigar=np.ascontiguousarray(cigar, dtype=np.uint32))


def concat_ordered(parts):
    """Host-side ordered concatenation of per-rank outputs (rank order == catalog order)."""
    return np.concatenate([np.asarray(p) for p in parts]) if parts else np.zeros(0)
