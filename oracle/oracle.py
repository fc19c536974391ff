"""ctypes front end for oracle/liboracle.so plus a pure-Python mirror for tiny cases.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
PARITY UNPINNED for `call`: no runnable reference and no golden `call` output exist.
The `outlier` rows are pinned by the reference's unit-test vectors (outlier.rs:147-168).

The pure-Python functions (`py_*`) restate the same reference lines a second
time, independently of the C file, so the two can be cross-checked:
  call.rs:377-413 -> py_call_from_cigar
  call.rs:497-522 -> py_median_str_length
  call.rs:279-374 -> py_genotype_locus
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_OK = 0
ORC_PANIC_BAD_HP = -1
ORC_PANIC_MEDIAN_EMPTY = -2
ORC_PANIC_START_LT_10 = -3
ORC_PANIC_BAD_INTERVAL = -4
ORC_PANIC_BAD_SA = -5

OPS = "MIDNSHP=X"


class _Reads(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64),
        ("contig", C.c_void_p),
        ("ref_start", C.c_void_p),
        ("ref_end", C.c_void_p),
        ("mapq", C.c_void_p),
        ("hp", C.c_void_p),
        ("flags", C.c_void_p),
        ("cigar_off", C.c_void_p),
        ("cigar", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_cohort.c", "oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_call_from_cigar.restype = C.c_int64
        L.orc_call_from_cigar.argtypes = [C.c_int32, C.c_void_p, C.c_uint64, C.c_uint32,
                                          C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
        L.orc_median_str_length.restype = C.c_double
        L.orc_median_str_length.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_int)]
        L.orc_cigar_to_rlen.restype = C.c_int64
        L.orc_cigar_to_rlen.argtypes = [C.c_char_p]
        L.orc_is_accidental_2d.restype = C.c_int
        L.orc_is_accidental_2d.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_int64]
        L.orc_genotype_loci.restype = C.c_int
        L.orc_genotype_loci.argtypes = [C.POINTER(_Reads), C.c_int32, C.c_uint64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_human_compare.restype = C.c_int
        L.orc_human_compare.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_format_f64.restype = C.c_int
        L.orc_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
        L.orc_validate_interval.restype = C.c_int
        L.orc_validate_interval.argtypes = [C.c_int64, C.c_int64, C.c_int64]
        L.orc_repeat_lengths.restype = C.c_int
        L.orc_repeat_lengths.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.orc_std_deviation_and_mean.restype = None
        L.orc_std_deviation_and_mean.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_zscore_outliers.restype = None
        L.orc_zscore_outliers.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
        L.orc_mode.restype = C.c_int64
        L.orc_mode.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_dbscan_outliers.restype = C.c_int
        L.orc_dbscan_outliers.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------- helpers
def pack_cigar(text: str) -> np.ndarray:
    """'100M10I' -> BAM packed u32 words (len<<4 | op)."""
    out, num = [], ""
    for ch in text:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | OPS.index(ch))
            num = ""
    return np.asarray(out, dtype=np.uint32)


def cigar_ref_len(words: np.ndarray) -> int:
    """htslib bam_cigar2rlen: M,D,N,=,X consume the reference."""
    ops = words & 0xF
    lens = (words >> 4).astype(np.int64)
    consume = np.isin(ops, [0, 2, 3, 7, 8])
    return int(lens[consume].sum())


class Reads:
    """SoA read set in the layout both the oracle and the C ABI consume."""

    def __init__(self, contig, ref_start, ref_end, mapq, hp, flags, cigar_off, cigar):
        self.contig = np.ascontiguousarray(contig, dtype=np.int32)
        self.ref_start = np.ascontiguousarray(ref_start, dtype=np.int32)
        self.ref_end = np.ascontiguousarray(ref_end, dtype=np.int32)
        self.mapq = np.ascontiguousarray(mapq, dtype=np.uint8)
        self.hp = np.ascontiguousarray(hp, dtype=np.uint8)
        self.flags = np.ascontiguousarray(flags, dtype=np.uint8)
        self.cigar_off = np.ascontiguousarray(cigar_off, dtype=np.uint64)
        self.cigar = np.ascontiguousarray(cigar, dtype=np.uint32)
        n = len(self.contig)
        assert len(self.cigar_off) == n + 1
        for a in (self.ref_start, self.ref_end, self.mapq, self.hp, self.flags):
            assert len(a) == n

    @property
    def n(self) -> int:
        return len(self.contig)

    @classmethod
    def from_records(cls, recs):
        """recs: iterable of dicts(contig,pos,cigar(text),mapq,hp(None|int),is2d,unmapped)."""
        contig, rs, re_, mq, hp, fl, off, words = [], [], [], [], [], [], [0], []
        for r in recs:
            w = pack_cigar(r["cigar"])
            rlen = 0 if r.get("unmapped") else cigar_ref_len(w)
            if rlen == 0:
                rlen = 1  # htslib bam_endpos
            contig.append(r.get("contig", 0))
            rs.append(r["pos"])
            re_.append(r["pos"] + rlen)
            mq.append(r.get("mapq", 60))
            h = r.get("hp", None)
            hp.append(0xFF if h is None else h)
            fl.append(1 if r.get("is2d") else 0)
            words.append(w)
            off.append(off[-1] + len(w))
        cig = np.concatenate(words) if words else np.zeros(0, np.uint32)
        return cls(contig, rs, re_, mq, hp, fl, off, cig)

    def _struct(self):
        return _reads_struct(self)


_DTYPES = dict(contig=np.int32, ref_start=np.int32, ref_end=np.int32, mapq=np.uint8, hp=np.uint8,
               flags=np.uint8, cigar_off=np.uint64, cigar=np.uint32)


def _reads_struct(reads):
    """C view of any object carrying the SoA attributes (oracle.Reads, synth.ReadSet)."""
    s = _Reads()
    s.n_reads = len(reads.contig)
    for name, dt in _DTYPES.items():
        a = getattr(reads, name)
        assert a.dtype == dt and a.flags["C_CONTIGUOUS"], name
        setattr(s, name, a.ctypes.data)
    return s


# --------------------------------------------------------------------------- C oracle
def call_from_cigar(ref_start, words, minlen, start_ext, end_ext, is2d=False):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    clip = C.c_int(0)
    v = lib().orc_call_from_cigar(int(ref_start), words.ctypes.data, len(words), int(minlen),
                                  int(start_ext), int(end_ext), int(bool(is2d)), C.byref(clip))
    return int(v), bool(clip.value)


def median_str_length(values, clips, support):
    v = np.ascontiguousarray(values, dtype=np.int64)
    k = np.ascontiguousarray(clips, dtype=np.uint8)
    p = C.c_int(0)
    r = lib().orc_median_str_length(v.ctypes.data, k.ctypes.data, len(v), int(support), C.byref(p))
    return float(r), bool(p.value)


def cigar_to_rlen(text: str) -> int:
    return int(lib().orc_cigar_to_rlen(text.encode()))


def is_accidental_2d(is_reverse, sa, ref_start, ref_end) -> bool:
    return bool(lib().orc_is_accidental_2d(int(bool(is_reverse)),
                                           None if sa is None else sa.encode(),
                                           int(ref_start), int(ref_end)))


def genotype_loci(reads: Reads, n_contigs, locus_contig, locus_start, locus_end, minlen=5,
                  support=3, unphased=False, threads=1):
    """-> (rc, phase1[f64], phase2[f64], op_visits). Loci in any order; output in input order."""
    lc = np.ascontiguousarray(locus_contig, dtype=np.int32)
    ls = np.ascontiguousarray(locus_start, dtype=np.uint32)
    le = np.ascontiguousarray(locus_end, dtype=np.uint32)
    n = len(lc)
    p1 = np.full(n, np.nan, dtype=np.float64)
    p2 = np.full(n, np.nan, dtype=np.float64)
    visits = C.c_uint64(0)
    s = _reads_struct(reads)
    rc = lib().orc_genotype_loci(C.byref(s), int(n_contigs), n, lc.ctypes.data, ls.ctypes.data,
                                 le.ctypes.data, int(minlen), int(support), int(bool(unphased)),
                                 int(threads), p1.ctypes.data, p2.ctypes.data, C.byref(visits))
    return int(rc), p1, p2, int(visits.value)


def human_compare(a: str, b: str) -> int:
    return int(lib().orc_human_compare(a.encode(), b.encode()))


def format_f64(v: float) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_format_f64(float(v), buf, 64)
    return buf.value.decode()


def validate_interval(start, end, chrom_len) -> int:
    return int(lib().orc_validate_interval(int(start), int(end), int(chrom_len)))


def format_row(chrom: str, start: int, end: int, p1: float, p2: float) -> str:
    """call.rs:57-65"""
    return f"{chrom}\t{start}\t{end}\t{format_f64(p1)}\t{format_f64(p2)}"


# --------------------------------------------------------------------------- pure-Python mirror
def py_call_from_cigar(ref_start, words, minlen, start_ext, end_ext, is2d=False):
    """call.rs:377-413, second independent restatement (u32 cursor)."""
    M32 = 0xFFFFFFFF
    pos = (ref_start + 1) & M32
    total, clipped = 0, False
    for w in words:
        w = int(w)
        op, ln = w & 0xF, w >> 4
        inside = start_ext < pos < end_ext
        if op in (0, 7, 8, 3):
            pos = (pos + ln) & M32
        elif op == 2:
            if ln > minlen and inside:
                total -= ln
            pos = (pos + ln) & M32
        elif op == 1:
            if ln > minlen and inside:
                total += ln
        elif op == 4:
            if (not is2d) and ln > minlen and inside:
                total += ln
                clipped = True
    return total, clipped


def py_median_str_length(calls, support):
    """call.rs:497-522. calls: list of (value, is_clip). Returns float or raises IndexError."""
    if len(calls) < support:
        return math.nan
    spanning = [v for v, c in calls if not c]
    clipped = [v for v, c in calls if c]
    if len(spanning) <= support:
        clipped.sort(key=lambda k: -k)
        spanning.extend(clipped[0:support - len(spanning)])
    spanning.sort()
    n = len(spanning)
    if n == 0:
        raise IndexError("median of empty vector (reference panics)")
    if n % 2 == 0:
        return float(spanning[n // 2 - 1] + spanning[n // 2]) / 2.0
    return float(spanning[n // 2])


def py_genotype_locus(reads: Reads, tid, start, end, minlen=5, support=3, unphased=False):
    """call.rs:279-374 for one locus, brute force over all reads (file order)."""
    if start < 10:
        raise OverflowError("start - 10 underflows u32")
    start_ext, end_ext = start - 10, end + 10
    buckets = {0: [], 1: [], 2: []}
    for r in range(reads.n):
        if reads.contig[r] != tid:
            continue
        rs, re_ = int(reads.ref_start[r]), int(reads.ref_end[r])
        if not (rs < end_ext and re_ > start_ext):  # htslib fetch
            continue
        mq, hp = int(reads.mapq[r]), int(reads.hp[r])
        if unphased:
            if start_ext < rs or re_ < end_ext or mq <= 10:
                continue
        else:
            if hp == 0xFF or (start_ext < rs and re_ < end_ext) or mq <= 10:
                continue
        a, b = int(reads.cigar_off[r]), int(reads.cigar_off[r + 1])
        if reads.flags[r] & 2:
            raise ValueError("is_accidental_2d panics on this read's SA tag (call.rs:431,439-450)")
        call = py_call_from_cigar(rs, reads.cigar[a:b], minlen, start_ext, end_ext,
                                  bool(reads.flags[r] & 1))
        if unphased:
            buckets[0].append(call)
        else:
            if hp not in (0, 1, 2):
                raise KeyError("HP not in {0,1,2} (reference panics)")
            buckets[hp].append(call)
    if unphased:
        calls = sorted(buckets[0], key=lambda c: (c[0], c[1]))  # tie rule: Span before Clip
        half = len(calls) // 2
        return py_median_str_length(calls[:half], support), py_median_str_length(calls[half:], support)
    return py_median_str_length(buckets[1], support), py_median_str_length(buckets[2], support)


# --------------------------------------------------------------------------- cohort `outlier` rows (outlier.rs)
def repeat_lengths(row, minsize: int):
    """outlier.rs:75-97 -> (kept, cleaned f32 values)"""
    v = np.ascontiguousarray(row, dtype=np.float32)
    out = np.empty_like(v)
    kept = lib().orc_repeat_lengths(v.ctypes.data, len(v), int(minsize), out.ctypes.data)
    return bool(kept), out


def std_deviation_and_mean(values):
    v = np.ascontiguousarray(values, dtype=np.float32)
    m, sd = C.c_float(), C.c_float()
    lib().orc_std_deviation_and_mean(v.ctypes.data, len(v), C.byref(m), C.byref(sd))
    return np.float32(m.value), np.float32(sd.value)


def zscore_outliers(values, cutoff: float) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.float32)
    flag = np.zeros(len(v), np.uint8)
    lib().orc_zscore_outliers(v.ctypes.data, len(v), float(cutoff), flag.ctypes.data)
    return flag


def mode(values) -> int:
    v = np.ascontiguousarray(values, dtype=np.float32)
    return int(lib().orc_mode(v.ctypes.data, len(v)))


def dbscan_outliers(values, mincluster: int):
    """-> (rc, noise flags); rc = -1 when the row has no positive value (the reference panics)"""
    v = np.ascontiguousarray(values, dtype=np.float32)
    flag = np.zeros(len(v), np.uint8)
    rc = lib().orc_dbscan_outliers(v.ctypes.data, len(v), int(mincluster), flag.ctypes.data)
    return rc, flag


def outlier_matrix(matrix, minsize=10, cutoff=3.0, method="zscore"):
    """Row loop of outlier.rs:33-72 over a rows x cols f32 matrix.
    -> (kept[rows] u8, flags[rows, cols] u8, status[rows] i8: -1 where dbscan has no mode)"""
    m = np.ascontiguousarray(matrix, dtype=np.float32)
    rows, cols = m.shape
    kept = np.zeros(rows, np.uint8)
    flags = np.zeros((rows, cols), np.uint8)
    status = np.zeros(rows, np.int8)
    mincluster = int(np.log2(cols)) if cols else 0          # samples.len().ilog2(), outlier.rs:39
    for r in range(rows):
        k, v = repeat_lengths(m[r], minsize)
        kept[r] = k
        if not k:
            continue
        if method == "zscore":
            flags[r] = zscore_outliers(v, cutoff)
        else:
            rc, f = dbscan_outliers(v, mincluster)
            status[r] = rc
            if rc == 0:
                flags[r] = f
    return kept, flags, status


def py_zscore_outliers(values, cutoff):
    """independent pure-Python mirror of outlier.rs:18-31,99-113 (numpy float32 scalars, sequential)"""
    v = [np.float32(x) for x in values]
    s = np.float32(0.0)
    for x in v:
        s = np.float32(s + x)
    n = np.float32(len(v))
    with np.errstate(all="ignore"):
        mean = np.float32(s / n)
        var = np.float32(0.0)
        for x in v:
            d = np.float32(mean - x)
            var = np.float32(var + np.float32(d * d))
        sd = np.float32(np.sqrt(np.float32(var / n)))
        return np.array([1 if np.float32(np.float32(x - mean) / sd) >= np.float32(cutoff) else 0 for x in v], np.uint8)


def py_dbscan_outliers(values, mincluster):
    """order-independent characterisation: noise = not core and no core point within eps"""
    v = np.asarray(values, dtype=np.float64)
    pos = [int(x) for x in values if x > 0]
    if not pos:
        return -1, np.zeros(len(v), np.uint8)
    cnt = {}
    for x in pos:
        cnt[x] = cnt.get(x, 0) + 1
    best = max(cnt.values())
    md = min(k for k, c in cnt.items() if c == best)
    eps = float(max(2 * md, 10))
    d = np.abs(v[:, None] - v[None, :])
    core = (d < eps).sum(1) >= mincluster
    near_core = ((d < eps) & core[None, :]).any(1)
    return 0, (~core & ~near_core).astype(np.uint8)
