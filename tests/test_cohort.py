"""Cohort `outlier` rows: oracle known-answer tests from the reference's own unit
tests (outlier.rs:147-168) and GPU parity of inq_outlier (include/inqcohort.h) against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O

REF_VALUES = [1.0, 2.0, 2.0, 3.0, 1.0, 5.0, 3.0, 2.0, 2.0, 1.0, 120.0]      # outlier.rs:149,161


# ------------------------------------------------------------------ oracle, pinned by the reference's tests
def test_reference_kat_zscore():
    # outlier.rs:159-168: expected ["s11"] at cutoff 2.0
    for f in (O.zscore_outliers, O.py_zscore_outliers):
        assert list(np.flatnonzero(f(REF_VALUES, 2.0))) == [10]


def test_reference_kat_dbscan():
    # outlier.rs:147-157: mincluster = len.ilog2() = 3, expected ["s11"]
    mc = int(np.log2(len(REF_VALUES)))
    assert mc == 3 and O.mode(REF_VALUES) == 2
    for f in (O.dbscan_outliers, O.py_dbscan_outliers):
        rc, flag = f(REF_VALUES, mc)
        assert rc == 0 and list(np.flatnonzero(flag)) == [10]


def test_repeat_lengths_and_edge_semantics():
    kept, v = O.repeat_lengths([np.nan, 3.0, 9.5], 10)
    assert not kept and list(v) == [0.0, 3.0, 9.5]                           # NaN -> 0, max 9.5 < 10
    assert O.repeat_lengths([np.nan, 10.0], 10)[0]                            # `<` is strict (outlier.rs:91)
    assert not O.repeat_lengths([np.nan, np.nan], 10)[0]
    # identical values: sd = 0 -> 0/0 = NaN -> no outlier; one large value among zeros: +inf never happens
    assert O.zscore_outliers([12.0] * 6, 3.0).sum() == 0
    # only expansions are reported (outlier.rs:106-107)
    z = O.zscore_outliers([50.0] * 20 + [0.0], 2.0)
    assert z.sum() == 0
    # dbscan: no positive value -> the reference panics ("No mode found")
    assert O.dbscan_outliers([0.0, 0.0, -1.0], 1)[0] == -1
    # mode ignores zeros (outlier.rs:135-137) and truncates (as usize)
    assert O.mode([0, 0, 0, 7.5, 7.0, 3.0]) == 7
    # eps is at least 10: 25 is within 10 of the 16-18 cluster's core points -> Edge, 40 is Noise
    rc, f = O.dbscan_outliers([2, 2, 2, 16, 17, 18, 18, 25, 40], 3)
    assert rc == 0 and list(np.flatnonzero(f)) == [8]


def test_oracle_restatements_agree_on_random_rows():
    rng = np.random.default_rng(5)
    for _ in range(200):
        x = random_row(rng, int(rng.integers(2, 120)))
        assert np.array_equal(O.zscore_outliers(x, 2.5), O.py_zscore_outliers(x, 2.5))
        mc = int(np.log2(len(x)))
        a, b = O.dbscan_outliers(x, mc), O.py_dbscan_outliers(x, mc)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])


def random_row(rng, n):
    x = np.round(rng.gamma(2, 15, n) * 2) / 2
    x[rng.random(n) < 0.15] = np.nan
    if rng.random() < 0.4:
        x[rng.integers(0, n, max(1, n // 40))] = rng.integers(100, 3000)
    if rng.random() < 0.1:
        x[:] = np.nan                                                         # dropped by the minsize test
    return x.astype(np.float32)


def random_matrix(seed, rows, cols):
    rng = np.random.default_rng(seed)
    return np.stack([random_row(rng, cols) for _ in range(rows)])


# ------------------------------------------------------------------ GPU parity through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(1, 11), (300, 7), (1000, 64), (517, 536), (129, 1000), (70, 2100)])   # 2100 columns: tiled z-score kernel
@pytest.mark.parametrize("method", ["zscore", "dbscan"])
def test_gpu_outlier_matches_oracle(rows, cols, method):
    from inquistr_b200 import cohort
    m = random_matrix(rows * 1000 + cols, rows, cols)
    if rows == 1:
        m[0, :] = REF_VALUES
    for minsize, cutoff in ((10, 3.0), (1, 2.0), (200, 1.5)):
        kept, flags, status = O.outlier_matrix(m, minsize, cutoff, method)
        assert (status == 0).all()
        k, hr, hc, ms = cohort.outlier(m, minsize, cutoff, method)
        assert np.array_equal(k, kept)
        er, ec = np.nonzero(flags)
        assert np.array_equal(hr, er) and np.array_equal(hc, ec), (method, minsize, cutoff)


@pytest.mark.gpu
@pytest.mark.parametrize("cols", [3, 31, 33, 63, 64, 65, 100, 128, 129, 255, 257, 511, 513, 1023, 1024, 1025, 1500, 4096])
def test_gpu_dbscan_every_sort_engine(cols):
    """The dbscan kernel picks a register sort by padded width (2..32 keys per lane) or the shared-memory network
    (tiny and > 1024-column rows); every boundary against the oracle, with the inputs that stress the value-only
    sort: heavy ties, +-0, infinities, negative values and an outlier value that occurs in several columns."""
    from inquistr_b200 import cohort
    rng = np.random.default_rng(cols)
    rows = 48
    m = random_matrix(7000 + cols, rows, cols)
    m[1, :] = 30.0                                                            # one value only
    m[2, : cols // 2] = 0.0
    m[2, cols // 2:] = -0.0
    m[2, -1] = 40.0
    if cols >= 31:
        m[3, :] = 12.0
        m[3, [1, 5, cols - 1]] = 900.0                                        # the same noise value in three columns
        m[4, 0] = np.inf
        m[4, 1] = -np.inf
        m[5, : cols // 3] = -rng.integers(1, 50, cols // 3)
        m[6, :] = rng.integers(10, 14, cols)                                  # four distinct values
    kept, flags, status = O.outlier_matrix(m, 10, 3.0, "dbscan")
    assert (status == 0).all()
    k, hr, hc, _ = cohort.outlier(m, 10, 3.0, "dbscan")
    er, ec = np.nonzero(flags)
    assert np.array_equal(k, kept) and np.array_equal(hr, er) and np.array_equal(hc, ec)
    if cols >= 31:
        assert set(hc[hr == 3]) == {1, 5, cols - 1}


@pytest.mark.gpu
def test_gpu_outlier_reference_vectors_and_errors():
    from inquistr_b200 import cohort
    from inquistr_b200.api import InqError
    m = np.asarray([REF_VALUES], np.float32)
    for method in ("zscore", "dbscan"):
        k, hr, hc, _ = cohort.outlier(m, 10, 2.0, method)
        assert list(k) == [1] and list(hr) == [0] and list(hc) == [10]       # "s11"
    with pytest.raises(InqError) as e:
        cohort.outlier(np.zeros((3, 8), np.float32), 0, 3.0, "dbscan")        # kept (max 0 >= 0) but no mode
    assert e.value.code == cohort.INQ_ERR_NO_MODE
    k, hr, hc, _ = cohort.outlier(np.zeros((0, 8), np.float32), 10, 3.0, "zscore")
    assert len(k) == 0 and len(hr) == 0


# ------------------------------------------------------------------ CLI: `combine` (host only) and `outlier`
import gzip
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cli():
    from inquistr_b200 import build
    build.build_libinqcall()
    return build.build_cli()


def run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True, timeout=300)


def py_combine(texts):
    """combine.rs:27-59 restated: first file's lines in full, columns 4.. of every other file appended"""
    lines = [t.split("\n")[:-1] if t.endswith("\n") else t.split("\n") for t in texts]
    out = []
    for i, first in enumerate(lines[0]):
        row = [first]
        for other in lines[1:]:
            row += other[i].split("\t")[3:]
        out.append("\t".join(row))
    return "\n".join(out) + "\n"


def make_inq(rng, loci, name):
    rows = [f"chrom\tbegin\tend\t{name}_H1\t{name}_H2"]
    for c, b, e in loci:
        vals = []
        for _ in range(2):
            u = rng.random()
            vals.append("NaN" if u < 0.1 else O.format_f64(float(np.round(rng.gamma(2, 12) * 2) / 2) if u < 0.95 else float(rng.integers(300, 4000))))
        rows.append(f"{c}\t{b}\t{e}\t{vals[0]}\t{vals[1]}")
    return "\n".join(rows) + "\n"


@pytest.fixture(scope="module")
def cohort_files(tmp_path_factory):
    rng = np.random.default_rng(77)
    d = tmp_path_factory.mktemp("cohort")
    loci = [(f"chr{1 + i // 200}", 1000 + 37 * i, 1000 + 37 * i + 20) for i in range(600)]
    texts, paths = [], []
    for s in range(24):
        t = make_inq(rng, loci, f"sample{s}")
        p = str(d / f"s{s}.inq") + (".gz" if s % 3 == 1 else "")
        if p.endswith(".gz"):
            with gzip.open(p, "wt") as f:
                f.write(t)
        else:
            open(p, "w").write(t)
        texts.append(t)
        paths.append(p)
    return d, texts, paths


def test_cli_combine_plain_and_gz(cli, cohort_files):
    d, texts, paths = cohort_files
    r = run(cli, "combine", *paths)
    assert r.returncode == 0, r.stderr
    assert r.stdout.decode() == py_combine(texts)
    # a missing file panics before anything is printed (combine.rs:29-33)
    r = run(cli, "combine", paths[0], str(d / "missing.inq"))
    assert r.returncode == 101 and r.stdout == b"" and b"does not exist" in r.stderr
    # a shorter second file: unwrap on None (combine.rs:45)
    short = str(d / "short.inq")
    open(short, "w").write("\n".join(texts[1].split("\n")[:5]) + "\n")
    r = run(cli, "combine", paths[0], short)
    assert r.returncode == 101
    assert run(cli, "combine").returncode == 2


def test_cli_outlier_usage_errors(cli, cohort_files, tmp_path):
    d, texts, paths = cohort_files
    assert run(cli, "outlier").returncode == 2
    assert run(cli, "outlier", "--method", "kmeans", paths[0]).returncode == 2
    r = run(cli, "outlier", str(d / "nope.tsv"))
    assert r.returncode == 101 and b"Combined file does not exist!" in r.stderr           # main.rs:210-212
    r = run(cli, "outlier", "-s", "a", "-S", paths[0], paths[0])
    assert r.returncode == 101 and b"Cannot use both -s and -S" in r.stderr               # main.rs:214-216
    bad = tmp_path / "bad.tsv"
    bad.write_text("chrom\tbegin\tend\ta_H1\ta_H2\nchr1\t1\t2\t3.0\tx\n")
    r = run(cli, "outlier", str(bad))
    assert r.returncode == 101 and b"Failed to parse number" in r.stderr                  # outlier.rs:79
    r = run(cli, "outlier", "--help")
    assert r.returncode == 0 and b"[possible values: zscore, dbscan]" in r.stdout


def expected_outlier_tsv(combined_text, minsize, cutoff, method, subset=None):
    lines = combined_text.split("\n")[:-1]
    samples = lines[0].split("\t")[3:]
    names = [s.replace("_H1", "").replace("_H2", "") for s in samples]
    m = np.asarray([[np.float32(x) for x in l.split("\t")[3:]] for l in lines[1:]], np.float32)
    kept, flags, status = O.outlier_matrix(m, minsize, cutoff, method)
    out = ["chrom\tbegin\tend\toutliers"]
    for i, l in enumerate(lines[1:]):
        cols = np.flatnonzero(flags[i])
        if len(cols) == 0:
            continue
        ex = [names[c] for c in cols]
        if subset is not None and not any(e in subset for e in ex):
            continue
        out.append("\t".join(l.split("\t")[:3]) + "\t" + ",".join(ex))
    return "\n".join(out) + "\n"


@pytest.mark.gpu
def test_cli_outlier_tsv_matches_oracle(cli, cohort_files, tmp_path):
    d, texts, paths = cohort_files
    combined = py_combine(texts)
    cpath = tmp_path / "combined.tsv.gz"
    with gzip.open(cpath, "wt") as f:
        f.write(combined)
    for method, extra in (("zscore", []), ("zscore", ["-z", "2", "--minsize", "50"]), ("dbscan", []), ("dbscan", ["--minsize", "1"])):
        minsize = int(extra[extra.index("--minsize") + 1]) if "--minsize" in extra else 10
        cutoff = float(extra[extra.index("-z") + 1]) if "-z" in extra else 3.0
        r = run(cli, "outlier", "--method", method, *extra, str(cpath))
        assert r.returncode == 0, r.stderr
        assert r.stdout.decode() == expected_outlier_tsv(combined, minsize, cutoff, method)
    r = run(cli, "outlier", "-s", "sample3", str(cpath))
    assert r.returncode == 0 and r.stdout.decode() == expected_outlier_tsv(combined, 10, 3.0, "zscore", {"sample3"})
    sub = tmp_path / "subset.txt"
    sub.write_text("sample1\nsample20\n")
    r = run(cli, "outlier", "-S", str(sub), "--method", "dbscan", str(cpath))
    assert r.returncode == 0 and r.stdout.decode() == expected_outlier_tsv(combined, 10, 3.0, "dbscan", {"sample1", "sample20"})


@pytest.mark.gpu
def test_cli_outlier_prints_rows_before_a_panic(cli, tmp_path):
    """The reference writes each row as it is decided (outlier.rs:70-120), so what precedes the row it panics on is
    already on stdout; the batched CLI must flush those rows before it exits 101."""
    samples = [f"s{i}" for i in range(12)]
    head = "chrom\tbegin\tend\t" + "\t".join(f"{s}_H1\t{s}_H2" for s in samples)
    normal = ["20.0"] * 24
    good = list(normal)
    good[5] = "900.0"                                                     # s2_H2: a clear expansion
    good_row = "chr1\t100\t200\t" + "\t".join(good)
    bad = tmp_path / "bad_later.tsv"
    bad.write_text(head + "\n" + good_row + "\nchr1\t300\t400\t" + "\t".join(normal[:-1] + ["x"]) + "\n")
    r = run(cli, "outlier", str(bad))
    assert r.returncode == 101 and b"Failed to parse number" in r.stderr
    assert r.stdout.decode() == "chrom\tbegin\tend\toutliers\nchr1\t100\t200\ts2\n"
    nomode = tmp_path / "nomode.tsv"
    nomode.write_text(head + "\n" + good_row + "\nchr1\t300\t400\t" + "\t".join(["0.0"] * 24) + "\n" + good_row + "\n")
    r = run(cli, "outlier", "--method", "dbscan", "--minsize", "0", str(nomode))
    assert r.returncode == 101 and b"No mode found for repeat" in r.stderr  # outlier.rs:144
    assert r.stdout.decode() == "chrom\tbegin\tend\toutliers\nchr1\t100\t200\ts2\n"


def test_outlier_fails_loudly_without_gpu(cli, cohort_files, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from inquistr_b200 import cohort
    from inquistr_b200.api import InqError
    with pytest.raises(InqError) as e:
        cohort.outlier(np.asarray([REF_VALUES], np.float32), 10, 2.0, "zscore")
    assert e.value.code == -1                                            # INQ_ERR_CUDA: no CPU implementation behind the ABI
    d, texts, paths = cohort_files
    c = tmp_path / "c.tsv"
    c.write_text(py_combine(texts))
    r = run(cli, "outlier", str(c))
    assert r.returncode == 1 and r.stdout == b"chrom\tbegin\tend\toutliers\n" and b"inquistr-b200:" in r.stderr
