"""Known-answer tests on the only reference-produced bytes in the checkout (/root/reference/test-data):
the htslib-written `small-test.bam.bai` pins the BAI reader (`BamIndexedReader::load_index` /
`query_chunks`), and `file{1,2,3}.inq[.gz]` -- the inputs of the reference's own
`combine` tests (combine.rs:61-78) -- pin `inquistr-b200 combine`. CPU only: no BAM, no GPU. The files
are read where they lie (nothing from /root/reference is copied into the repo); on a box without the
reference checkout these tests skip."""
import gzip
import json
import os
import struct
import subprocess

import pytest

REF = "/root/reference/test-data"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this box")


@pytest.fixture(scope="module")
def cli():
    from inquistr_b200 import build
    build.build_libinqcall()
    return build.build_cli()


# ---------------------------------------------------------------- independent BAI parse (SAM spec 5.2 / 5.3)
def parse_bai(path):
    b = open(path, "rb").read()
    assert b[:4] == b"BAI\x01"
    n_ref = struct.unpack_from("<i", b, 4)[0]
    p = 8
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", b, p)[0]; p += 4
        bins, meta = {}, None
        for _ in range(n_bin):
            bin_, n_chunk = struct.unpack_from("<Ii", b, p); p += 8
            ch = [struct.unpack_from("<QQ", b, p + 16 * i) for i in range(n_chunk)]; p += 16 * n_chunk
            if bin_ == 37450:
                meta = ch
            else:
                bins[bin_] = ch
        n_intv = struct.unpack_from("<i", b, p)[0]; p += 4
        lin = list(struct.unpack_from(f"<{n_intv}Q", b, p)); p += 8 * n_intv
        refs.append((bins, lin, meta))
    n_no_coor = struct.unpack_from("<Q", b, p)[0] if p + 8 <= len(b) else 0
    return refs, n_no_coor


def reg2bins(beg, end):
    end -= 1
    out = [0]
    for shift, base in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        out += list(range(base + (beg >> shift), base + (end >> shift) + 1))
    return out


def expected_chunks(ref, beg, end):
    bins, lin, _ = ref
    min_off = (lin[beg >> 14] if (beg >> 14) < len(lin) else lin[-1]) if lin else 0
    ch = sorted(c for b in reg2bins(beg, end) for c in bins.get(b, []) if c[1] > min_off)
    merged = []
    for a, e in ch:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([a, e])
    return merged


def test_reference_bai_loads_and_answers_the_test_bed_window(cli):
    bai = os.path.join(REF, "small-test.bam.bai")
    chrom, s, e = open(os.path.join(REF, "test.bed")).read().split()[:3]
    assert (chrom, int(s), int(e)) == ("chr7", 154778571, 154779363)        # test-data/test.bed:1
    beg, end = int(s) - 10, int(e) + 10                                     # the fetch window of call.rs:285-288
    refs, n_no_coor = parse_bai(bai)
    populated = [t for t, r in enumerate(refs) if r[0]]
    assert len(refs) == 195 and populated == [6]                            # hg38 order: chr7 is tid 6
    r = subprocess.run([cli, "baistat", bai, f"6:{beg}-{end}"], capture_output=True, timeout=60)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    assert got["refs"] == 195 and got["n_no_coor"] == n_no_coor
    assert got["mapped"] == {"6": [8105, 0]}                                # metadata pseudo-bin: 8,105 mapped reads
    assert refs[6][2][1] == (8105, 0)
    exp = expected_chunks(refs[6], beg, end)
    assert got["chunks"] == exp and len(exp) > 0
    # every chunk lies inside the (absent) 73.7 MB BAM: below the largest virtual offset the index mentions
    max_voff = max(c[1] for ch in refs[6][0].values() for c in ch)
    assert 73_000_000 < (max_voff >> 16) < 74_500_000
    assert all(a < b <= max_voff for a, b in got["chunks"])
    # a window on an unpopulated reference, and one far from any read, yield nothing
    r = subprocess.run([cli, "baistat", bai, "3:1000-2000"], capture_output=True, timeout=60)
    assert json.loads(r.stdout)["chunks"] == []
    r = subprocess.run([cli, "baistat", bai, "6:1000-2000"], capture_output=True, timeout=60)
    assert json.loads(r.stdout)["chunks"] == expected_chunks(refs[6], 1000, 2000)


def test_bai_rejects_garbage(cli, tmp_path):
    p = tmp_path / "x.bai"
    p.write_bytes(b"BAM\x01" + b"\0" * 32)
    assert subprocess.run([cli, "baistat", str(p)], capture_output=True).returncode == 1
    good = open(os.path.join(REF, "small-test.bam.bai"), "rb").read()
    p.write_bytes(good[:5000])                                               # truncated
    assert subprocess.run([cli, "baistat", str(p)], capture_output=True).returncode == 1


# ---------------------------------------------------------------- combine on the reference's own fixtures
def py_combine(texts):
    """combine.rs:40-58: the first file's line in full, then columns 4.. (split on tab) of every other file"""
    lines = [t.split("\n")[:-1] if t.endswith("\n") else t.split("\n") for t in texts]
    out = []
    for i, first in enumerate(lines[0]):
        row = [first]
        for other in lines[1:]:
            row += other[i].split("\t")[3:]
        out.append("\t".join(row))
    return "\n".join(out) + "\n"


def test_combine_on_reference_fixtures_plain_and_gz(cli):
    plain = [os.path.join(REF, f"file{i}.inq") for i in (1, 2, 3)]          # combine.rs:61-68
    gz = [p + ".gz" for p in plain]                                          # combine.rs:70-78
    texts = [open(p).read() for p in plain]
    assert "4027.0  4081.0" in texts[0].split("\n")[0]                       # file1.inq:1 has spaces where a tab should be
    a = subprocess.run([cli, "combine", *plain], capture_output=True, timeout=60)
    b = subprocess.run([cli, "combine", *gz], capture_output=True, timeout=60)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    assert a.stdout.decode() == py_combine(texts)
    gz_texts = [gzip.open(p, "rt").read() for p in gz]
    assert b.stdout.decode() == py_combine(gz_texts)
    if gz_texts == texts:
        assert a.stdout == b.stdout
    # mixed plain / gz and a different first file
    c = subprocess.run([cli, "combine", gz[1], plain[0], gz[2]], capture_output=True, timeout=60)
    assert c.returncode == 0 and c.stdout.decode() == py_combine([gz_texts[1], texts[0], gz_texts[2]])
