"""GPU BGZF inflate prototype (include/inqbgzf.h): every block byte-identical to zlib's output, CRC32 and ISIZE of the
BGZF trailer verified, on synthetic BAMs and on hand-made blocks that force stored / fixed / dynamic deflate blocks,
long codes, long matches and distance-1 runs."""
import struct
import zlib

import numpy as np
import pytest


def bgzf_block(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 26
    assert bsize <= 65536
    return (struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize - 1) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def test_bgzf_binding_loads_without_gpu():
    from inquistr_b200 import bgzf
    rows, crcs, total = bgzf.scan_blocks(bgzf_block(b"hello world" * 10) + bgzf_block(b""))
    assert len(rows) == 2 and total == 110 and rows["out_len"].tolist() == [110, 0]
    L = bgzf._lib()
    for name in bgzf.EXPORTS:
        assert hasattr(L, name)


def check_image(image: bytes):
    from inquistr_b200 import bgzf
    rows, crcs, total = bgzf.scan_blocks(image)
    out, status, ms = bgzf.inflate(image, rows, total)
    assert np.all(status == 0), (np.flatnonzero(status)[:10], status[status != 0][:10])
    for r, crc in zip(rows, crcs):
        a, n = int(r["out_off"]), int(r["out_len"])
        ref = zlib.decompress(image[int(r["in_off"]):int(r["in_off"]) + int(r["in_len"])], -15)
        got = out[a:a + n].tobytes()
        assert got == ref and (zlib.crc32(got) & 0xFFFFFFFF) == int(crc)
    return ms


@pytest.mark.gpu
def test_handmade_blocks_match_zlib():
    rng = np.random.default_rng(5)
    payloads = [
        b"", b"a", b"ab" * 3, bytes(range(256)) * 40,
        b"A" * 65280,                                                  # distance-1 run, 258-byte matches
        rng.integers(0, 256, 65280, dtype=np.uint8).tobytes(),         # incompressible: stored blocks at any level
        rng.integers(0, 4, 65280, dtype=np.uint8).tobytes(),           # 2-bit alphabet: short codes, many short matches
        (b"ACGT" * 7 + b"N") * 2200,                                   # period 29
        rng.integers(0, 256, 300, dtype=np.uint8).tobytes() * 200,     # long-distance matches
        np.repeat(rng.integers(0, 256, 700, dtype=np.uint8), rng.integers(1, 90, 700)).tobytes()[:65000],
        bytes(rng.choice(np.arange(256, dtype=np.uint8), 60000, p=np.r_[0.5, 0.2, np.full(254, 0.3 / 254)])),   # skewed: long codes
    ]
    image = b""
    for p in payloads:
        for level, strat in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                             (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            image += bgzf_block(p, level, strat)
    check_image(image)


@pytest.mark.gpu
def test_synthetic_bam_matches_zlib(tmp_path):
    from synth import synth as S
    w = S.make_workload(3, scale=0.002, threads=4)
    for with_seq in (False, True):
        path = str(tmp_path / f"s{int(with_seq)}.bam")
        S.write_bam(w, path, with_seq=with_seq)
        check_image(open(path, "rb").read())


@pytest.mark.gpu
def test_corrupt_blocks_are_reported_not_crashed():
    from inquistr_b200 import bgzf
    rng = np.random.default_rng(7)
    good = bgzf_block(bytes(rng.integers(0, 8, 50000, dtype=np.uint8)), 6)
    bad = bytearray(good)
    for k in range(40, 400, 7):
        bad[k] ^= 0x5A                                                 # scramble the Huffman tables / stream
    image = good + bytes(bad) + good
    rows, crcs, total = bgzf.scan_blocks(image)
    out, status, _ = bgzf.inflate(image, rows, total)
    assert status[0] == 0 and status[2] == 0
    n = int(rows["out_len"][0])
    assert out[:n].tobytes() == out[2 * n:3 * n].tobytes()
    ok_mid = status[1] == 0 and (zlib.crc32(out[n:2 * n].tobytes()) & 0xFFFFFFFF) == int(crcs[1])
    assert status[1] != 0 or not ok_mid                                # either declined, or caught by the CRC check
