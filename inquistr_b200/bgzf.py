"""ctypes binding of include/inqbgzf.h (GPU BGZF inflate prototype, part of libinqcall.so). No CPU fallback: blocks
the kernel declines are reported in `status`, inflating them elsewhere is the caller's business."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from .api import INQ_OK, InqError, load_library

EXPORTS = ["inq_bgzf_inflate", "inq_bgzf_last_error", "inq_bgzf_engine_create", "inq_bgzf_engine_destroy", "inq_bgzf_engine_run",
           "inq_host_register", "inq_host_unregister"]
_BOUND = False

ZBLOCK = np.dtype([("in_off", np.uint64), ("out_off", np.uint64), ("in_len", np.uint32), ("out_len", np.uint32)])


def _lib():
    global _BOUND
    L = load_library()
    if not _BOUND:
        L.inq_bgzf_inflate.restype = C.c_int
        L.inq_bgzf_inflate.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p,
                                       C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.inq_bgzf_last_error.restype = C.c_char_p
        L.inq_bgzf_last_error.argtypes = []
        _BOUND = True
    return L


def scan_blocks(data: bytes | np.ndarray, max_blocks: int | None = None):
    """BGZF block table of a file image (SAM spec 4.1): -> (ZBLOCK array, crc32[], total output bytes)."""
    buf = memoryview(data)
    p, out_off = 0, 0
    rows, crcs = [], []
    n = len(buf)
    while p + 18 <= n and (max_blocks is None or len(rows) < max_blocks):
        if buf[p] != 31 or buf[p + 1] != 139:
            raise ValueError("not a BGZF block")
        xlen = struct.unpack_from("<H", buf, p + 10)[0]
        bsize = None
        q = p + 12
        while q + 4 <= p + 12 + xlen:
            si1, si2, slen = buf[q], buf[q + 1], struct.unpack_from("<H", buf, q + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", buf, q + 4)[0] + 1
            q += 4 + slen
        if bsize is None:
            raise ValueError("BGZF block without BC subfield")
        in_off = p + 12 + xlen
        in_len = bsize - 12 - xlen - 8
        crc, isize = struct.unpack_from("<II", buf, p + bsize - 8)
        rows.append((in_off, out_off, in_len, isize))
        crcs.append(crc)
        out_off += isize
        p += bsize
    return np.array(rows, dtype=ZBLOCK), np.asarray(crcs, np.uint32), out_off


def inflate(comp, blocks, out_bytes: int, device: int = 0, out=None):
    """-> (out uint8[out_bytes], status uint32[n_blocks], dict(ms_h2d, ms_kernel, ms_d2h))"""
    comp = np.frombuffer(comp, dtype=np.uint8) if not isinstance(comp, np.ndarray) else comp
    blocks = np.ascontiguousarray(blocks, dtype=ZBLOCK)
    if out is None:
        out = np.empty(max(out_bytes, 1), np.uint8)
    status = np.full(len(blocks), 0xFFFFFFFF, np.uint32)
    a, b, c = C.c_float(0), C.c_float(0), C.c_float(0)
    L = _lib()
    rc = L.inq_bgzf_inflate(int(device), comp.ctypes.data, comp.nbytes, blocks.ctypes.data, len(blocks), out.ctypes.data, int(out_bytes),
                            status.ctypes.data, C.byref(a), C.byref(b), C.byref(c))
    if rc != INQ_OK:
        raise InqError(rc, L.inq_bgzf_last_error().decode())
    return out[:out_bytes], status, {"ms_h2d": a.value, "ms_kernel": b.value, "ms_d2h": c.value}
