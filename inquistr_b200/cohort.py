"""ctypes binding of include/inqcohort.h (cohort `outlier` rows, part of libinqcall.so). No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import INQ_OK, InqError, load_library

ZSCORE, DBSCAN = 0, 1
INQ_ERR_NO_MODE, INQ_ERR_HITS_CAP = -15, -16
EXPORTS = ["inq_outlier", "inq_cohort_last_error"]

_BOUND = False


def _lib():
    global _BOUND
    L = load_library()
    if not _BOUND:
        L.inq_outlier.restype = C.c_int
        L.inq_outlier.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float,
                                  C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]
        L.inq_cohort_last_error.restype = C.c_char_p
        L.inq_cohort_last_error.argtypes = []
        _BOUND = True
    return L


def outlier(matrix, minsize: int = 10, cutoff: float = 3.0, method: str = "zscore", device: int = 0):
    """Outliers of every row of a rows x cols f32 matrix (outlier.rs:41-71).
    -> (row_kept[rows] u8, rows_of_hits, cols_of_hits, kernel milliseconds); hits sorted by (row, col)."""
    m = np.ascontiguousarray(matrix, dtype=np.float32)
    rows, cols = m.shape
    meth = {"zscore": ZSCORE, "dbscan": DBSCAN}[method]
    kept = np.zeros(rows, np.uint8)
    n = C.c_uint64(0)
    ms = C.c_float(0)
    cap = max(1024, rows // 4)
    L = _lib()
    while True:
        hits = np.zeros(cap, np.uint64)
        rc = L.inq_outlier(device, meth, rows, cols, m.ctypes.data, int(minsize), float(cutoff), kept.ctypes.data,
                           C.byref(n), hits.ctypes.data, cap, C.byref(ms))
        if rc == INQ_ERR_HITS_CAP:
            cap = int(n.value)
            continue
        if rc != INQ_OK:
            raise InqError(rc, L.inq_cohort_last_error().decode())
        h = hits[:int(n.value)]
        return kept, (h >> np.uint64(32)).astype(np.int64), (h & np.uint64(0xFFFFFFFF)).astype(np.int64), float(ms.value)
