#!/usr/bin/env python
"""Measures the cohort `outlier` kernels (include/inqcohort.h) on a synthetic combined matrix:
rows = loci, cols = 2 x samples. Prints one JSON line per method with the kernel time (CUDA events,
H2D excluded), the algorithmic bytes (4 B per value, read once) per second, and the oracle's CPU time on
a sample of the rows. Not the headline benchmark (bench.py stays on `call`)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--samples", type=int, default=268)       # the cohort size quoted in the reference's README
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cutoff", type=float, default=3.0)
    ap.add_argument("--methods", default="zscore,dbscan")
    args = ap.parse_args()
    from inquistr_b200 import cohort
    from oracle import oracle as O
    rng = np.random.default_rng(9)
    cols = 2 * args.samples
    m = (np.round(rng.gamma(2.0, 15.0, (args.rows, cols)) * 2) / 2).astype(np.float32)
    m[rng.random(m.shape) < 0.05] = np.nan
    big = rng.random(args.rows) < 0.2
    m[big, rng.integers(0, cols, big.sum())] = rng.integers(150, 3000, big.sum())
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for method in args.methods.split(","):
        best = None
        for _ in range(args.reps):
            kept, hr, hc, ms = cohort.outlier(m, 10, args.cutoff, method)
            best = ms if best is None else min(best, ms)
        n = min(args.rows, 2000)
        t0 = time.perf_counter()
        k2, f2, st = O.outlier_matrix(m[:n], 10, args.cutoff, method)
        cpu_s = time.perf_counter() - t0
        er, ec = np.nonzero(f2)
        sel = hr < n
        ok = bool(np.array_equal(kept[:n], k2) and np.array_equal(hr[sel], er) and np.array_equal(hc[sel], ec))
        gbps = m.nbytes / (best * 1e-3) / 1e9
        print(json.dumps({"method": method, "rows": args.rows, "cols": cols, "kernel_ms": best, "algorithmic_GBps": gbps,
                          "frac_of_measured_hbm": gbps / peak, "rows_per_s": args.rows / (best * 1e-3), "outliers": int(len(hr)),
                          "rows_kept": int(kept.sum()), "cpu_oracle_rows_per_s_1core": n / cpu_s,
                          "parity_first_rows": {"rows": n, "bit_exact_vs_oracle": ok}}))


if __name__ == "__main__":
    main()
