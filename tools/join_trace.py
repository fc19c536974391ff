#!/usr/bin/env python
"""Experiments: when do the CTAs of k_join_ranges run relative to each other? (INQ_JOIN_TRACE=file python bench.py ...)
Prints a histogram of CTA start and end times (50 us bins from the first start) and CTA durations."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 3)
t0 = a[:, 0].min()
st = (a[:, 0] - t0).astype(np.float64) / 1e3
en = (a[:, 1] - t0).astype(np.float64) / 1e3
dur = en - st
print(f"{len(a)} CTAs, span {en.max():.1f} us, duration per CTA: median {np.median(dur):.1f} us, p10 {np.percentile(dur, 10):.1f}, p90 {np.percentile(dur, 90):.1f}, max {dur.max():.1f}")
bins = np.arange(0, en.max() + 50, 50)
hs, _ = np.histogram(st, bins)
he, _ = np.histogram(en, bins)
for i in range(len(bins) - 1):
    print(f"{bins[i]:7.0f} us  started {hs[i]:6d}  finished {he[i]:6d}")
print("SMs used:", len(np.unique(a[:, 2])), " resident CTAs per SM at t = span/2:",
      np.mean([np.sum((st[a[:, 2] == s] < en.max() / 2) & (en[a[:, 2] == s] > en.max() / 2)) for s in np.unique(a[:, 2])]))
