#!/usr/bin/env python
"""End-to-end check of the C++ host: synthetic BAM + BED -> `inquistr-b200 call` -> TSV, timed, and
compared byte for byte with the TSV the oracle implies. Usage:
  python tools/bench_bam.py [--config 3] [--scale 0.02] [--with-seq] [-t 8] [--keep DIR]"""
import argparse, functools, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
from synth import synth as S

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--scale", type=float, default=0.02)
ap.add_argument("--with-seq", action="store_true")
ap.add_argument("-t", "--threads", type=int, default=8)
ap.add_argument("--keep", default=None)
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--devices", default="0", help="passed to `inquistr-b200 call --devices`")
ap.add_argument("--gen-threads", type=int, default=0)
a = ap.parse_args()

from inquistr_b200 import build
cli = build.build_cli()
w = S.make_workload(a.config, scale=a.scale, threads=a.gen_threads)
d = a.keep or tempfile.mkdtemp(prefix="inqbam")
os.makedirs(d, exist_ok=True)
bam, bed = os.path.join(d, "sample.bam"), os.path.join(d, "loci.bed")
t0 = time.perf_counter()
nbytes = S.write_bam(w, bam, with_seq=a.with_seq, threads=a.gen_threads)
t_write = time.perf_counter() - t0
rows = S.write_bed(w, bed, shuffle_seed=1)
stats = os.path.join(d, "stats.json")
args = [cli, "call", "-R", bed, "-t", str(a.threads), "--devices", a.devices, "--stats-json", stats] + (["-u"] if w.unphased else []) + [bam]
t0 = time.perf_counter()
r = subprocess.run(args, capture_output=True)
wall = time.perf_counter() - t0
assert r.returncode == 0, r.stderr.decode()[-2000:]
st = json.load(open(stats))
ok = None
if not a.no_check:
    sel = np.asarray([i for *_, i in rows])
    rc, p1, p2, _ = O.genotype_loci(w.reads, w.n_contigs, w.locus_contig[sel], w.locus_start[sel].astype(np.uint32),
                                    w.locus_end[sel].astype(np.uint32), w.minlen, w.support, w.unphased, threads=os.cpu_count())
    idx = list(range(len(rows)))
    if a.threads > 1:
        def cmp(x, y):
            c = O.human_compare(rows[x][0], rows[y][0])
            return c if c else (rows[x][1] > rows[y][1]) - (rows[x][1] < rows[y][1])
        idx.sort(key=functools.cmp_to_key(cmp))
    exp = "chromosome\tbegin\tend\tsample_H1\tsample_H2\n" + "".join(
        O.format_row(rows[i][0], rows[i][1], rows[i][2], p1[i], p2[i]) + "\n" for i in idx)
    ok = r.stdout == exp.encode()
scan_s = max(st["s_bam_scan"], 1e-9)
print(json.dumps({"workload": w.name, "with_seq": a.with_seq, "bam_bytes": nbytes, "bam_write_s": round(t_write, 2),
                  "devices": a.devices, "cli_wall_s": round(wall, 3), "loci": w.n_loci, "reads": w.reads.n, "loci_per_s": w.n_loci / wall,
                  "inflated_GB": st["bytes_inflated"] / 1e9, "inflate_GBps": st["bytes_inflated"] / 1e9 / wall,
                  "inflate_GBps_during_scan": st["bytes_inflated"] / 1e9 / scan_s,
                  "phases_s": {"cuda_ctx_create+set_loci (under the scan)": st["s_ctx_create_set_loci"],
                               "bam_scan (read + inflate + parse + route; pushes overlap it)": st["s_bam_scan"],
                               "h2d_push (worker threads, under the scan)": st["s_push_under_scan"],
                               "flush + genotype + join": st["s_flush_genotype"],
                               "tsv + exit": max(0.0, st["s_total"] - st["s_bam_scan"] - st["s_flush_genotype"]),
                               "total_in_process": st["s_total"]},
                  "gpu_busy_ms": st["ms_total"], "gpu_idle_s": max(0.0, st["s_total"] - st["ms_total"] / 1e3 - st["ms_h2d"] / 1e3),
                  "tsv_identical_to_oracle": ok, "cli_stats": st}))
if not a.keep:
    import shutil
    shutil.rmtree(d, ignore_errors=True)
