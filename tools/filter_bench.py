#!/usr/bin/env python
"""`polypolish filter` then `polypolish polish` (the CLI binary, one process each) on synthetic paired SAM, timed per phase (the library's own stderr line), device text path vs host.
usage: python tools/filter_bench.py [contig_len] [depth] [outdir (default /dev/shm)]"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
from polypolish_b200 import api  # noqa: E402

clen = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
depth = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
base = sys.argv[3] if len(sys.argv) > 3 else ("/dev/shm" if os.path.isdir("/dev/shm") else None)
d = tempfile.mkdtemp(prefix="pp_fb_", dir=base)
try:
    syn = api.Synth(seed=2, n_contigs=1, contig_len=clen, depth=depth)
    fa, sams = syn.write(d)
    exe = os.path.join(ROOT, "build", "polypolish")
    for extra in ([], ["--host-parse"]):
        for rep in range(2):
            t0 = time.perf_counter()
            p = subprocess.run([exe, "filter", "--in1", sams[0], "--in2", sams[1], "--out1", os.path.join(d, "o1.sam"), "--out2", os.path.join(d, "o2.sam")] + extra,
                               capture_output=True, text=True)
            dt = time.perf_counter() - t0
            last = [ln for ln in p.stderr.strip().split("\n") if "device" in ln or "phases" in ln]
            print(f"{'host' if extra else 'device'} text path, process wall {dt * 1e3:.0f} ms rc={p.returncode}: {' | '.join(last) if last else p.stderr[-200:]}", flush=True)
    # the whole `polypolish polish` process on the filtered files (CUDA start-up included), both parsers
    for extra in ([], ["--host-parse"]):
        for rep in range(2):
            t0 = time.perf_counter()
            p = subprocess.run([exe, "polish", fa, os.path.join(d, "o1.sam"), os.path.join(d, "o2.sam")] + extra, capture_output=True)
            dt = time.perf_counter() - t0
            err = p.stderr.decode(errors="replace")
            last = [ln for ln in err.strip().split("\n") if "GPU job" in ln or "tokeniser" in ln]
            print(f"polish, {'host' if extra else 'device'} parser, process wall {dt * 1e3:.0f} ms rc={p.returncode} out={len(p.stdout)} B: {' | '.join(last)}", flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
