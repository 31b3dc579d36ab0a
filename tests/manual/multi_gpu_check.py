#!/usr/bin/env python
"""`polypolish polish --gpus N` (contigs sharded over N GPUs, SAM tokenised on the first one) against the CPU oracle.
usage: python tests/manual/multi_gpu_check.py [n_gpus] [n_contigs] [contig_len] [depth]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
from polypolish_b200 import api  # noqa: E402
import oracle_lib  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
clen = int(sys.argv[3]) if len(sys.argv) > 3 else 300_000
depth = float(sys.argv[4]) if len(sys.argv) > 4 else 60.0
with tempfile.TemporaryDirectory() as d:
    syn = api.Synth(seed=7, n_contigs=n_contigs, contig_len=clen, depth=depth)
    fa, sams = syn.write(d)
    t0 = time.perf_counter()
    exp = oracle_lib.load().polish(fa, sams)["fasta"]
    t_orc = time.perf_counter() - t0
    exe = os.path.join(ROOT, "build", "polypolish")
    for extra in ([], ["--host-parse"]):
        for gpus in (1, n_gpus):
            t0 = time.perf_counter()
            p = subprocess.run([exe, "polish", "--gpus", str(gpus), fa] + sams + extra, capture_output=True)
            dt = time.perf_counter() - t0
            ok = p.returncode == 0 and p.stdout == exp
            jobs = [ln for ln in p.stderr.decode(errors="replace").split("\n") if ln.startswith("GPU job")]
            print(f"--gpus {gpus} {' '.join(extra) or 'device parser'}: identical to the oracle = {ok}, process wall {dt * 1e3:.0f} ms (oracle {t_orc:.1f} s); {len(jobs)} GPU jobs", flush=True)
            if not ok:
                sys.stderr.write(p.stderr.decode(errors="replace")[-2000:])
                sys.exit(1)
print("multi-GPU check OK")
