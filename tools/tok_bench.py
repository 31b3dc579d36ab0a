#!/usr/bin/env python
"""Device SAM tokeniser: upload and kernel time against the number of reader threads (5 Mbp x 100x synthetic SAM).
usage: python tools/tok_bench.py [contig_len] [depth] [readers,...]"""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
import polypolish_b200 as pp  # noqa: E402
from polypolish_b200 import api  # noqa: E402

clen = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
depth = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
readers = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 4, 8, 12, 16]
only = os.environ.get("TOK_ONCE") == "1"
syn = api.Synth(seed=2, n_contigs=1, contig_len=clen, depth=depth)
base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 4 << 30 else None
d = tempfile.mkdtemp(prefix="pp_tok_", dir=base)
try:
    fa, sams = syn.write(d)
    fasta = pp.load_fasta(fa)
    nbytes = sum(os.path.getsize(s) for s in sams)
    with pp.Context(0) as ctx:
        for r in readers:
            ctx.set_readers(r)
            best = None
            for rep in range(1 if only else 4):
                t0 = time.perf_counter()
                rc, st = ctx.tokenise(fasta, sams)
                wall = (time.perf_counter() - t0) * 1e3
                assert rc == 0
                cur = dict(readers=r, wall_ms=round(wall, 2), h2d_ms=[round(s["h2d_ms"], 2) for s in st],
                           device_ms=[round(s["device_ms"], 3) for s in st], text_gb_s=round(nbytes / 1e6 / sum(s["h2d_ms"] for s in st), 2))
                if best is None or cur["wall_ms"] < best["wall_ms"]:
                    best = cur
            print(json.dumps(best), flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
