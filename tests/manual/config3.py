#!/usr/bin/env python
"""BASELINE config 3 at full size: `polypolish filter` (insert size) then `polypolish polish` on 5 Mbp x 100x synthetic
paired SAM, through the file-level C ABI calls, timed, and compared byte for byte with the CPU oracle on the same files.
usage: python tests/manual/config3.py [contig_len] [depth] [out.json]      (writes nothing else; temp files under /dev/shm)"""
import hashlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import polypolish_b200 as pp  # noqa: E402
from polypolish_b200 import api  # noqa: E402
import oracle_lib  # noqa: E402

clen = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
depth = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
out_json = sys.argv[3] if len(sys.argv) > 3 else None
with_oracle = os.environ.get("NO_ORACLE") != "1"


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 24), b""):
            h.update(b)
    return h.hexdigest()


base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 8 << 30 else None
d = tempfile.mkdtemp(prefix="pp_c3_", dir=base)
res = {"config": f"{clen} bp x {depth:g}x, filter then polish", "host_cores": os.cpu_count()}
try:
    syn = api.Synth(seed=2, n_contigs=1, contig_len=clen, depth=depth)
    fa, sams = syn.write(d)
    res["sam_text_bytes"] = sum(os.path.getsize(s) for s in sams)
    f1, f2 = os.path.join(d, "f1.sam"), os.path.join(d, "f2.sam")
    with pp.Context(0) as ctx:
        ts = []
        for rep in range(2):
            t0 = time.perf_counter()
            ctx.filter_files(sams[0], sams[1], f1, f2)
            ts.append((time.perf_counter() - t0) * 1e3)
        res["filter_ms"] = min(ts)
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            fasta = ctx.polish_files(fa, [f1, f2])
            ts.append((time.perf_counter() - t0) * 1e3)
        res["polish_ms"] = min(ts)
    res["filter_then_polish_mbp_s"] = clen / 1e6 / ((res["filter_ms"] + res["polish_ms"]) / 1e3)
    n_fail = 0
    for p in (f1, f2):
        with open(p, "rb") as f:
            n_fail += f.read().count(b"\tZP:Z:fail")
    res["zp_fail_records"] = n_fail
    res["gpu_sha256"] = {"f1": sha(f1), "f2": sha(f2), "fasta": hashlib.sha256(fasta).hexdigest()}
    if with_oracle:
        orc = oracle_lib.load()
        o1, o2 = os.path.join(d, "o1.sam"), os.path.join(d, "o2.sam")
        t0 = time.perf_counter()
        fo = orc.filter(sams[0], sams[1])
        res["oracle_filter_s"] = time.perf_counter() - t0
        open(o1, "wb").write(fo["out1"])
        open(o2, "wb").write(fo["out2"])
        t0 = time.perf_counter()
        po = orc.polish(fa, [o1, o2])
        res["oracle_polish_s"] = time.perf_counter() - t0
        res["oracle_thresholds"] = {"low": fo["low"], "high": fo["high"], "orientation": fo["orientation"]}
        res["identical"] = {"f1": sha(o1) == res["gpu_sha256"]["f1"], "f2": sha(o2) == res["gpu_sha256"]["f2"],
                            "fasta": hashlib.sha256(po["fasta"]).hexdigest() == res["gpu_sha256"]["fasta"]}
        res["oracle_filter_then_polish_mbp_s"] = clen / 1e6 / (res["oracle_filter_s"] + res["oracle_polish_s"])
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(res, indent=1))
if out_json:
    json.dump(res, open(out_json, "w"), indent=1)
