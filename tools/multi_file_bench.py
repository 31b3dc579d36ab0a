#!/usr/bin/env python
"""The file-level multi-GPU path timed on the GPUs of this box: `polypolish polish --gpus N` = pp_polish_files_multi.

  python tools/multi_file_bench.py [n_contigs] [contig_bp]

One config-5-shaped assembly (repeat families that cross contigs) as SAM text in /dev/shm; the same command on 1 GPU, on every GPU
with device-side shards (every GPU tokenises the text itself, parser 0) and with the host packer + host sharder (parser 1).  All
outputs must be byte-identical.  Prints one JSON line."""
import hashlib
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import polypolish_b200 as pp  # noqa: E402
from polypolish_b200 import api  # noqa: E402


def main():
    n_contigs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    clen = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    import __graft_entry__ as g
    g.build()
    n_gpu = torch.cuda.device_count()
    d = tempfile.mkdtemp(prefix="pp_multi_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        syn = api.Synth(seed=5, n_contigs=n_contigs, contig_len=clen, depth=100, cross_contig=0.03)
        fa, sams = syn.write(d)
        text = sum(os.path.getsize(s) for s in sams)
        out = {"assembly_bp": n_contigs * clen, "sam_text_bytes": text, "gpus": n_gpu, "runs": {}}
        ref = None

        def timed(name, fn, reps=3):
            nonlocal ref
            ts = []
            for _ in range(reps):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = fn()
                ts.append((time.perf_counter() - t0) * 1e3)
            h = hashlib.sha256(r).hexdigest()
            if ref is None:
                ref = h
            out["runs"][name] = {"ms_best": round(min(ts), 1), "ms_all": [round(t, 1) for t in ts], "same_fasta": h == ref}

        ctxs = [pp.Context(i) for i in range(n_gpu)]          # one context per GPU, reused by every run (a resident host process)
        timed("1_gpu_pp_polish_files", lambda: ctxs[0].polish_files(fa, sams))
        timed("%d_gpus_device_shards" % n_gpu, lambda: api.polish_files_multi(fa, sams, contexts=ctxs, parser=0), reps=4)
        timed("%d_gpus_host_packer_host_sharder" % n_gpu, lambda: api.polish_files_multi(fa, sams, contexts=ctxs, parser=1), reps=2)
        ctxs[0].set_parser(0)
        os.environ["POLYPOLISH_TIMING"] = "1"
        api.polish_files_multi(fa, sams, contexts=ctxs, parser=0, verbose=True)          # the library's own timing lines -> stderr
        os.environ.pop("POLYPOLISH_TIMING")
        if n_gpu == 1:
            four = [pp.Context(0) for _ in range(4)]
            timed("4_contexts_one_gpu_device_shards", lambda: api.polish_files_multi(fa, sams, contexts=four, parser=0), reps=3)
            for c in four:
                c.close()
        for c in ctxs:
            c.close()
        out["all_identical"] = all(r["same_fasta"] for r in out["runs"].values())
        print(json.dumps(out))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
