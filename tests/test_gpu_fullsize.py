"""BASELINE.json configs 2-5 at (or near) full size, GPU against the CPU oracle, byte for byte.  Part of the default
`-m gpu` run: nothing here is opt-in.

  config 2  5 Mbp x 100x, `polish` from SAM text (device tokeniser) and from the packed arrays
  config 3  the same reads through `filter` (insert size) then `polish`: both filtered SAM files and the FASTA
  config 4  1000x depth (counter stress, long ordered-depth lists) on 1 Mbp
  config 5  a multi-contig assembly whose repeat families cross contigs (a read's k spans shards; the ghost-record rule
            of alignment.rs:283-288 under contig sharding), through pp_polish_files_multi over every visible GPU and
            with more contexts than GPUs

The synthetic files (1.2-2.4 GB of SAM text each) live in /dev/shm when it has room."""
import hashlib
import os
import shutil
import tempfile

import pytest

import polypolish_b200 as pp
from polypolish_b200 import api

pytestmark = pytest.mark.gpu


def _workdir(need_gb):
    shm = "/dev/shm"
    base = shm if os.path.isdir(shm) and shutil.disk_usage(shm).free > need_gb * (1 << 30) else None
    return tempfile.mkdtemp(prefix="pp_full_", dir=base)


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def _sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 24), b""):
            h.update(b)
    return h.hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g
    g.build()
    c = pp.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def config2():
    """5 Mbp x 100x, multi-mapped 150 bp pairs (bench.py's default workload, same seed)."""
    d = _workdir(6)
    syn = api.Synth(seed=2, contig_len=5_000_000, depth=100)
    fa, sams = syn.write(d)
    yield syn, fa, sams, d
    shutil.rmtree(d, ignore_errors=True)


def test_config2_full_size(ctx, oracle, config2):
    syn, fa, sams, d = config2
    exp = oracle.polish(fa, sams)
    assert sum(exp["changed"]) > 200
    ctx.set_parser(0)
    got = ctx.polish_files(fa, sams)
    assert _sha(got) == _sha(exp["fasta"])
    # the packed-array entry point (what bench.py times) on the same records
    f = syn.fasta()
    p = syn.pack(f)
    r = ctx.polish_packed(f.view, p.view)
    fasta = b"".join(b">" + f.names[i].encode() + b" " + f.descriptions[i].encode() + b" polypolish\n" + r["sequences"][i] + b"\n"
                     for i in range(len(f.names)))
    assert fasta == exp["fasta"]
    assert r["changed"] == exp["changed"] and r["zero_depth"] == exp["zero_depth"] and r["n_aln_used"] == exp["used_total"]


def test_config3_filter_then_polish_full_size(ctx, oracle, config2):
    syn, fa, sams, d = config2
    f1, f2 = os.path.join(d, "f1.sam"), os.path.join(d, "f2.sam")
    o1, o2 = os.path.join(d, "o1.sam"), os.path.join(d, "o2.sam")
    ctx.set_parser(0)
    ctx.filter_files(sams[0], sams[1], f1, f2)
    fo = oracle.filter(sams[0], sams[1])
    open(o1, "wb").write(fo["out1"])
    open(o2, "wb").write(fo["out2"])
    n_fail = fo["out1"].count(b"\tZP:Z:fail") + fo["out2"].count(b"\tZP:Z:fail")
    del fo
    assert n_fail > 1000
    assert _sha_file(f1) == _sha_file(o1) and _sha_file(f2) == _sha_file(o2)
    exp = oracle.polish(fa, [o1, o2])
    assert _sha(ctx.polish_files(fa, [f1, f2])) == _sha(exp["fasta"])
    # and as one call, without the 1.25 GB of intermediate files (pp_filter_polish_files)
    assert _sha(ctx.filter_polish_files(fa, sams[0], sams[1])) == _sha(exp["fasta"])
    for p in (f1, f2, o1, o2):
        os.remove(p)


def test_config4_depth_1000(ctx, oracle):
    d = _workdir(5)
    try:
        syn = api.Synth(seed=4, contig_len=1_000_000, depth=1000)
        fa, sams = syn.write(d)
        exp = oracle.polish(fa, sams)
        assert _sha(ctx.polish_files(fa, sams)) == _sha(exp["fasta"])
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_config5_contig_sharded_cross_contig_multimaps(ctx, oracle):
    """8 contigs x 1 Mbp x 100x; 3 % of the assembly in repeat families whose copies lie on different contigs."""
    import torch
    n_gpu = torch.cuda.device_count()
    d = _workdir(6)
    try:
        syn = api.Synth(seed=5, n_contigs=8, contig_len=1_000_000, depth=100, cross_contig=0.03)
        fa, sams = syn.write(d)
        exp = oracle.polish(fa, sams)
        want = _sha(exp["fasta"])
        assert _sha(ctx.polish_files(fa, sams)) == want                                     # one GPU, one call
        assert _sha(api.polish_files_multi(fa, sams, devices=list(range(n_gpu)))) == want   # every visible GPU
        assert _sha(api.polish_files_multi(fa, sams, devices=[i % n_gpu for i in range(8)])) == want   # 8 shards
        # the shards really carry foreign records: ghosts exist and every shard's k needs them
        f = syn.fasta()
        p = syn.pack(f)
        sh = api.Shards(f.view, p.view, 8)
        ghosts = 0
        for s in range(8):
            c, a, cmap, n_home = sh.get(s)
            ghosts += a.n_aln - n_home
        assert ghosts > 1000
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_depth_70000_counters_are_32_bit(ctx, oracle):
    """A 2 kbp contig at 70,000x (amplicon / plasmid depths): more than 65,535 alignments over every position.  The reference
    counts in u32 (pileup.rs:33-37) and so do the tile's shared-memory counters; one tile holds the whole contig (a list of
    ~950,000 alignments, far more queued reads than the queue holds, depth walks over lists of 10^5 entries)."""
    d = _workdir(2)
    try:
        syn = api.Synth(seed=7, contig_len=2_000, depth=70_000)
        fa, sams = syn.write(d)
        exp = oracle.polish(fa, sams)
        assert max(exp["total_depth"]) / 2000 > 60_000
        assert ctx.polish_files(fa, sams) == exp["fasta"]
        f = syn.fasta()
        p = syn.pack(f)
        r = ctx.polish_packed(f.view, p.view)
        assert r["changed"] == exp["changed"] and r["zero_depth"] == exp["zero_depth"]
    finally:
        shutil.rmtree(d, ignore_errors=True)
