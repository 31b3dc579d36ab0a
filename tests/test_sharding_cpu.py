"""Host-side logic of the multi-GPU path, on CPU: the contig sharder (no GPU needed) and a world_size-2 gloo run of
the per-rank shard selection + max-over-ranks reduction bench.py uses."""
import os
import subprocess
import sys

import numpy as np
import pytest

import polypolish_b200 as pp
from polypolish_b200 import api
from tests import fuzzgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GHOST = 0x20


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    g.build()


def shard_case(tmp_path, seed, n_shards):
    case = fuzzgen.make_case(seed, n_contigs=3, multimap=0.6)
    fa, sams = case.write(tmp_path)
    f = pp.load_fasta(fa)
    p = pp.pack_sams(f, sams)
    return f, p, api.Shards(f.view, p.view, n_shards)


@pytest.mark.parametrize("seed,n_shards", [(1, 2), (2, 3), (3, 2), (4, 8)])
def test_shards_partition(tmp_path, seed, n_shards):
    f, p, sh = shard_case(tmp_path, seed, n_shards)
    full = p.arrays()
    n_contigs = f.view.n_contigs
    seen_contigs, home_total = [], 0
    home_keys = []
    for s in range(n_shards):
        c, a, cmap, n_home = sh.get(s)
        seen_contigs += cmap
        arr = api.view_arrays(a)
        ghost = (arr["flags"] & GHOST) != 0
        assert int((~ghost).sum()) == n_home
        home_total += n_home
        off = np.ctypeslib.as_array(__import__("ctypes").cast(c.off, __import__("ctypes").POINTER(__import__("ctypes").c_uint64)),
                                    shape=(c.n_contigs + 1,))
        # contig bases are the original contigs, in input order within the shard
        assert cmap == sorted(cmap)
        for lc, oc in enumerate(cmap):
            got = bytes(np.ctypeslib.as_array(__import__("ctypes").cast(c.bases, __import__("ctypes").POINTER(__import__("ctypes").c_uint8)),
                                              shape=(int(off[-1]),))[int(off[lc]):int(off[lc + 1])])
            assert got == f.sequence(oc)
        # read groups stay whole and consecutive; ids dense and non-decreasing
        rid = arr["read_id"]
        assert (np.diff(rid.astype(np.int64)) >= 0).all() and (len(rid) == 0 or rid[-1] + 1 == arr["n_reads"])
        # home alignments keep (original contig, start, nm, cigar) in SAM order
        for i in np.nonzero(~ghost)[0]:
            oc = cmap[arr["contig"][i]] if arr["contig"][i] != 0xFFFFFFFF else 0xFFFFFFFF
            ops = tuple(arr["cigar_ops"][arr["cigar_off"][i]:arr["cigar_off"][i] + arr["n_cigar"][i]].tolist())
            home_keys.append((oc, int(arr["ref_start"][i]), int(arr["nm"][i]), ops, int(arr["flags"][i]) & 0x1F))
    assert sorted(seen_contigs) == list(range(n_contigs))
    assert home_total == len(full["contig"])
    ref_keys = []
    for i in range(len(full["contig"])):
        ops = tuple(full["cigar_ops"][full["cigar_off"][i]:full["cigar_off"][i] + full["n_cigar"][i]].tolist())
        ref_keys.append((int(full["contig"][i]), int(full["ref_start"][i]), int(full["nm"][i]), ops, int(full["flags"][i]) & 0x1F))
    assert sorted(home_keys) == sorted(ref_keys)


def test_ghosts_keep_groups_whole(tmp_path):
    f, p, sh = shard_case(tmp_path, 7, 3)
    full = p.arrays()
    sizes = np.bincount(full["read_id"])
    for s in range(3):
        c, a, cmap, n_home = sh.get(s)
        arr = api.view_arrays(a)
        if len(arr["read_id"]) == 0:
            continue
        local = np.bincount(arr["read_id"])
        # every group present in a shard has all its records there (home + ghost)
        assert set(local.tolist()) <= set(sizes.tolist())


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import numpy as np
import polypolish_b200 as pp
from polypolish_b200 import api
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
syn = api.Synth(seed=5, n_contigs=4, contig_len=20000, depth=20)
f = syn.fasta(); p = syn.pack(f)
sh = api.Shards(f.view, p.view, world)
c, a, cmap, n_home = sh.get(rank)                  # this rank's contigs and alignments
t = torch.tensor([float(n_home), float(c.n_contigs), 10.0 + rank], dtype=torch.float64)
tot = t.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)     # bench.py: time = max over ranks
if rank == 0:
    assert int(tot[0]) == p.view.n_aln, (tot, p.view.n_aln)
    assert int(tot[1]) == 4
    assert mx[2] == 10.0 + world - 1
    print("OK", int(tot[0]))
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


def model_shards(f, full, n_shards, seq_bits):
    """Python restatement of the sharding rule (shard.cpp header): LPT bin packing of whole contigs on contig length + aligned bases,
    every read group copied - in SAM order - to each shard one of its records lands on, foreign records as ghosts, one copy of
    a sequence per (shard, group)."""
    nc = f.view.n_contigs
    blk = 16 if seq_bits == 4 else 32
    weight = [int(f.off[i + 1] - f.off[i]) for i in range(nc)]
    for c, L in zip(full["contig"].tolist(), full["seq_len"].tolist()):
        if c != 0xFFFFFFFF and c < nc:
            weight[c] += L
    order = sorted(range(nc), key=lambda i: -weight[i])              # stable, like std::stable_sort
    load, shard_of = [0] * n_shards, [0] * nc
    for ci in order:
        b = min(range(n_shards), key=lambda s: load[s])
        shard_of[ci] = b
        load[b] += weight[ci]
    cmap = [[c for c in range(nc) if shard_of[c] == s] for s in range(n_shards)]
    local = {c: cmap[shard_of[c]].index(c) for c in range(nc)}
    out = [dict(contig=[], ref_start=[], read_id=[], seq_off=[], seq_len=[], cigar_off=[], n_cigar=[], nm=[], flags=[], cigar_ops=[],
                pool=bytearray(), n_reads=0) for _ in range(n_shards)]
    n = len(full["contig"])
    rid = full["read_id"]
    g0 = 0
    while g0 < n:
        g1 = g0 + 1
        while g1 < n and rid[g1] == rid[g0]:
            g1 += 1
        sh_of = [0 if (full["contig"][i] == 0xFFFFFFFF or full["contig"][i] >= nc) else shard_of[int(full["contig"][i])] for i in range(g0, g1)]
        for s in sorted(set(sh_of)):
            o = out[s]
            r = o["n_reads"]
            o["n_reads"] += 1
            seen = {}
            for i, so in zip(range(g0, g1), sh_of):
                fl = int(full["flags"][i])
                c = int(full["contig"][i])
                home = so == s
                soff = 0
                if home:
                    if not fl & 0x10:                                   # PP_FLAG_NOSEQ
                        key = int(full["seq_off"][i])
                        if key in seen and fl & 0x04:                    # PP_FLAG_SEQSTAR shares the group's copy
                            soff = seen[key]
                        else:
                            nb = (int(full["seq_len"][i]) + 31) // 32
                            soff = len(o["pool"]) // blk
                            o["pool"] += bytes(full["seq_pool"][key * blk:(key + nb) * blk])
                            seen.setdefault(key, soff)
                else:
                    fl |= GHOST
                unknown = c == 0xFFFFFFFF or c >= nc
                o["contig"].append((0xFFFFFFFF if unknown else local[c]) if home else 0)
                o["ref_start"].append(int(full["ref_start"][i])); o["read_id"].append(r); o["seq_off"].append(soff)
                o["seq_len"].append(int(full["seq_len"][i])); o["cigar_off"].append(len(o["cigar_ops"])); o["n_cigar"].append(int(full["n_cigar"][i]))
                o["nm"].append(int(full["nm"][i])); o["flags"].append(fl)
                co = int(full["cigar_off"][i])
                o["cigar_ops"] += full["cigar_ops"][co:co + int(full["n_cigar"][i])].tolist()
        g0 = g1
    return cmap, out


@pytest.mark.parametrize("seed,n_shards", [(11, 2), (12, 3), (13, 4), (14, 7)])
def test_shards_equal_python_model(tmp_path, seed, n_shards):
    """Every array of every shard, against the restated rule (the sharder builds the shards on parallel threads)."""
    case = fuzzgen.make_case(seed, n_contigs=4, multimap=0.6, exotic=0.5 if seed == 14 else 0.0)
    fa, sams = case.write(tmp_path)
    f = pp.load_fasta(fa)
    p = pp.pack_sams(f, sams)
    full = p.arrays()
    cmap_m, out_m = model_shards(f, full, n_shards, int(full["seq_bits"]))
    sh = api.Shards(f.view, p.view, n_shards)
    for s in range(n_shards):
        c, a, cmap, n_home = sh.get(s)
        arr = api.view_arrays(a)
        m = out_m[s]
        assert cmap == cmap_m[s]
        for k in ("contig", "ref_start", "read_id", "seq_off", "seq_len", "cigar_off", "n_cigar", "nm", "flags", "cigar_ops"):
            assert arr[k].tolist() == m[k], (s, k)
        assert bytes(arr["seq_pool"]) == bytes(m["pool"])
        assert arr["n_reads"] == m["n_reads"]


def test_filtered_generation_equals_the_sharders_shard(built):
    """bench.py --gpus N: every rank generates only the reads of the ONE data set that have a record on its contigs
    (pp_synth_set_shard_filter, per-pair random streams) and runs the sharder on them.  That must be, array for array, the
    shard the sharder makes from the whole data set - ghost records included."""
    syn = api.Synth(seed=5, n_contigs=6, contig_len=40_000, depth=30, cross_contig=0.05)
    f = syn.fasta()
    full = syn.pack(f)
    n = 3
    assign = [c % n for c in range(6)]
    whole = api.Shards(f.view, full.view, n, shard_of_contig=assign)
    total_home = 0
    for s in range(n):
        syn.set_shard_filter(n, s, assign)
        part = syn.pack(f)
        mine = api.Shards(f.view, part.view, n, shard_of_contig=assign, only_shard=s)
        c1, a1, m1, h1 = whole.get(s)
        c2, a2, m2, h2 = mine.get(s)
        x, y = api.view_arrays(a1), api.view_arrays(a2)
        assert m1 == m2 and h1 == h2 and a1.n_aln == a2.n_aln and a1.n_aln > h1 > 0          # ghosts exist
        for k in x:
            if isinstance(x[k], np.ndarray):
                assert np.array_equal(x[k], y[k]), k
        assert part.view.n_aln < full.view.n_aln
        total_home += h1
    assert total_home == full.view.n_aln
    syn.set_shard_filter(0, 0)
    assert syn.pack(f).view.n_aln == full.view.n_aln


def test_threaded_generation_gives_the_same_bytes(built, tmp_path):
    """pp_synth_set_threads (bench.py --gpus N uses it to shorten its setup): blocks of pairs generated by worker threads and
    emitted in pair order are byte for byte the sequential generator's text, with and without the shard filter."""
    import hashlib

    def digests(threads, filt):
        syn = api.Synth(seed=9, n_contigs=5, contig_len=30_000, depth=40, cross_contig=0.05)
        if filt:
            syn.set_shard_filter(2, 1, [c % 2 for c in range(5)])
        syn.set_threads(threads)
        fa, sams = syn.write(str(tmp_path))
        out = [hashlib.sha256(open(p, "rb").read()).hexdigest() for p in sams]
        syn.close()
        return out
    for filt in (False, True):
        assert digests(1, filt) == digests(3, filt) == digests(0, filt)
