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
