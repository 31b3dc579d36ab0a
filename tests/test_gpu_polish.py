"""GPU parity tests: the CUDA polish path (through the C ABI) against the CPU oracle, byte for byte."""
import os

import numpy as np
import pytest

import polypolish_b200 as pp
from polypolish_b200 import api
from tests import fuzzgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g
    g.build()
    c = pp.Context(0)
    yield c
    c.close()


def both(ctx, oracle, fa, sams, **opts):
    """Runs oracle and GPU on the same files; returns (oracle_fasta or error msg, gpu_fasta or error msg)."""
    try:
        exp = ("ok", oracle.polish(fa, sams, **opts))
    except Exception as e:
        exp = ("err", e.msg)
    try:
        got = ("ok", ctx.polish_files(fa, sams, **opts))
    except pp.PolypolishError as e:
        got = ("err", e.msg)
    return exp, got


def assert_parity(ctx, oracle, fa, sams, **opts):
    exp, got = both(ctx, oracle, fa, sams, **opts)
    assert exp[0] == got[0], (exp[0], exp[1] if exp[0] == "err" else "", got[1] if got[0] == "err" else "")
    if exp[0] == "ok":
        assert got[1] == exp[1]["fasta"]
    return exp, got


# ---- the reference's vote vectors (pileup.rs:209-295), driven end to end -----------------------------------
REF = "ACGTTGCAAGCTTAGGCATCGATTACGGATCCATGCAAGTCCGATAGGCT"       # 50 bp, position 20 is the tested base
POS = 20
DUMP = "TTGACCGTAGCTAGGATCCGATCGGATTAGCCTAGGCTTAACGGATCGAT" * 2


def vote_case(tmp_path, orig, adds, k=1):
    """adds = list of alleles seen at POS (each one read).  k > 1: every read also has k-1 good alignments on a
    second contig, so each contributes 1/k of depth (pileup.rs test 05 uses 0.1)."""
    draft = REF[:POS] + orig + REF[POS + 1:]
    fa = tmp_path / "v.fasta"
    fa.write_text(f">t\n{draft}\n>dump\n{DUMP}\n")
    lines = []
    for i, al in enumerate(adds):
        left, right = draft[5:POS], draft[POS + 1:45]
        if al == "-":
            seq, cig = left + right, f"{len(left)}M1D{len(right)}M"
        else:
            seq = left + al + right
            cig = f"{len(left) + 1}M{len(al) - 1}I{len(right)}M" if len(al) > 1 else f"{len(seq)}M"
        lines.append(f"q{i}\t0\tt\t6\t60\t{cig}\t*\t0\t0\t{seq}\t*\tNM:i:1")
        for j in range(k - 1):
            lines.append(f"q{i}\t256\tdump\t{1 + j}\t0\t{len(seq)}M\t*\t0\t0\t*\t*\tNM:i:0")
    sam = tmp_path / "v.sam"
    sam.write_text("\n".join(lines) + "\n")
    return fa, [sam]


VOTE = [
    ("A", ["A"] * 50, 1, 0.2, "A"), ("G", ["A", "T"] + ["G"] * 50, 1, 0.2, "G"), ("T", ["C"] + ["A"] * 99, 1, 0.2, "A"),
    ("A", ["T", "C", "G"], 1, 0.2, "A"), ("C", ["A"] * 123 + ["T"] * 321, 10, 0.2, "C"),
    ("T", ["A"] * 6 + ["C"] * 4, 1, 0.2, "T"), ("T", ["A"] * 9 + ["C"], 1, 0.1, "T"), ("T", ["A"] * 19 + ["C"], 1, 0.1, "A"),
    ("T", ["-"] * 30, 1, 0.2, ""), ("T", ["TGA"] * 30 + ["T"] * 3, 1, 0.2, "TGA"), ("N", ["N"] * 20, 1, 0.2, "N"),
    ("N", ["C"] * 20 + ["N"], 1, 0.2, "C"), ("A", ["N"] * 20, 1, 0.2, "N"),
]


@pytest.mark.parametrize("orig,adds,k,fi,expect", VOTE)
def test_vote_vectors_end_to_end(ctx, oracle, tmp_path, orig, adds, k, fi, expect):
    fa, sams = vote_case(tmp_path, orig, adds, k)
    exp, got = assert_parity(ctx, oracle, fa, sams, fraction_invalid=fi)
    seq = got[1].split(b"\n")[1].decode()
    draft = REF[:POS] + orig + REF[POS + 1:]
    assert seq == draft[:POS] + expect + draft[POS + 1:]


def test_tiny_hand_derived(ctx, oracle, tmp_path):
    fa = tmp_path / "asm.fasta"
    fa.write_text(">c1 desc here\nACGTACGTACGTACGTACGT\n")
    read = "ACGTACGAACGTACGTACGT"
    sam = tmp_path / "r.sam"
    sam.write_text("@SQ\tSN:c1\tLN:20\n" + "".join(
        f"r{i}\t0\tc1\t1\t60\t20M\t*\t0\t0\t{read}\t{'I' * 20}\tNM:i:1\n" for i in range(6)))
    assert ctx.polish_files(fa, [sam]) == b">c1 desc here polypolish\nACGTACGAACGTACGTACGT\n"
    # zero SAM files: legal, draft unchanged (main.rs:107)
    assert ctx.polish_files(fa, []) == b">c1 desc here polypolish\nACGTACGTACGTACGTACGT\n"


@pytest.mark.parametrize("seed", range(100, 180))
def test_fuzz_parity(ctx, oracle, tmp_path, seed):
    case = fuzzgen.make_case(seed, exotic=0.5 if seed % 4 == 0 else 0.0)
    fa, sams = case.write(tmp_path)
    assert_parity(ctx, oracle, fa, sams, **case.opts)


@pytest.mark.parametrize("seed", range(300, 312))
def test_fuzz_parity_deep_multimap(ctx, oracle, tmp_path, seed):
    """More reads per position and most reads multi-mapped: non-dyadic k everywhere (ordered f64 depth)."""
    case = fuzzgen.make_case(seed, n_contigs=2, contig_len=(200, 400), depth=(150, 300), multimap=0.8,
                             opts=dict(careful=False))
    fa, sams = case.write(tmp_path)
    assert_parity(ctx, oracle, fa, sams, **case.opts)


@pytest.mark.parametrize("opts", [dict(), dict(careful=True), dict(min_depth=0), dict(max_errors=2),
                                  dict(fraction_invalid=0.05, fraction_valid=0.95)])
def test_synth_small_parity(ctx, oracle, tmp_path, opts):
    syn = api.Synth(seed=1, contig_len=50_000, depth=100)        # BASELINE config 1
    fa, sams = syn.write(tmp_path)
    exp, got = assert_parity(ctx, oracle, fa, sams, **opts)
    assert exp[0] == "ok"


def test_synth_packed_path_and_stats(ctx, oracle, tmp_path):
    """The streamed-pack route (no SAM text on disk) and the per-contig statistics."""
    syn = api.Synth(seed=3, n_contigs=3, contig_len=40_000, depth=80)
    fa, sams = syn.write(tmp_path)
    exp = oracle.polish(fa, sams)
    f = syn.fasta()
    p = syn.pack(f)
    r = ctx.polish_packed(f.view, p.view)
    fasta = b"".join(b">" + f.names[i].encode() + b" " + f.descriptions[i].encode() + b" polypolish\n" + r["sequences"][i] + b"\n"
                     for i in range(3))
    assert fasta == exp["fasta"]
    assert r["changed"] == exp["changed"] and r["zero_depth"] == exp["zero_depth"]
    assert r["n_aln_used"] == exp["used_total"]
    # resident route: same answer, repeatable, options can change between calls
    ctx.upload(f.view, p.view)
    r2 = ctx.polish_resident()
    assert r2["sequences"] == r["sequences"]
    r3 = ctx.polish_resident(careful=True)
    assert b"".join(b">" + f.names[i].encode() + b" " + f.descriptions[i].encode() + b" polypolish\n" + r3["sequences"][i] + b"\n"
                    for i in range(3)) == oracle.polish(fa, sams, careful=True)["fasta"]
    r4 = ctx.polish_resident()
    assert r4["sequences"] == r["sequences"]


def test_synth_medium_parity(ctx, oracle, tmp_path):
    syn = api.Synth(seed=2, contig_len=600_000, depth=100)
    fa, sams = syn.write(tmp_path)
    assert_parity(ctx, oracle, fa, sams)


def test_device_error_messages(ctx, oracle, tmp_path):
    fa = tmp_path / "a.fasta"
    fa.write_text(">c1\nACGTACGTACGTACGTACGTACGTACGTAC\n")
    cases = [
        "r1\t0\tnope\t1\t60\t10M\t*\t0\t0\tACGTACGTAC\t*\tNM:i:0\n",               # unknown reference (good alignment)
        "r1\t0\tc1\t1\t60\t10M\t*\t0\t0\tACGTACGTACG\t*\tNM:i:0\n",                # CIGAR / SEQ length mismatch
        "r1\t0\tc1\t1\t60\t4M2N4M\t*\t0\t0\tACGTACGT\t*\tNM:i:0\n",                # N inside
    ]
    for i, sam in enumerate(cases):
        s = tmp_path / f"e{i}.sam"
        s.write_text(sam)
        exp, got = both(ctx, oracle, fa, [s])
        assert exp[0] == "err" and got[0] == "err"
        assert got[1] == exp[1]
    # not an error when the alignment is not "good" (soft clip): unknown reference is never looked at
    s = tmp_path / "ok.sam"
    s.write_text("r1\t0\tnope\t1\t60\t2S8M\t*\t0\t0\tACGTACGTAC\t*\tNM:i:0\n")
    assert_parity(ctx, oracle, fa, [s])


def test_full_size_properties(ctx):
    """BASELINE config 2 size (5 Mbp x 100x): size-independent properties instead of the oracle:
    determinism across runs and agreement of the host-buffer and device-resident entry points."""
    syn = api.Synth(seed=2, n_contigs=2, contig_len=2_500_000, depth=100)
    f = syn.fasta()
    p = syn.pack(f)
    r1 = ctx.polish_packed(f.view, p.view)
    r2 = ctx.polish_packed(f.view, p.view)
    assert r1["sequences"] == r2["sequences"]
    ctx.upload(f.view, p.view)
    r3 = ctx.polish_resident()
    assert r3["sequences"] == r1["sequences"]
    assert sum(r1["changed"]) > 100


@pytest.mark.parametrize("seed,n_shards", [(20, 2), (21, 3), (22, 2), (23, 5)])
def test_contig_sharding_invariance(ctx, oracle, tmp_path, seed, n_shards):
    """polishing each contig shard on its own (with ghost records keeping k) == polishing everything."""
    case = fuzzgen.make_case(seed, n_contigs=3, multimap=0.6, opts=dict(careful=(seed == 22)))
    fa, sams = case.write(tmp_path)
    exp = oracle.polish(fa, sams, **case.opts)
    f = pp.load_fasta(fa)
    p = pp.pack_sams(f, sams, careful=case.opts["careful"])
    sh = api.Shards(f.view, p.view, n_shards)
    seqs, used = [None] * 3, 0
    for s in range(n_shards):
        c, a, cmap, n_home = sh.get(s)
        if c.n_contigs == 0:
            assert a.n_aln == 0
            continue
        r = ctx.polish_packed(c, a, **case.opts)
        for lc, oc in enumerate(cmap):
            seqs[oc] = r["sequences"][lc]
        used += r["n_aln_used"]
    fasta = b"".join(b">" + f.names[i].encode() + ((b" " + f.descriptions[i].encode()) if f.descriptions[i] else b"") +
                     b" polypolish\n" + seqs[i] + b"\n" for i in range(3))
    assert fasta == exp["fasta"]
    assert used == exp["used_total"]


def test_multi_context_file_path(oracle, tmp_path):
    """pp_polish_files_multi with two contexts (both on GPU 0 here): host threads, sharding, merge in input order."""
    case = fuzzgen.make_case(31, n_contigs=3, multimap=0.5, opts=dict(careful=False))
    fa, sams = case.write(tmp_path)
    exp = oracle.polish(fa, sams, **case.opts)["fasta"]
    assert api.polish_files_multi(fa, sams, devices=[0, 0], **case.opts) == exp
    assert api.polish_files_multi(fa, sams, devices=[0, 0, 0, 0], **case.opts) == exp
    syn = api.Synth(seed=9, n_contigs=4, contig_len=30_000, depth=50)
    d2 = tmp_path / "syn"
    d2.mkdir()
    fa2, sams2 = syn.write(d2)
    assert api.polish_files_multi(fa2, sams2, devices=[0, 0, 0]) == oracle.polish(fa2, sams2)["fasta"]


@pytest.mark.parametrize("seed", range(40, 52))
def test_device_side_shards_fuzz(oracle, tmp_path, seed):
    """Several contexts, no host in the middle: every context tokenises the text and keeps its own contigs (pp_tok_set_shard: foreign
    records become ghosts on the device).  Same FASTA as the oracle, as the host packer + host sharder, and the reference's errors."""
    case = fuzzgen.make_case(seed, n_contigs=2 + seed % 3, multimap=0.6, exotic=0.5 if seed % 4 == 0 else 0.0,
                             opts=dict(careful=(seed % 3 == 0)))
    fa, sams = case.write(tmp_path)
    try:
        exp = ("ok", oracle.polish(fa, sams, **case.opts)["fasta"])
    except Exception as e:
        exp = ("err", e.msg)
    for parser in (0, 1):
        for n in (2, 3):
            try:
                got = ("ok", api.polish_files_multi(fa, sams, devices=[0] * n, parser=parser, **case.opts))
            except pp.PolypolishError as e:
                got = ("err", e.msg)
            assert got == exp, (parser, n)


def test_device_side_shards_errors(oracle, tmp_path):
    """Data errors under device-side sharding are the reference's, raised once: an RNAME that is not in the assembly, a CIGAR that does
    not match its SEQ, on a record of the second shard."""
    fa = tmp_path / "a.fasta"
    fa.write_text(">c1\n" + "ACGTTGCAAGCTTAGGCATCGATTACGGATCCATGCAAGTCCGATAGGCT" * 2 + "\n>c2\n" + "TTGACCGTAGCTAGGATCCGATCGGATTAGCCTAGGCTTAACGGATCGAT" + "\n")
    good = "r1\t0\tc1\t1\t60\t20M\t*\t0\t0\tACGTTGCAAGCTTAGGCATC\t*\tNM:i:0"
    for bad in ("r2\t0\tnope\t1\t60\t10M\t*\t0\t0\tACGTTGCAAG\t*\tNM:i:0",
                "r2\t0\tc2\t1\t60\t12M\t*\t0\t0\tTTGACCGTAG\t*\tNM:i:0"):
        sam = tmp_path / "a.sam"
        sam.write_text(good + "\n" + bad + "\n")
        with pytest.raises(Exception) as e1:
            oracle.polish(str(fa), [str(sam)])
        with pytest.raises(pp.PolypolishError) as e2:
            api.polish_files_multi(str(fa), [str(sam)], devices=[0, 0])
        assert e2.value.msg == e1.value.msg


def test_cli_binary(oracle, tmp_path):
    """The drop-in command line: same flags as the reference, FASTA on stdout, Error + exit 1 on user errors."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "polypolish")
    syn = api.Synth(seed=4, contig_len=30_000, depth=40)
    fa, sams = syn.write(tmp_path)
    r = subprocess.run([exe, "polish", "-m", "8", "--min_depth", "4", fa] + sams, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == oracle.polish(fa, sams, max_errors=8, min_depth=4)["fasta"]
    r = subprocess.run([exe, "polish", "-i", "0.6", fa] + sams, capture_output=True)
    assert r.returncode == 1 and b"Error: --fraction_invalid must be less than --fraction_valid" in r.stderr
    r = subprocess.run([exe, "-V"], capture_output=True)
    assert r.stdout.strip() == b"Polypolish v0.6.1"
    o1, o2 = tmp_path / "f1.sam", tmp_path / "f2.sam"
    r = subprocess.run([exe, "filter", "--in1", sams[0], "--in2", sams[1], "--out1", str(o1), "--out2", str(o2)], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    ef = oracle.filter(sams[0], sams[1])
    assert open(o1, "rb").read() == ef["out1"] and open(o2, "rb").read() == ef["out2"]


@pytest.mark.parametrize("seed", [100, 101, 104, 105, 300, 303])
def test_debug_tsv_parity(ctx, oracle, tmp_path, seed):
    """--debug per-base TSV (polish.rs:230-266): depth, thresholds, sorted allele counts, status, new base."""
    kw = dict(n_contigs=2, contig_len=(200, 400), depth=(150, 300), multimap=0.8, opts=dict(careful=False)) if seed >= 300 else {}
    case = fuzzgen.make_case(seed, exotic=0.5 if seed % 4 == 0 else 0.0, **kw)
    fa, sams = case.write(tmp_path)
    exp = oracle.polish(fa, sams, debug=True, **case.opts)
    dbg = tmp_path / "debug.tsv"
    got = ctx.polish_files(fa, sams, debug=dbg, **case.opts)
    assert got == exp["fasta"]
    assert dbg.read_bytes() == exp["debug_tsv"]
    # recording is off again afterwards and the plain path still works
    assert ctx.polish_files(fa, sams, **case.opts) == exp["fasta"]


def test_debug_tsv_synth(ctx, oracle, tmp_path):
    syn = api.Synth(seed=6, n_contigs=2, contig_len=40_000, depth=60)
    fa, sams = syn.write(tmp_path)
    exp = oracle.polish(fa, sams, debug=True)
    dbg = tmp_path / "debug.tsv"
    assert ctx.polish_files(fa, sams, debug=dbg) == exp["fasta"]
    assert dbg.read_bytes() == exp["debug_tsv"]


def test_depth_1000_parity(ctx, oracle, tmp_path):
    """BASELINE config 4 in miniature (1000x depth: counter / atomic stress, long fix-up lists), against the oracle."""
    syn = api.Synth(seed=4, contig_len=200_000, depth=1000)
    fa, sams = syn.write(tmp_path)
    exp = oracle.polish(fa, sams)
    assert ctx.polish_files(fa, sams) == exp["fasta"]


def test_cli_clap_argument_forms_and_log_numbers(oracle, tmp_path):
    """clap accepts --name=value, -m5 and `--`; the stderr log carries the reference's per-contig numbers
    (polish.rs:206-227: mean read depth, zero-depth bp, changed positions, estimated accuracy)."""
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "polypolish")
    syn = api.Synth(seed=8, n_contigs=2, contig_len=30_000, depth=40)
    fa, sams = syn.write(tmp_path)
    exp = oracle.polish(fa, sams, max_errors=5, min_depth=4, fraction_invalid=0.1)
    r = subprocess.run([exe, "polish", "-m5", "--min_depth=4", "-i=0.1", "--", fa] + sams, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == exp["fasta"]
    log = r.stderr.decode()
    depths = re.findall(r"mean read depth: ([0-9.]+)x", log)
    lens = [len(s) for _, _, s in pp.load_fasta(fa).records()]
    assert depths == ["%.1f" % (exp["total_depth"][i] / lens[i]) for i in range(2)]
    changed = [int(x.replace(",", "")) for x in re.findall(r"([0-9,]+) positions? changed", log)]
    assert changed == exp["changed"]
    zero = [int(x.replace(",", "")) for x in re.findall(r"([0-9,]+) bp ha(?:s|ve) a depth of zero", log)]
    assert zero == exp["zero_depth"]
    assert len(re.findall(r"estimated pre-polishing sequence accuracy: [0-9.]+% \(Q[0-9.]+\)", log)) == 2
    # filter: --low=..., --high=...
    o1, o2 = tmp_path / "f1.sam", tmp_path / "f2.sam"
    r = subprocess.run([exe, "filter", "--in1=" + sams[0], "--in2", sams[1], "--out1", str(o1), "--out2=" + str(o2), "--low=1", "--high=99"],
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    ef = oracle.filter(sams[0], sams[1], low=1.0, high=99.0)
    assert open(o1, "rb").read() == ef["out1"] and open(o2, "rb").read() == ef["out2"]


def test_total_depth_statistic(ctx, oracle, tmp_path):
    """pp_polish_result.total_depth: the per-position f64 depths are the reference's; their per-contig sum agrees with the
    oracle's sequential sum to rounding (multi-mapped reads make the depths fractional)."""
    case = fuzzgen.make_case(303, n_contigs=2, contig_len=(200, 400), depth=(150, 300), multimap=0.8, opts=dict(careful=False))
    fa, sams = case.write(tmp_path)
    exp = oracle.polish(fa, sams, **case.opts)
    f = pp.load_fasta(fa)
    p = pp.pack_sams(f, sams)
    r = ctx.polish_packed(f.view, p.view, **case.opts)
    for got, want in zip(r["total_depth"], exp["total_depth"]):
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want))
