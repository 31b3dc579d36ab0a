"""GPU parity tests for `polypolish filter`: thresholds, orientation and the re-streamed SAM files against the
CPU oracle, byte for byte."""
import random

import pytest

import polypolish_b200 as pp
from polypolish_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g
    g.build()
    c = pp.Context(0)
    yield c
    c.close()


def run_both(ctx, oracle, in1, in2, tmp_path, **kw):
    o1, o2 = tmp_path / "o1.sam", tmp_path / "o2.sam"
    try:
        exp = ("ok", oracle.filter(in1, in2, out1=o1, out2=o2, **kw))
    except Exception as e:
        exp = ("err", e.msg)
    try:
        ctx.filter_files(in1, in2, o1, o2, **kw)
        got = ("ok", dict(out1=open(o1, "rb").read(), out2=open(o2, "rb").read()))
    except pp.PolypolishError as e:
        got = ("err", e.msg)
    assert exp[0] == got[0], (exp, got if got[0] == "err" else "")
    if exp[0] == "ok":
        assert got[1]["out1"] == exp[1]["out1"]
        assert got[1]["out2"] == exp[1]["out2"]
    else:
        assert got[1] == exp[1]
    return exp


def random_pairs(seed, n_pairs=300, orient="fr", eol="\n", shuffle_groups=False):
    """Small paired SAMs: unique pairs (insert ~ N(300,30)) plus multi-mapped reads whose mates sit at good and bad
    distances / strands / contigs; unaligned records; reads present in one file only."""
    rng = random.Random(seed)
    refs = ["ctgA", "ctgB"]
    f = [["@HD\tVN:1.6", "@SQ\tSN:ctgA\tLN:100000", "@SQ\tSN:ctgB\tLN:100000"], ["@HD\tVN:1.6"]]

    def rec(name, flag, ref, pos, cigar="100M", extra=()):
        return "\t".join([name, str(flag), ref, str(pos), "60", cigar, "*", "0", "0", "ACGT", "IIII", "NM:i:0", *extra])

    def pair_flags():
        if orient == "fr":
            return 0, 16
        if orient == "rf":
            return 16, 0
        if orient == "ff":
            return 0, 0
        return 16, 16
    for i in range(n_pairs):
        name = f"p{i}"
        ref = rng.choice(refs)
        a = rng.randint(1, 90000)
        ins = max(150, int(rng.gauss(300, 30)))
        fa, fb = pair_flags()
        left, right = a, a + ins - 100
        r = rng.random()
        recs1, recs2 = [rec(name, fa, ref, left, rng.choice(["100M", "50M2D50M", "40M3I57M", "5S95M"]))], [rec(name, fb, ref, right)]
        if r < 0.25:                                           # mate 1 multi-mapped: one good, some bad locations
            for _ in range(rng.randint(1, 4)):
                recs1.append(rec(name, 256 | rng.choice([0, 16]), rng.choice(refs), rng.randint(1, 90000)))
        elif r < 0.4:                                          # both multi-mapped
            for _ in range(rng.randint(1, 3)):
                recs1.append(rec(name, 256 | fa, ref, left + rng.choice([0, 1000, -50, 5])))
                recs2.append(rec(name, 256 | fb, ref, right + rng.choice([0, 1000, 3, 20000])))
        elif r < 0.45:
            recs2 = []                                         # mate missing entirely
        elif r < 0.5:
            recs2 = [rec(name, 4, "*", 0, "*")]                # mate unaligned
        elif r < 0.55:
            recs1.append(rec(name, 4, "*", 0, "*"))
        if rng.random() < 0.5:
            rng.shuffle(recs1)
        f[0].extend(recs1)
        f[1].extend(recs2)
    if shuffle_groups:                                         # names not consecutive: the reference groups by name anyway
        body = f[0][3:]
        rng.shuffle(body)
        f[0] = f[0][:3] + body
    return eol.join(f[0]) + eol, eol.join(f[1]) + eol


@pytest.mark.parametrize("seed,orient", [(1, "fr"), (2, "rf"), (3, "ff"), (4, "rr"), (5, "fr"), (6, "fr")])
def test_random_pairs(ctx, oracle, tmp_path, seed, orient):
    t1, t2 = random_pairs(seed, orient=orient, eol="\r\n" if seed == 5 else "\n", shuffle_groups=(seed == 6))
    i1, i2 = tmp_path / "i1.sam", tmp_path / "i2.sam"
    i1.write_bytes(t1.encode()); i2.write_bytes(t2.encode())
    exp = run_both(ctx, oracle, i1, i2, tmp_path)
    assert exp[0] == "ok" and exp[1]["orientation"] == orient
    assert exp[1]["before_count"] > exp[1]["after_count"] > 0


@pytest.mark.parametrize("kw", [dict(orientation="rf"), dict(orientation="bogus"), dict(low=5.0, high=95.0),
                                dict(low=0.0), dict(high=100.0), dict(low=49.9, high=50.1)])
def test_options_and_errors(ctx, oracle, tmp_path, kw):
    t1, t2 = random_pairs(11)
    i1, i2 = tmp_path / "i1.sam", tmp_path / "i2.sam"
    i1.write_text(t1); i2.write_text(t2)
    run_both(ctx, oracle, i1, i2, tmp_path, **kw)


def test_text_errors(ctx, oracle, tmp_path):
    good1, good2 = random_pairs(12, n_pairs=20)
    cases = [("@HD\n\n", good2), ("@HD\nr1\t4\t*\t0\t0\t*\t*\t0\t0\tAC\tII\n", "@HD\nr1\t4\t*\t0\t0\t*\t*\t0\t0\tAC\tII\n"),
             ("r1\t0\tc\t1\t60\t4M\n", good2), (good1, good2.replace("\n", "\n\n", 1))]
    for k, (a, b) in enumerate(cases):
        d = tmp_path / f"c{k}"
        d.mkdir()
        i1, i2 = d / "i1.sam", d / "i2.sam"
        i1.write_text(a); i2.write_text(b)
        exp = run_both(ctx, oracle, i1, i2, d)
        assert exp[0] == "err"
    # all four paths must differ
    i1, i2 = tmp_path / "i1.sam", tmp_path / "i2.sam"
    i1.write_text(good1); i2.write_text(good2)
    with pytest.raises(pp.PolypolishError) as e:
        ctx.filter_files(i1, i2, i1, tmp_path / "o2.sam")
    assert "must all have unique values" in e.value.msg


def test_auto_orientation_tie_is_an_error(ctx, oracle, tmp_path):
    a1, a2 = random_pairs(21, n_pairs=10, orient="fr")
    b1, b2 = random_pairs(21, n_pairs=10, orient="ff")
    strip = lambda t: "\n".join(l for l in t.split("\n") if l and not l.startswith("@"))
    rename = lambda t: t.replace("p", "q")
    i1, i2 = tmp_path / "i1.sam", tmp_path / "i2.sam"
    i1.write_text(strip(a1) + "\n" + rename(strip(b1)) + "\n")
    i2.write_text(strip(a2) + "\n" + rename(strip(b2)) + "\n")
    exp = run_both(ctx, oracle, i1, i2, tmp_path)
    # (either an exact tie -> error, or a unique winner: parity is what matters)
    run_both(ctx, oracle, i1, i2, tmp_path, orientation="ff")


@pytest.mark.parametrize("seed", [1, 2])
def test_synth_filter_then_polish(ctx, oracle, tmp_path, seed):
    """BASELINE config 3 in miniature: filter (insert size) then polish, both stages against the oracle."""
    syn = api.Synth(seed=seed, contig_len=60_000, depth=60)
    fa, sams = syn.write(tmp_path)
    exp = run_both(ctx, oracle, sams[0], sams[1], tmp_path)
    assert exp[1]["orientation"] == "fr" and 200 <= exp[1]["low"] < exp[1]["high"] <= 700
    f1, f2 = tmp_path / "f1.sam", tmp_path / "f2.sam"
    ctx.filter_files(sams[0], sams[1], f1, f2)
    assert b"ZP:Z:fail" in open(f1, "rb").read()
    got = ctx.polish_files(fa, [f1, f2])
    assert got == oracle.polish(fa, [f1, f2])["fasta"]


# ---- the device text path of `filter` (tok_kernels.cu) against the host text path -------------------------------------
@pytest.mark.parametrize("seed", [31, 32, 33, 34])
def test_device_and_host_text_paths_agree(ctx, oracle, tmp_path, seed):
    """Same bytes from both text paths (pp_set_parser 0 / 1), and from the oracle: CRLF, shuffled groups, last line unterminated."""
    t1, t2 = random_pairs(seed, n_pairs=500, eol="\r\n" if seed == 32 else "\n", shuffle_groups=(seed == 33))
    if seed == 34:
        t1, t2 = t1[:-1], t2[:-1]
    i1, i2 = tmp_path / "i1.sam", tmp_path / "i2.sam"
    i1.write_bytes(t1.encode()); i2.write_bytes(t2.encode())
    exp = oracle.filter(i1, i2)
    for mode in (0, 1):
        o1, o2 = tmp_path / f"o1_{mode}.sam", tmp_path / f"o2_{mode}.sam"
        ctx.set_parser(mode)
        try:
            ctx.filter_files(i1, i2, o1, o2)
        finally:
            ctx.set_parser(0)
        assert open(o1, "rb").read() == exp["out1"] and open(o2, "rb").read() == exp["out2"]


def test_device_text_path_large_files(ctx, oracle, tmp_path):
    """~100 MB per mate: every pinned slot is reused in both directions, 0.6 M names are interned, the second file is parsed
    while ... the first is already resident; the result is the oracle's, byte for byte."""
    syn = api.Synth(seed=21, n_contigs=2, contig_len=400_000, depth=100.0)
    fa, sams = syn.write(tmp_path)
    o1, o2 = tmp_path / "o1.sam", tmp_path / "o2.sam"
    ctx.filter_files(sams[0], sams[1], o1, o2)
    exp = oracle.filter(sams[0], sams[1])
    assert open(o1, "rb").read() == exp["out1"]
    assert open(o2, "rb").read() == exp["out2"]
    assert exp["out1"].count(b"\tZP:Z:fail") > 100


def test_device_text_path_same_name_in_both_columns(ctx, oracle, tmp_path):
    """A QNAME that is also an RNAME, an empty RNAME, names that differ only in the last byte: the interning keeps them apart."""
    rows1 = ["ctgA\t0\tctgA\t100\t60\t50M\t*\t0\t0\tAC\tII", "ctgB\t0\tctgA\t500\t60\t50M\t*\t0\t0\tAC\tII", "x1\t0\t\t100\t60\t50M\t*\t0\t0\tAC\tII",
             "x2\t0\tctgB\t100\t60\t50M\t*\t0\t0\tAC\tII", "x2\t256\tctgA\t9000\t60\t50M\t*\t0\t0\tAC\tII"]
    rows2 = ["ctgA\t16\tctgA\t350\t60\t50M\t*\t0\t0\tAC\tII", "ctgB\t16\tctgA\t760\t60\t50M\t*\t0\t0\tAC\tII", "x1\t16\t\t340\t60\t50M\t*\t0\t0\tAC\tII",
             "x2\t16\tctgB\t330\t60\t50M\t*\t0\t0\tAC\tII", "x2\t272\tctgA\t9300\t60\t50M\t*\t0\t0\tAC\tII", "x2\t272\tctgB\t50000\t60\t50M\t*\t0\t0\tAC\tII"]
    more1, more2 = random_pairs(41, n_pairs=200)
    i1, i2 = tmp_path / "i1.sam", tmp_path / "i2.sam"
    i1.write_text(more1 + "\n".join(rows1) + "\n"); i2.write_text(more2 + "\n".join(rows2) + "\n")
    run_both(ctx, oracle, i1, i2, tmp_path)


def test_filter_to_pipes_and_devices(ctx, oracle, tmp_path):
    """Outputs that are not regular files (the reference streams through a BufWriter, filter.rs:296-349): a named FIFO read
    by another thread, and /dev/null.  The device text path cannot truncate or write those at offsets; it streams in order."""
    import os
    import threading
    syn = api.Synth(seed=12, contig_len=40_000, depth=40)
    fa, sams = syn.write(tmp_path)
    exp = oracle.filter(sams[0], sams[1])
    fifo = tmp_path / "out1.fifo"
    os.mkfifo(fifo)
    got = {}

    def reader():
        with open(fifo, "rb") as f:
            got["out1"] = f.read()
    t = threading.Thread(target=reader)
    t.start()
    ctx.set_parser(0)
    ctx.filter_files(sams[0], sams[1], fifo, "/dev/null")
    t.join(timeout=60)
    assert not t.is_alive()
    assert got["out1"] == exp["out1"]
    # and the other way round, second output through the pipe
    t = threading.Thread(target=reader)
    t.start()
    ctx.filter_files(sams[1], sams[0], fifo, tmp_path / "plain.sam")
    t.join(timeout=60)
    assert got["out1"] == exp["out2"]
    assert open(tmp_path / "plain.sam", "rb").read() == exp["out1"]


@pytest.mark.parametrize("seed,opts", [(14, {}), (15, dict(careful=True)), (16, dict(max_errors=3, min_depth=3))])
def test_filter_then_polish_in_one_call(ctx, oracle, tmp_path, seed, opts):
    """pp_filter_polish_files: `filter` + `polish` without the intermediate files == the two reference commands
    one after the other (filter.rs:334-342 writes ZP:Z:fail, alignment.rs:72-74 reads it back)."""
    syn = api.Synth(seed=seed, n_contigs=2, contig_len=30_000, depth=40)
    fa, sams = syn.write(tmp_path)
    fo = oracle.filter(sams[0], sams[1])
    o1, o2 = tmp_path / "o1.sam", tmp_path / "o2.sam"
    o1.write_bytes(fo["out1"])
    o2.write_bytes(fo["out2"])
    assert fo["out1"].count(b"ZP:Z:fail") + fo["out2"].count(b"ZP:Z:fail") > 20
    exp = oracle.polish(fa, [o1, o2], **opts)["fasta"]
    ctx.set_parser(0)
    assert ctx.filter_polish_files(fa, sams[0], sams[1], **opts) == exp                     # nothing written
    f1, f2 = tmp_path / "f1.sam", tmp_path / "f2.sam"
    assert ctx.filter_polish_files(fa, sams[0], sams[1], f1, f2, **opts) == exp             # filtered files written on the way
    assert f1.read_bytes() == fo["out1"] and f2.read_bytes() == fo["out2"]
    ctx.set_parser(1)                                                                       # host text path: the two commands through files
    try:
        assert ctx.filter_polish_files(fa, sams[0], sams[1], **opts) == exp
    finally:
        ctx.set_parser(0)


def test_filter_then_polish_errors_and_odd_text(ctx, oracle, tmp_path):
    """The fused call hands anything unusual to the text code: same errors as the reference's two commands."""
    syn = api.Synth(seed=18, contig_len=20_000, depth=30)
    fa, sams = syn.write(tmp_path)
    # a read group without SEQ in file 1 -> polish's error, worded by the host path
    bad = tmp_path / "bad_1.sam"
    bad.write_bytes(open(sams[0], "rb").read() + b"zz\t0\tcontig_1\t100\t60\t50M\t*\t0\t0\t*\t*\tNM:i:0\n")
    with pytest.raises(pp.PolypolishError) as e:
        ctx.filter_polish_files(fa, bad, sams[1])
    assert "no alignments for read zz contain sequence" in e.value.msg
    with pytest.raises(pp.PolypolishError) as e:
        ctx.filter_polish_files(fa, sams[0], sams[1], low=60.0)
    assert "--low must be greater than 0 and less than 50" in e.value.msg


def test_cli_filter_polish(oracle, tmp_path):
    """The additive `polypolish filter-polish` command = `filter` then `polish` of its output."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "polypolish")
    syn = api.Synth(seed=19, contig_len=25_000, depth=40)
    fa, sams = syn.write(tmp_path)
    fo = oracle.filter(sams[0], sams[1])
    o1, o2 = tmp_path / "o1.sam", tmp_path / "o2.sam"
    o1.write_bytes(fo["out1"])
    o2.write_bytes(fo["out2"])
    exp = oracle.polish(fa, [o1, o2], min_depth=4)["fasta"]
    r = subprocess.run([exe, "filter-polish", "--in1", sams[0], "--in2=" + sams[1], "-d4", "--quiet", fa], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == exp
