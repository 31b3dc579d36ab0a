"""GPU tests of the device SAM tokeniser (tok_kernels.cu): the arrays it builds in HBM must equal the host packer's
(sam_pack.cpp) bit for bit, and `polish_files` must print the same bytes / raise the same errors with either parser."""
import numpy as np
import pytest

import polypolish_b200 as pp
from polypolish_b200 import api
from tests import fuzzgen
from tests.test_tok_cpu import BAD_LINES, EDGE_TEXTS, FA, line

pytestmark = pytest.mark.gpu
PP_OK, PP_TOK_HOST, PP_TOK_NEED8 = 0, 1, 2


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g
    g.build()
    c = pp.Context(0)
    yield c
    c.close()


def host_arrays(fa_path, sources, careful):
    f = pp.load_fasta(fa_path)
    p = api.Packed(f, careful)
    p.set_threads(1)
    for s in sources:
        if isinstance(s, (bytes, bytearray)):
            p.add_text(s)
        else:
            p.add_file(s)
    p.finish()
    return f, p


def device_arrays(ctx, f, sources, careful, bits):
    rc, stats = ctx.tokenise(f, sources, careful, bits)
    if rc != PP_OK:
        return rc, None, stats
    return rc, ctx.dataset_arrays(), stats


def assert_same(dev, host):
    for k, v in host.items():
        if isinstance(v, np.ndarray):
            assert v.shape == dev[k].shape, k
            assert np.array_equal(v, dev[k]), k
        else:
            assert v == dev[k], k


@pytest.mark.parametrize("seed", range(300, 324))
def test_tokeniser_equals_host_packer_on_fuzz(ctx, tmp_path, seed):
    case = fuzzgen.make_case(seed, exotic=0.5 if seed % 6 == 0 else 0.0, multimap=0.5 if seed % 2 else 0.25)
    fa, sams = case.write(tmp_path)
    careful = case.opts["careful"]
    f, p = host_arrays(fa, sams, careful)
    host = p.arrays()
    rc, dev, stats = device_arrays(ctx, f, sams, careful, 4)
    if host["seq_bits"] == 8:
        assert rc == PP_TOK_NEED8
        rc, dev, stats = device_arrays(ctx, f, sams, careful, 8)
    assert rc == PP_OK
    assert_same(dev, host)
    for i, st in enumerate(stats):
        na, nr = api.C.c_uint64(), api.C.c_uint64()
        api.lib().pp_pack_file_stats(p.h, i, api.C.byref(na), api.C.byref(nr))
        assert (st["alignments"], st["reads"]) == (na.value, nr.value)
    # the tokenised dataset polishes to the same bytes as the uploaded host arrays
    opts = {k: v for k, v in case.opts.items()}
    got = ctx.polish_resident(**opts)
    exp = ctx.polish_packed(f.view, p.view, **opts)
    assert got["sequences"] == exp["sequences"] and got["n_aln_used"] == exp["n_aln_used"]


@pytest.mark.parametrize("i", range(len(EDGE_TEXTS)))
@pytest.mark.parametrize("careful", [False, True])
def test_tokeniser_edge_texts(ctx, tmp_path, i, careful):
    fa = tmp_path / "a.fasta"
    fa.write_text(FA)
    t = EDGE_TEXTS[i].encode("latin-1")
    try:
        f, p = host_arrays(fa, [t], careful)
    except pp.PolypolishError:
        f = pp.load_fasta(fa)
        assert device_arrays(ctx, f, [t], careful, 4)[0] == PP_TOK_HOST
        return
    host = p.arrays()
    rc, dev, _ = device_arrays(ctx, f, [t], careful, int(host["seq_bits"]))
    assert rc == PP_OK
    assert_same(dev, host)


@pytest.mark.parametrize("bad", BAD_LINES)
def test_tokeniser_hands_bad_text_to_the_host(ctx, oracle, tmp_path, bad):
    """A malformed line anywhere: the tokeniser answers PP_TOK_HOST, and polish_files ends with the reference's message
    (the oracle's) whichever parser is selected."""
    fa = tmp_path / "a.fasta"
    fa.write_text(FA)
    good = "".join(line(f"g{i}") for i in range(300))
    s = tmp_path / "s.sam"
    s.write_text(good + bad + good)
    f = pp.load_fasta(fa)
    assert device_arrays(ctx, f, [s], False, 4)[0] == PP_TOK_HOST
    try:
        oracle.polish(fa, [s])
        ref_msg = None                      # a limit of this implementation, not a reference error
    except Exception as e:
        ref_msg = e.msg
    msgs = []
    for mode in (0, 1):
        ctx.set_parser(mode)
        try:
            with pytest.raises(pp.PolypolishError) as ei:
                ctx.polish_files(fa, [s])
        finally:
            ctx.set_parser(0)
        msgs.append(ei.value.msg)
    assert msgs[0] == msgs[1]
    if ref_msg is not None and not ref_msg.startswith("panic") and "not supported" not in msgs[0]:
        assert msgs[0] == ref_msg


def test_tokeniser_empty_and_missing_files(ctx, oracle, tmp_path):
    fa = tmp_path / "a.fasta"
    fa.write_text(FA)
    f = pp.load_fasta(fa)
    e = tmp_path / "empty.sam"
    e.write_text("")
    h = tmp_path / "hdr.sam"
    h.write_text("@HD\tVN:1.6\n@SQ\tSN:c1\tLN:40\n")
    for s in (e, h):
        assert device_arrays(ctx, f, [s], False, 4)[0] == PP_TOK_HOST
        with pytest.raises(Exception) as eo:
            oracle.polish(fa, [s])
        with pytest.raises(pp.PolypolishError) as ei:
            ctx.polish_files(fa, [s])
        assert ei.value.msg == eo.value.msg
    assert device_arrays(ctx, f, [tmp_path / "nope.sam"], False, 4)[0] == PP_TOK_HOST


def test_tokeniser_synth_two_files(ctx, oracle, tmp_path):
    """50 kbp x 100x: two SAM files (one per mate), ~12 MB of text each; arrays, per-file statistics and polished bytes."""
    syn = api.Synth(seed=11, n_contigs=3, contig_len=50_000, depth=60.0)
    fa, sams = syn.write(tmp_path)
    f, p = host_arrays(fa, sams, False)
    rc, dev, stats = device_arrays(ctx, f, sams, False, 4)
    assert rc == PP_OK
    assert_same(dev, p.arrays())
    assert stats[0]["lines"] > 10_000 and stats[0]["launches"] >= 10
    exp = oracle.polish(fa, sams)["fasta"]
    for mode in (0, 1):
        ctx.set_parser(mode)
        try:
            assert ctx.polish_files(fa, sams) == exp
        finally:
            ctx.set_parser(0)


def test_tokeniser_large_file_streaming(ctx, tmp_path):
    """A file several times the pinned ring (4 readers x 2 slots x 8 MiB): every slot is reused, the text crosses thousands of
    16 KiB tiles, and the second file makes every dataset array grow while keeping the first file's records."""
    syn = api.Synth(seed=12, n_contigs=1, contig_len=800_000, depth=100.0)
    fa, sams = syn.write(tmp_path)
    import os
    assert os.path.getsize(sams[0]) > 80 << 20
    f, p = host_arrays(fa, sams, False)
    rc, dev, stats = device_arrays(ctx, f, sams, False, 4)
    assert rc == PP_OK
    assert_same(dev, p.arrays())
    # unterminated last line + CRLF endings on a big text
    raw = open(sams[0], "rb").read()
    crlf = raw.replace(b"\n", b"\r\n")[:-2]
    c = tmp_path / "crlf.sam"
    c.write_bytes(crlf)
    f2, p2 = host_arrays(fa, [c], False)
    rc, dev2, _ = device_arrays(ctx, f2, [c], False, 4)
    assert rc == PP_OK
    assert_same(dev2, p2.arrays())


def test_tokeniser_unknown_reference_goes_to_the_polish_error(ctx, oracle, tmp_path):
    fa = tmp_path / "a.fasta"
    fa.write_text(FA)
    s = tmp_path / "s.sam"
    s.write_text("".join(line(f"g{i}") for i in range(50)) + line("bad", ref="c3") + line("z"))
    f = pp.load_fasta(fa)
    rc, dev, _ = device_arrays(ctx, f, [s], False, 4)
    assert rc == PP_OK and dev["contig"][50] == 0xFFFFFFFF
    with pytest.raises(Exception) as eo:
        oracle.polish(fa, [s])
    with pytest.raises(pp.PolypolishError) as ei:
        ctx.polish_files(fa, [s])
    assert ei.value.msg == eo.value.msg and "c3" in ei.value.msg


def test_tokeniser_many_contigs_three_files(ctx, oracle, tmp_path):
    """3000 contig names (hash-table probing, names that are prefixes of each other), three SAM files, records of every file
    hitting every contig; plus the polished bytes through both parsers."""
    import random
    rng = random.Random(77)
    names = [f"c{i}" for i in range(1500)] + [f"c{i}_x" for i in range(1500)]
    seqs = ["".join(rng.choice("ACGT") for _ in range(60)) for _ in names]
    fa = tmp_path / "many.fasta"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in zip(names, seqs)))
    files = []
    for k in range(3):
        rows = []
        for j in range(4000):
            c = rng.randrange(len(names))
            st = rng.randrange(0, 20)
            rows.append("\t".join([f"r{k}_{j // 2}", str(rng.choice([0, 16])), names[c], str(st + 1), "60", "40M", "*", "0", "0", seqs[c][st:st + 40], "I" * 40,
                                   "NM:i:0"]))
        p = tmp_path / f"s{k}.sam"
        p.write_text("\n".join(rows) + "\n")
        files.append(p)
    f, p = host_arrays(fa, files, False)
    rc, dev, stats = device_arrays(ctx, f, files, False, 4)
    assert rc == PP_OK and len(stats) == 3
    assert_same(dev, p.arrays())
    exp = oracle.polish(fa, files)["fasta"]
    for mode in (0, 1):
        ctx.set_parser(mode)
        try:
            assert ctx.polish_files(fa, files) == exp
        finally:
            ctx.set_parser(0)


def test_tokeniser_slice_boundaries(ctx, tmp_path):
    """Lines straddling the 4 MiB upload slices and the 16 KiB newline tiles, CRLF pairs split by a slice boundary, a last line
    without newline: a text built so that '\\r' and '\\n' of one line end sit on either side of byte 4 Mi."""
    fa = tmp_path / "a.fasta"
    fa.write_text(FA)
    body = []
    total = 0
    i = 0
    target = 4 << 20
    while total + 6000 < target:
        ln = line(f"q{i}", seq="ACGT" * 10, cigar="40M", qual="I" * 40).replace("\n", "\r\n")
        body.append(ln)
        total += len(ln)
        i += 1
    want = target - total + 1                    # the filler line's '\r' lands on byte 4 Mi - 1, its '\n' on byte 4 Mi
    filler = line("z", seq="ACGT" * 10, cigar="40M", qual="I" * 40, tags=("NM:i:0", "XX:Z:" + "y" * 4000))
    base_len = len(filler.replace("\n", "\r\n"))
    extra = want - base_len
    filler = line("z", seq="ACGT" * 10, cigar="40M", qual="I" * 40, tags=("NM:i:0", "XX:Z:" + "y" * (4000 + extra))).replace("\n", "\r\n")
    text = "".join(body) + filler + "".join(body[:500]) + line("last")[:-1]
    raw = text.encode()
    assert raw[target - 1:target + 1] == b"\r\n"
    s = tmp_path / "b.sam"
    s.write_bytes(raw)
    f, p = host_arrays(fa, [s], False)
    rc, dev, _ = device_arrays(ctx, f, [s], False, 4)
    assert rc == PP_OK
    assert_same(dev, p.arrays())
