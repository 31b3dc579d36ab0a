"""The 2-bit wire format of a packed batch (pp_alignments_to_2bit, pp_abi.h seq_bits == 2): the host conversion decoded back on the
CPU, and (GPU) the polish of a 2-bit upload against the oracle and against the 4-bit upload of the same batch."""
import numpy as np
import pytest

import polypolish_b200 as pp
from polypolish_b200 import api
from tests import fuzzgen

NIB = "=ACMGRSVTWYHKDBN"
PP_FLAG_SEQSTAR, PP_FLAG_NOSEQ, PP_FLAG_ESC, PP_FLAG_NEWGROUP = 0x04, 0x10, 0x40, 0x80


def _flag_values():
    import re, os
    text = open(os.path.join(api.ROOT, "include", "pp_abi.h")).read()
    return {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define PP_FLAG_(\w+)\s+0x([0-9a-fA-F]+)", text)}


def test_flag_constants_match_the_header():
    f = _flag_values()
    assert (f["SEQSTAR"], f["NOSEQ"], f["ESC"], f["NEWGROUP"]) == (PP_FLAG_SEQSTAR, PP_FLAG_NOSEQ, PP_FLAG_ESC, PP_FLAG_NEWGROUP)


def _seq4(pool, off, n):
    b = pool[off * 16: off * 16 + (n + 1) // 2]
    return "".join(NIB[(int(b[i >> 1]) >> ((i & 1) * 4)) & 15] for i in range(n))


def _seq2(pool, off, n):
    b = pool[off * 8: off * 8 + (n + 3) // 4]
    return "".join("ACGT"[(int(b[i >> 2]) >> ((i & 3) * 2)) & 3] for i in range(n))


def _check_roundtrip(p):
    src = api.view_arrays(p.view)
    tb = api.TwoBit(p.view)
    two = api.view_arrays(tb.view)
    assert two["seq_bits"] == 2 and len(two["seq_pool"]) * 2 == len(src["seq_pool"])
    n_esc = 0
    for i in range(p.view.n_aln):
        fl = int(src["flags"][i])
        assert int(two["flags"][i]) & ~(PP_FLAG_ESC | PP_FLAG_NEWGROUP) == fl
        if fl & PP_FLAG_NOSEQ:
            assert not int(two["flags"][i]) & PP_FLAG_ESC
            continue
        n = int(src["seq_len"][i])
        want = _seq4(src["seq_pool"], int(src["seq_off"][i]), n)
        if int(two["flags"][i]) & PP_FLAG_ESC:
            n_esc += 1
            assert set(want) - set("ACGT")
            assert _seq4(two["esc_pool"], int(two["seq_off"][i]), n) == want
        else:
            assert not set(want) - set("ACGT")
            assert two["seq_off"][i] == src["seq_off"][i]
            assert _seq2(two["seq_pool"], int(two["seq_off"][i]), n) == want
    # the two arrays the device rebuilds: left out exactly when they are what the packer makes them, and then recoverable
    n = p.view.n_aln
    if not tb.view.cigar_off:
        assert list(np.concatenate(([0], np.cumsum(src["n_cigar"].astype(np.int64))[:-1]))) == list(src["cigar_off"]) if n else True
    else:
        assert list(two["cigar_off"]) == list(src["cigar_off"])
    if not tb.view.read_id:
        starts = (two["flags"].astype(np.int64) >> 7) & 1
        assert list(np.cumsum(starts) - 1) == list(src["read_id"])
    else:
        assert list(two["read_id"]) == list(src["read_id"]) and not (two["flags"] & PP_FLAG_NEWGROUP).any()
    assert (not tb.view.cigar_off) and (not tb.view.read_id)        # (the host packer's batches always qualify)
    # untouched arrays are shared, not copied
    assert tb.view.contig == p.view.contig and tb.view.cigar_ops == p.view.cigar_ops
    tb.close()
    return n_esc


@pytest.mark.parametrize("seed", range(900, 912))
def test_host_conversion_round_trip(tmp_path, seed):
    case = fuzzgen.make_case(seed)
    fa, sams = case.write(tmp_path)
    f = api.Fasta(fa)
    try:
        p = api.pack_sams(f, sams, **{k: v for k, v in case.opts.items() if k == "careful"})
    except pp.PolypolishError:
        pytest.skip("case the packer rejects")
    if p.view.seq_bits != 4:
        with pytest.raises(ValueError):
            api.TwoBit(p.view)
        return
    _check_roundtrip(p)


def test_escapes_and_shared_sequences(tmp_path):
    """A read with an N (escaped), its SEQ="*" secondary (shares the escaped copy), a plain read whose secondary shares the 2-bit copy,
    an unmapped line, and a 33-base read (two blocks)."""
    fa = tmp_path / "a.fasta"
    fa.write_text(">c1\n" + "ACGTTGCAAGCTTAGGCATCGATTACGGATCCATGCAAGTCCGATAGGCT" * 2 + "\n")
    rows = [
        "r1\t0\tc1\t1\t60\t20M\t*\t0\t0\tACGTTGCAAGNTTAGGCATC\t*\tNM:i:1",
        "r1\t256\tc1\t51\t0\t20M\t*\t0\t0\t*\t*\tNM:i:1",
        "r2\t0\tc1\t3\t60\t33M\t*\t0\t0\tGTTGCAAGCTTAGGCATCGATTACGGATCCATG\t*\tNM:i:0",
        "r2\t272\tc1\t53\t0\t33M\t*\t0\t0\t*\t*\tNM:i:0",
        "r3\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t*",
        "r4\t16\tc1\t11\t60\t10M\t*\t0\t0\tCTTAGGCATC\t*\tNM:i:0",
    ]
    sam = tmp_path / "a.sam"
    sam.write_text("\n".join(rows) + "\n")
    p = api.pack_sams(api.Fasta(str(fa)), [str(sam)])
    assert _check_roundtrip(p) == 2


@pytest.fixture(scope="module")
def ctx():
    import __graft_entry__ as g
    g.build()
    c = pp.Context(0)
    yield c
    c.close()


def _fasta_bytes(f, r):
    return b"".join(b">" + f.names[i].encode() + (b" " + f.descriptions[i].encode() if f.descriptions[i] else b"") + b" polypolish\n" +
                    r["sequences"][i] + b"\n" for i in range(len(f.names)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(900, 912)) + list(range(300, 304)))
def test_two_bit_upload_parity(ctx, oracle, tmp_path, seed):
    case = fuzzgen.make_case(seed) if seed >= 900 else fuzzgen.make_case(seed, n_contigs=2, contig_len=(200, 400), depth=(150, 300), multimap=0.8,
                                                                           opts=dict(careful=False))
    fa, sams = case.write(tmp_path)
    try:
        exp = oracle.polish(fa, sams, **case.opts)
    except Exception:
        pytest.skip("error case (covered by test_gpu_polish)")
    f = api.Fasta(fa)
    p = api.pack_sams(f, sams, careful=bool(case.opts.get("careful")))
    if p.view.seq_bits != 4:
        pytest.skip("8-bit batch")
    tb = api.TwoBit(p.view)
    r4 = ctx.polish_packed(f.view, p.view, **case.opts)
    r2 = ctx.polish_packed(f.view, tb.view, **case.opts)
    assert r2["sequences"] == r4["sequences"] and r2["changed"] == r4["changed"] and r2["n_aln_used"] == r4["n_aln_used"]
    assert r2["n_aln_used"] == exp["used_total"] and r2["changed"] == exp["changed"] and r2["zero_depth"] == exp["zero_depth"]
    assert [s.decode() for s in r2["sequences"]] == [s.decode() for s in exp["sequences"]] if "sequences" in exp else True
    tb.close()


@pytest.mark.gpu
def test_two_bit_upload_synth(ctx, oracle, tmp_path):
    syn = api.Synth(seed=11, n_contigs=2, contig_len=300_000, depth=60)
    fa, sams = syn.write(tmp_path)
    exp = oracle.polish(fa, sams)
    f = syn.fasta()
    p = syn.pack(f)
    tb = api.TwoBit(p.view)
    a = api.view_arrays(tb.view)
    assert tb.view.seq_pool_bytes * 2 == p.view.seq_pool_bytes
    r = ctx.polish_packed(f.view, tb.view)
    fasta = b"".join(b">" + f.names[i].encode() + b" " + f.descriptions[i].encode() + b" polypolish\n" + r["sequences"][i] + b"\n"
                     for i in range(len(f.names)))
    assert fasta == exp["fasta"]
    # resident route through the same upload
    ctx.upload(f.view, tb.view)
    assert ctx.polish_resident()["sequences"] == r["sequences"]
    tb.close()


@pytest.mark.gpu
def test_two_bit_argument_errors(ctx, tmp_path):
    syn = api.Synth(seed=1, contig_len=20_000, depth=20)
    f = syn.fasta()
    p = syn.pack(f)
    tb = api.TwoBit(p.view)
    bad = api.Alignments()
    import ctypes as C
    C.memmove(C.byref(bad), C.byref(tb.view), C.sizeof(api.Alignments))
    bad.seq_pool_bytes = tb.view.seq_pool_bytes - 3
    with pytest.raises(pp.PolypolishError):
        ctx.polish_packed(f.view, bad)
    # sharding works on 4-/8-bit batches only
    with pytest.raises(Exception):
        api.Shards(f.view, tb.view, 2)
    tb.close()


def test_sharded_batches_convert_too():
    """What bench.py --gpus N puts on the wire: a contig shard (ghost records included) in the 2-bit format, both rebuilt arrays left out."""
    syn = api.Synth(seed=5, n_contigs=4, contig_len=20_000, depth=30, cross_contig=0.05)
    f = syn.fasta()
    p = syn.pack(f)
    sh = api.Shards(f.view, p.view, 2)
    for s in range(2):
        c, a, cmap, n_home = sh.get(s)
        tb = api.TwoBit(a)
        two, src = api.view_arrays(tb.view), api.view_arrays(a)
        assert not tb.view.cigar_off and not tb.view.read_id
        assert list(np.cumsum((two["flags"].astype(np.int64) >> 7) & 1) - 1) == list(src["read_id"])
        assert list(np.concatenate(([0], np.cumsum(src["n_cigar"].astype(np.int64))[:-1]))) == list(src["cigar_off"])
        tb.close()
