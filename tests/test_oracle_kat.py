"""Known-answer tests: every vector carried by the reference's own 21 unit tests that touches the
polishing path, checked against the CPU oracle.  These pin the oracle; the GPU
path is then compared with the oracle (tests/test_gpu_*.py)."""
import gzip

import pytest


# alignment.rs:386-392 test_get_expanded_cigar_good
@pytest.mark.parametrize("cigar,n,exp", [
    ("10M", 10, "MMMMMMMMMM"), ("3M1I7M", 11, "MMMIMMMMMMM"), ("5M2D4M", 9, "MMMMMDDMMMM"),
    ("5=2X3=", 10, "=====XX==="), ("*", 1, "")])
def test_expanded_cigar_good(oracle, cigar, n, exp):
    assert oracle.expanded_cigar(cigar, n) == exp


# alignment.rs:395-399 test_get_expanded_cigar_bad
@pytest.mark.parametrize("cigar", ["10Q", "10MM1I10M", "100M5"])
def test_expanded_cigar_bad(oracle, cigar):
    assert oracle.expanded_cigar(cigar, 10) is None


# alignment.rs:402-422 test_get_ref_positions
@pytest.mark.parametrize("cigar,end", [("4M", 1003), ("2=1X1=", 1003), ("2M1I1M", 1002), ("2M1D1M", 1003)])
def test_ref_positions(oracle, cigar, end):
    line = f"r_1\t0\tx\t1000\t60\t{cigar}\t*\t0\t0\tACTG\tKKKK\tNM:i:0"
    rs, re_, nm, qc = oracle.alignment_new(line)
    assert (rs, re_, nm, qc) == (999, end, 0, True)


# filter.rs:396-424 test_get_orientation
@pytest.mark.parametrize("p1,p2,f1,f2,exp", [
    (100000, 200000, 0, 16, "fr"), (200000, 100000, 16, 0, "fr"),
    (200000, 100000, 0, 16, "rf"), (100000, 200000, 16, 0, "rf"),
    (100000, 200000, 0, 0, "ff"), (200000, 100000, 16, 16, "ff"),
    (200000, 100000, 0, 0, "rr"), (100000, 200000, 16, 16, "rr")])
def test_get_orientation(oracle, p1, p2, f1, f2, exp):
    l1 = f"r_1\t{f1}\tx\t{p1}\t60\t150M\t*\t0\t0\tACTG\tKKKK\tNM:i:0"
    l2 = f"r_2\t{f2}\tx\t{p2}\t60\t150M\t*\t0\t0\tACTG\tKKKK\tNM:i:0"
    o, ins = oracle.orientation(l1, l2)
    assert o == exp
    assert ins == 100150


# filter.rs:427-447 test_auto_determine_orientation
@pytest.mark.parametrize("counts,exp", [((3, 1, 1, 1), "fr"), ((1, 3, 1, 1), "rf"), ((1, 1, 3, 1), "ff"),
                                        ((1, 1, 1, 3), "rr"), ((3, 3, 1, 1), None)])
def test_auto_orientation(oracle, counts, exp):
    assert oracle.auto_orientation(counts) == exp


# filter.rs:450-462 test_get_percentile
@pytest.mark.parametrize("p,exp", [(0.1, 15), (19.9, 15), (20.1, 20), (39.9, 20), (40.1, 35), (59.9, 35),
                                   (60.1, 40), (79.9, 40), (80.1, 50), (99.9, 50)])
def test_percentile(oracle, p, exp):
    assert oracle.percentile([15, 20, 35, 40, 50], p) == exp
    assert oracle.percentile([], p) == 0


# pileup.rs:209-295 test_pileupbase_01..08
VOTE = [
    ("A", [("A", 1.0)] * 50, 0.2, "Ax50", "A", "kept"),
    ("G", [("A", 1.0), ("T", 1.0)] + [("G", 1.0)] * 50, 0.2, "Ax1,Gx50,Tx1", "G", "kept"),
    ("T", [("C", 1.0)] + [("A", 1.0)] * 99, 0.2, "Ax99,Cx1", "A", "changed"),
    ("A", [("T", 1.0), ("C", 1.0), ("G", 1.0)], 0.2, "Cx1,Gx1,Tx1", "A", "low_depth"),
    ("C", [("A", 0.1)] * 123 + [("T", 0.1)] * 321, 0.2, "Ax123,Tx321", "C", "multiple"),
    ("T", [("A", 1.0)] * 6 + [("C", 1.0)] * 4, 0.2, "Ax6,Cx4", "T", "too_close"),
    ("T", [("A", 1.0)] * 9 + [("C", 1.0)], 0.1, "Ax9,Cx1", "T", "too_close"),
    ("T", [("A", 1.0)] * 19 + [("C", 1.0)], 0.1, "Ax19,Cx1", "A", "changed"),
]


@pytest.mark.parametrize("orig,adds,fi,cstr,new,status", VOTE)
def test_pileupbase(oracle, orig, adds, fi, cstr, new, status):
    nb, st, cs, _ = oracle.vote(orig, adds, min_depth=5, fv=0.5, fi=fi)
    assert (cs, nb, st) == (cstr, new, status)


# misc.rs:280-296 test_bankers_rounding
@pytest.mark.parametrize("x,exp", [(0.0, 0), (123.0, 123), (98765.0, 98765), (0.4999, 0), (0.5, 0), (0.5001, 1),
                                   (42.45, 42), (42.5, 42), (42.55, 43), (12345.4998, 12345), (12345.5, 12346),
                                   (12345.5002, 12346)])
def test_bankers_rounding(oracle, x, exp):
    assert oracle.bankers_rounding(x) == exp


# misc.rs:299-304 test_reverse_complement
def test_reverse_complement(oracle):
    assert oracle.reverse_complement("GGTATCACTCAGGAAGC") == "GCTTCCTGAGTGATACC"
    assert oracle.reverse_complement("GGGGaaaaaaaatttatatat") == "atatataaattttttttCCCC"
    assert oracle.reverse_complement("atatataaattttttttCCCC") == "GGGGaaaaaaaatttatatat"
    assert oracle.reverse_complement("ACGT123") == "NNNACGT"


# misc.rs:246-267 test_load_fasta_1 / _2
FASTA_TXT = ">seq_1 123 456\nACGAT\n>seq_2 abc\nGGTA\n>seq_3\nCTCGCATCAG\n"
FASTA_EXP = [("seq_1", "123 456", "ACGAT"), ("seq_2", "abc", "GGTA"), ("seq_3", "", "CTCGCATCAG")]


def test_load_fasta_plain(oracle, tmp_path):
    p = tmp_path / "a.fasta"
    p.write_text(FASTA_TXT)
    assert oracle.load_fasta(p) == FASTA_EXP


def test_load_fasta_gz(oracle, tmp_path):
    p = tmp_path / "a.fasta.gz"
    with gzip.open(p, "wt") as f:
        f.write(FASTA_TXT)
    assert oracle.load_fasta(p) == FASTA_EXP


# Hand-derived from alignment.rs:175-201 and :364-378; not reference tests.
def test_walk_and_trim_hand_derived(oracle):
    assert oracle.walk("10M", "ACGTACGTAA") == [(i, i + 1) for i in range(7)]
    assert oracle.walk("2M1I2M", "ACGTT") == [(0, 1)]
    assert oracle.walk("2M1D2M", "ACGT") == [(0, 1), (1, 2), (2, 2)]
    # D followed by I turns the deletion entry into a one-base entry (Appendix A.4)
    assert oracle.walk("2M1D1I3M", "ACGTCA") == [(0, 1), (1, 2), (2, 3), (3, 4)]
    # whole read is one homopolymer: everything is trimmed
    assert oracle.walk("5M", "AAAAA") == []


def test_end_to_end_tiny(oracle, tmp_path):
    """Smallest whole-command case, worked by hand from polish.rs:157-203: 6 identical reads fix one
    substitution; trimmed read tails leave the last bases at low depth."""
    fa = tmp_path / "asm.fasta"
    fa.write_text(">c1 desc here\nACGTACGTACGTACGTACGT\n")
    read = "ACGTACGAACGTACGTACGT"   # T->A at 0-based position 7
    sam = tmp_path / "r.sam"
    lines = ["@SQ\tSN:c1\tLN:20"]
    for i in range(6):
        lines.append(f"r{i}\t0\tc1\t1\t60\t20M\t*\t0\t0\t{read}\t{'I' * 20}\tNM:i:1")
    sam.write_text("\n".join(lines) + "\n")
    out = oracle.polish(fa, [sam])
    assert out["fasta"] == b">c1 desc here polypolish\nACGTACGAACGTACGTACGT\n"
    assert out["changed"] == [1]
    assert out["alignment_total"] == 6 and out["used_total"] == 6
    assert out["zero_depth"] == [2]       # trim pops the final T and one more
