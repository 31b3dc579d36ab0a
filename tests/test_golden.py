"""Committed vectors (tests/golden/, oracle-generated: see make_golden.py for what they do and do not pin)."""
import json
import pathlib

import pytest

import polypolish_b200 as pp

G = pathlib.Path(__file__).resolve().parent / "golden"
POLISH = sorted(p.name for p in G.glob("polish_*"))


@pytest.mark.parametrize("name", POLISH)
def test_oracle_reproduces_golden(oracle, name):
    d = G / name
    opts = json.loads((d / "opts.json").read_text())
    sams = sorted(d.glob("reads_*.sam"))
    r = oracle.polish(d / "asm.fasta", sams, debug=True, **opts)
    assert r["fasta"] == (d / "expected.fasta").read_bytes()
    assert r["debug_tsv"] == (d / "expected_debug.tsv").read_bytes()
    st = json.loads((d / "expected_stats.json").read_text())
    assert (r["changed"], r["zero_depth"], r["used_total"]) == (st["changed"], st["zero_depth"], st["used_total"])


def test_oracle_reproduces_golden_filter(oracle):
    d = G / "filter_6"
    r = oracle.filter(d / "in_1.sam", d / "in_2.sam")
    assert r["out1"] == (d / "expected_1.sam").read_bytes() and r["out2"] == (d / "expected_2.sam").read_bytes()
    e = json.loads((d / "expected.json").read_text())
    assert (r["low"], r["high"], r["orientation"]) == (e["low"], e["high"], e["orientation"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", POLISH)
def test_gpu_matches_golden(name):
    d = G / name
    opts = json.loads((d / "opts.json").read_text())
    sams = sorted(d.glob("reads_*.sam"))
    assert pp.polish(d / "asm.fasta", sams, **opts) == (d / "expected.fasta").read_bytes()


@pytest.mark.gpu
def test_gpu_matches_golden_filter(tmp_path):
    d = G / "filter_6"
    o1, o2 = tmp_path / "o1.sam", tmp_path / "o2.sam"
    pp.filter_sams(d / "in_1.sam", d / "in_2.sam", o1, o2)
    assert o1.read_bytes() == (d / "expected_1.sam").read_bytes() and o2.read_bytes() == (d / "expected_2.sam").read_bytes()
