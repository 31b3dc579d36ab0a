"""Offline campaign: the two restatements of the reference (oracle/polypolish_oracle.cpp and oracle/pyport.py) against each other on
many fuzz cases - FASTA, debug TSV, statistics, error text; `filter` on synthetic pairs with every orientation.  CPU only.
  python tests/manual/pyport_campaign.py [first_seed] [n_cases]"""
import importlib.util
import pathlib
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import fuzzgen, oracle_lib          # noqa: E402
from polypolish_b200 import api                # noqa: E402

spec = importlib.util.spec_from_file_location("pyport", ROOT / "oracle" / "pyport.py")
pyport = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pyport)


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    oracle_lib.build()
    orc = oracle_lib.load()
    ok = err = 0
    for seed in range(first, first + n):
        kw = {}
        if seed % 5 == 0:
            kw = dict(n_contigs=2, contig_len=(200, 400), depth=(150, 300), multimap=0.8)
        case = fuzzgen.make_case(seed, exotic=0.5 if seed % 4 == 0 else 0.0, **kw)
        with tempfile.TemporaryDirectory() as d:
            fa, sams = case.write(pathlib.Path(d))
            try:
                exp = ("ok", orc.polish(fa, sams, debug=True, **case.opts))
            except Exception as e:
                exp = ("err", e.msg)
            try:
                got = ("ok", pyport.polish(fa, sams, debug=True, **case.opts))
            except pyport.RefError as e:
                got = ("err", str(e))
        assert exp[0] == got[0], (seed, exp[1] if exp[0] == "err" else "", got[1] if got[0] == "err" else "")
        if exp[0] == "ok":
            for k in ("fasta", "debug_tsv", "changed", "zero_depth", "used_total"):
                assert got[1][k] == exp[1][k], (seed, k)
            ok += 1
        else:
            assert got[1] == exp[1], (seed, got[1], exp[1])
            err += 1
    print("polish: %d cases identical (%d of them the same error text)" % (ok + err, err), flush=True)
    nf = 0
    for seed in range(first, first + max(4, n // 25)):
        with tempfile.TemporaryDirectory() as d:
            syn = api.Synth(seed=seed, contig_len=15_000 + 1000 * (seed % 7), depth=20 + seed % 30)
            fa, sams = syn.write(d)
            for orient in ("auto", "fr", "rf", "ff", "rr"):
                kw = dict(orientation=orient)
                try:
                    exp = ("ok", orc.filter(sams[0], sams[1], **kw))
                except Exception as e:
                    exp = ("err", e.msg)
                try:
                    got = ("ok", pyport.filter_sams(sams[0], sams[1], **kw))
                except pyport.RefError as e:
                    got = ("err", str(e))
                assert exp[0] == got[0], (seed, orient, exp[1] if exp[0] == "err" else "", got[1] if got[0] == "err" else "")
                if exp[0] == "ok":
                    for k in ("out1", "out2", "low", "high", "orientation"):
                        assert got[1][k] == exp[1][k], (seed, orient, k)
                nf += 1
    print("filter: %d runs identical" % nf, flush=True)


if __name__ == "__main__":
    main()
