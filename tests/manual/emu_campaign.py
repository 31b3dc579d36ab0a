"""Offline campaign: the polish kernels themselves (polish_dev.cuh under tests/emu, one OS thread per CUDA thread) against the oracle
on many more fuzz cases than the default CPU suite runs - plain, exotic alphabets, deep multi-mapped (ordered depth), long reads,
several tiles.  CPU only, ~1 s per case.   python tests/manual/emu_campaign.py [first_seed] [n_cases]"""
import pathlib
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import fuzzgen, oracle_lib           # noqa: E402
from tests.test_emu_polish import check         # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    oracle_lib.build()
    orc = oracle_lib.load()
    done = 0
    for seed in range(first, first + n):
        kind = seed % 6
        if kind == 0:
            case = fuzzgen.make_case(seed, n_contigs=2, contig_len=(200, 400), depth=(150, 300), multimap=0.8, opts=dict(careful=False))
        elif kind == 1:
            case = fuzzgen.make_case(seed, n_contigs=2, contig_len=(2500, 5000), depth=(15, 30), read_len=(200, 900), multimap=0.4, opts=dict(careful=False))
        elif kind == 2:
            case = fuzzgen.make_case(seed, n_contigs=3, contig_len=(1500, 3000), depth=(30, 60), multimap=0.5)
        else:
            case = fuzzgen.make_case(seed, exotic=0.5 if seed % 4 == 0 else 0.0)
        with tempfile.TemporaryDirectory() as d:
            fa, sams = case.write(pathlib.Path(d))
            try:
                check(orc, fa, sams, grid_tiles=1 + seed % 3, **case.opts)
            except AssertionError:
                print("FAILED at seed", seed, flush=True)
                raise
        done += 1
        if done % 50 == 0:
            print(done, "cases identical", flush=True)
    print("emulated kernels == oracle on %d cases (seeds %d..%d)" % (done, first, first + n - 1), flush=True)


if __name__ == "__main__":
    main()
