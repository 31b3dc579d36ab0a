import sys, tempfile, pathlib
sys.path.insert(0,'.')
from tests import fuzzgen, oracle_lib
import polypolish_b200 as pp
o = oracle_lib.load()
ctx = pp.Context(0)
for seed in [101, 108, 109, 103]:
    case = fuzzgen.make_case(seed, exotic=0.5 if seed % 4 == 0 else 0.0)
    d = pathlib.Path(tempfile.mkdtemp())
    fa, sams = case.write(d)
    exp = o.polish(fa, sams, debug=True, **case.opts)
    got = ctx.polish_files(fa, sams, **case.opts)
    e = exp['fasta'].decode().split('\n'); g = got.decode().split('\n')
    print('seed', seed, 'equal', exp['fasta']==got, 'nlines', len(e), len(g))
    dbg = exp['debug_tsv'].decode().split('\n')
    for li,(a,b) in enumerate(zip(e,g)):
        if a!=b:
            if a.startswith('>'): print(' header diff', a, '|', b); continue
            k = next((i for i,(x,y) in enumerate(zip(a,b)) if x!=y), min(len(a),len(b)))
            print(' line', li, 'len', len(a), len(b), 'first diff at', k, a[max(0,k-10):k+10], '|', b[max(0,k-10):k+10])
            cname = e[li-1][1:].split()[0]
            rows=[r for r in dbg if r.startswith(cname+'\t')]
            # find debug rows near output index k: approximate by position k
            for r in rows[max(0,k-6):k+6]: print('   ', r)
            break
