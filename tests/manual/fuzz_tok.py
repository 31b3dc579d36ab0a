"""Offline fuzz campaign: the device tokeniser logic (tests/tok_harness.cpp) against the host packer on random, mutated SAM texts.
usage: python tests/manual/fuzz_tok.py <seed> <trials>   (needs build/tok_harness.so: run tests/test_tok_cpu.py once).  51,000 texts: 0 disagreements."""
import os, random, sys, pathlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_tok_cpu import *
import subprocess, ctypes as C
out = os.path.join(ROOT, "build", "tok_harness.so")
Hh = C.CDLL(out)
seed0 = int(sys.argv[1]); n_trials = int(sys.argv[2])
d = pathlib.Path(tempfile.mkdtemp())
alphabet = "\t\t\t\t0123456789MIDNSHP=X*ACGTNacgtn:@+-\r zpZPfailNM\x08\x01\x89\xff\n"
bad = 0
for t in range(n_trials):
    rng = random.Random(seed0 * 1000003 + t)
    n_lines = rng.randint(1, 12)
    base = []
    for i in range(n_lines):
        name = rng.choice(["q%d" % (i // 2), "q%d" % (i // 3), "", "x"])
        seq = rng.choice(["ACGT", "*", "acgn", "ACGTACGTACGTACGTACGTACGTACGTACGTACGTA", "", "AC.T", "A=CG"])
        base.append(line(name, rng.choice([0, 16, 4, 256, 272, 20]), rng.choice(["c1", "c2", "c10", "zz", "*", ""]), rng.choice([0, 1, 7, 30]),
                         rng.choice(["4M", "2M1I1M", "1M1D3M", "2S2M", "4=", "35M", "*", "0M4M", "1M"]), seq,
                         rng.choice([("NM:i:%d" % rng.randint(0, 12),), ("NM:i:1", "ZP:Z:fail"), (), ("XX:i:1", "NM:i:+3", ""), ("zp:z:FAIL", "NM:i:0")])))
    if rng.random() < 0.15:
        base.insert(rng.randrange(len(base) + 1), rng.choice(["@HD\tVN:1", "", "@"]) + "\n")
    lines = list(base)
    for _ in range(rng.choice([0, 0, 1, 1, 2, 3])):
        j = rng.randrange(len(lines)); s = list(lines[j])
        if len(s) < 2: continue
        k = rng.randrange(len(s) - 1); op = rng.random()
        if op < 0.4: s[k] = rng.choice(alphabet)
        elif op < 0.7: del s[k]
        else: s.insert(k, rng.choice(alphabet))
        lines[j] = "".join(s)
    text = "".join(lines)
    if rng.random() < 0.2: text = text.replace("\n", "\r\n")
    if rng.random() < 0.2 and text.endswith("\n"): text = text[:-1]
    tb = text.encode("latin-1")
    careful = rng.random() < 0.3
    try:
        f, p = host_pack(d, FA, [tb], careful)
        host = p.arrays()
    except pp.PolypolishError:
        host = None
    rc, dev, _ = tok_cpu(Hh, NAMES, [tb], careful, 4)
    try:
        if host is None:
            assert rc in (PP_TOK_HOST, PP_TOK_NEED8), ("host fails, harness rc", rc)
        elif rc == PP_OK:
            assert host["seq_bits"] == 4; assert_same(dev, host)
        elif rc == PP_TOK_NEED8:
            assert host["seq_bits"] == 8
            rc8, dev8, _ = tok_cpu(Hh, NAMES, [tb], careful, 8)
            assert rc8 == PP_OK; assert_same(dev8, host)
        else:
            raise AssertionError("harness refused a text the host packer accepts")
    except AssertionError as e:
        bad += 1
        print("DISAGREE", seed0, t, careful, e, repr(tb)[:600]); 
        if bad > 5: break
print("done", seed0, n_trials, "bad", bad)
