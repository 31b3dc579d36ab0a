"""Offline campaign: the QUAL-stripping upload emulated with random tiny slice sizes (tests/tok_harness.cpp upload_emulate), tokenised by the
harness and compared with the host packer on the original text.  usage: python tests/manual/fuzz_upload.py <seed_from> <seed_to>.  800 cases: 0 disagreements."""
import os, random, sys, pathlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_tok_cpu import *
import ctypes as C
Hh = C.CDLL(os.path.join(ROOT, "build", "tok_harness.so"))
d = pathlib.Path(tempfile.mkdtemp())
bad = 0; n3 = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = random.Random(seed)
    case = fuzzgen.make_case(seed, exotic=0.3 if seed % 7 == 0 else 0.0, multimap=0.4)
    texts = [t.encode("latin-1") for t in case.sam_texts]
    if rng.random() < 0.3 and not any(b"\r" in t for t in texts): texts = [t.replace(b"\n", b"\r\n") for t in texts]
    if rng.random() < 0.3 and not any(b"\r" in t for t in texts): texts = [t[:-1] for t in texts]
    careful = case.opts["careful"]
    try:
        f, p = host_pack(d, case.fasta_text, texts, careful); host = p.arrays()
    except pp.PolypolishError:
        continue
    S = rng.choice([33, 64, 100, 257, 1024, 5000]); look = rng.choice([64, 300, 1000, 8000])
    dev_texts = []
    for t in texts:
        rc, out, last, sent = upload_emulate(Hh, t, S, look)
        if rc == 3: n3 += 1; dev_texts.append(t); continue
        assert rc == 0 and b"\xee" not in out and len(out) == len(t) and last == t[-1], (seed, S, look)
        dev_texts.append(out)
    rc, dev, _ = tok_cpu(Hh, f.names, dev_texts, careful, int(host["seq_bits"]))
    try:
        assert rc == PP_OK; assert_same(dev, host)
    except AssertionError as e:
        bad += 1; print("DISAGREE", seed, S, look, e)
print("done", sys.argv[1:], "bad", bad, "fallbacks", n3)
