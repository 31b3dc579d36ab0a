"""CPU unit test of the 4-bit sequence helpers the scatter kernel is built on (nib_utils.h), against plain Python."""
import ctypes as C
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H():
    out = os.path.join(ROOT, "build", "nib_harness.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "nib_harness.cpp")])
    return C.CDLL(out)


def pack(codes, extra_words=4):
    n = (len(codes) + 15) // 16 + extra_words
    w = [0] * n
    for i, c in enumerate(codes):
        w[i // 16] |= c << (4 * (i % 16))
    return (C.c_uint64 * n)(*w)


def nibbles(lo, hi):
    return [(lo >> (4 * i)) & 15 for i in range(16)] + [(hi >> (4 * i)) & 15 for i in range(16)]


def brev4(c):
    return int(f"{c:04b}"[::-1], 2)


def test_load_nib32(H):
    rng = random.Random(1)
    codes = [rng.randint(1, 15) for _ in range(400)]
    w = pack(codes)
    out = (C.c_uint64 * 2)()
    for start in list(range(0, 70)) + [rng.randint(0, 360) for _ in range(200)]:
        H.h_load_nib32(w, start, out)
        assert nibbles(out[0], out[1]) == (codes + [0] * 64)[start:start + 32], start


@pytest.mark.parametrize("rc", [0, 1])
def test_load_read32(H, rc):
    rng = random.Random(2 + rc)
    out = (C.c_uint64 * 2)()
    for length in [1, 5, 16, 31, 32, 33, 64, 100, 149, 150, 151, 250, 257]:
        codes = [rng.randint(1, 15) for _ in range(length)]
        eff = [brev4(c) for c in reversed(codes)] if rc else codes
        w = pack(codes)
        for ri in sorted(set([0, 1, 15, 16, 17, 31, 32, 33, max(0, length - 33), max(0, length - 32), max(0, length - 31), max(0, length - 1)] +
                             [rng.randint(0, length - 1) for _ in range(20)])):
            if ri >= length:
                continue
            H.h_load_read32(w, length, rc, ri, out)
            got = nibbles(out[0], out[1])
            valid = min(32, length - ri)
            assert got[:valid] == eff[ri:ri + valid], (length, ri)


def test_mismatch_masks(H):
    rng = random.Random(3)
    out = (C.c_uint64 * 2)()
    for _ in range(500):
        r = [rng.randint(1, 15) for _ in range(32)]
        d = [x if rng.random() < 0.8 else rng.randint(0, 15) for x in r]
        vc = rng.randint(1, 32)
        rw, dw = pack(r, 0), pack(d, 0)
        H.h_mismatch(rw, dw, vc, out)
        m = [(out[0] >> (4 * i)) & 1 for i in range(16)] + [(out[1] >> (4 * i)) & 1 for i in range(16)]
        assert m == [1 if (i < vc and r[i] != d[i]) else 0 for i in range(32)]
