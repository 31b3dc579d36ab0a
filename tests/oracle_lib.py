"""ctypes binding of the CPU oracle (oracle/polypolish_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
Never imported from polypolish_b200/.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "build", "liboracle.so")
BIN = os.path.join(ORACLE_DIR, "build", "polypolish_oracle")


def build(force=False):
    src = os.path.join(ORACLE_DIR, "polypolish_oracle.cpp")
    stale = (not os.path.exists(LIB) or not os.path.exists(BIN)
             or os.path.getmtime(LIB) < os.path.getmtime(src))
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return LIB


class PolishResult(C.Structure):
    _fields_ = [("fasta", C.c_void_p), ("fasta_len", C.c_uint64),
                ("debug_tsv", C.c_void_p), ("debug_len", C.c_uint64),
                ("alignment_total", C.c_uint64), ("used_total", C.c_uint64),
                ("n_contigs", C.c_uint64),
                ("changed", C.POINTER(C.c_uint64)), ("zero_depth", C.POINTER(C.c_uint64)),
                ("new_length", C.POINTER(C.c_uint64)), ("total_depth", C.POINTER(C.c_double)),
                ("secs_load", C.c_double), ("secs_parse", C.c_double),
                ("secs_scatter", C.c_double), ("secs_vote", C.c_double)]


class FilterResult(C.Structure):
    _fields_ = [("out1", C.c_void_p), ("out1_len", C.c_uint64), ("out2", C.c_void_p), ("out2_len", C.c_uint64),
                ("low", C.c_uint32), ("high", C.c_uint32), ("orientation", C.c_int),
                ("pairs", C.c_uint64 * 4), ("before_count", C.c_uint64), ("after_count", C.c_uint64)]


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


STATUS = ["low_depth", "none", "multiple", "too_close", "kept", "changed"]
ORIENT = ["fr", "rf", "ff", "rr"]


class Oracle:
    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        L.orc_bankers_rounding.restype = C.c_uint32
        L.orc_bankers_rounding.argtypes = [C.c_double]
        L.orc_get_percentile.restype = C.c_uint32
        L.orc_get_percentile.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_double]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_get_expanded_cigar.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.orc_vote.argtypes = [C.c_char, C.c_char_p, C.POINTER(C.c_double), C.c_uint64, C.c_uint32, C.c_double,
                               C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                               C.POINTER(C.c_void_p)]
        L.orc_polish.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_double, C.c_double, C.c_uint32,
                                 C.c_uint32, C.c_int, C.c_int, C.POINTER(PolishResult), C.c_char_p, C.c_uint64]
        L.orc_filter.argtypes = [C.c_char_p] * 5 + [C.c_double, C.c_double, C.POINTER(FilterResult), C.c_char_p,
                                                    C.c_uint64]

    def _take(self, p):
        s = C.string_at(p)
        self.lib.orc_free(p)
        return s.decode("latin-1")

    def bankers_rounding(self, x):
        return self.lib.orc_bankers_rounding(x)

    def reverse_complement(self, s):
        out = C.create_string_buffer(len(s) + 1)
        self.lib.orc_reverse_complement(s.encode("latin-1"), out)
        return out.value.decode("latin-1")

    def expanded_cigar(self, cigar, n=0):
        p = C.c_void_p()
        rc = self.lib.orc_get_expanded_cigar(cigar.encode(), n, C.byref(p))
        if rc:
            return None
        return self._take(p)

    def alignment_new(self, line):
        rs, re_, nm, qc = C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_int()
        err = C.create_string_buffer(512)
        rc = self.lib.orc_alignment_new(line.encode("latin-1"), C.byref(rs), C.byref(re_), C.byref(nm), C.byref(qc),
                                        err, 512)
        if rc:
            raise OracleError(rc, err.value.decode())
        return rs.value, re_.value, nm.value, bool(qc.value)

    def orientation(self, line1, line2):
        out = C.create_string_buffer(3)
        ins = C.c_uint32()
        rc = self.lib.orc_get_orientation(line1.encode(), line2.encode(), out, C.byref(ins))
        assert rc == 0
        return out.value.decode(), ins.value

    def auto_orientation(self, counts):
        arr = (C.c_uint64 * 4)(*counts)
        i = self.lib.orc_auto_determine_orientation(arr)
        return None if i < 0 else ORIENT[i]

    def percentile(self, sorted_list, p):
        arr = (C.c_uint32 * len(sorted_list))(*sorted_list)
        return self.lib.orc_get_percentile(arr, len(sorted_list), p)

    def vote(self, original, adds, min_depth=5, fv=0.5, fi=0.2):
        """adds: list of (seq, contribution). Returns (new_base, status, count_str, debug_line)."""
        blob = b"".join(s.encode("latin-1") + b"\0" for s, _ in adds)
        contribs = (C.c_double * max(1, len(adds)))(*[c for _, c in adds])
        nb, cs, dl = C.c_void_p(), C.c_void_p(), C.c_void_p()
        st = C.c_int()
        self.lib.orc_vote(original.encode("latin-1"), blob, contribs, len(adds), min_depth, fv, fi, C.byref(nb),
                          C.byref(st), C.byref(cs), C.byref(dl))
        return self._take(nb), STATUS[st.value], self._take(cs), self._take(dl)

    def load_fasta(self, path):
        p = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.orc_load_fasta(str(path).encode(), C.byref(p), err, 512)
        if rc:
            raise OracleError(rc, err.value.decode())
        txt = self._take(p)
        return [tuple(l.split("\t")) for l in txt.split("\n") if l]

    def walk(self, cigar, seq):
        cap = 4 * (len(seq) + 16) + 64
        pairs = (C.c_uint64 * (2 * cap))()
        n = C.c_uint64()
        err = C.create_string_buffer(512)
        rc = self.lib.orc_walk(cigar.encode(), seq.encode("latin-1"), pairs, cap, C.byref(n), err, 512)
        if rc:
            raise OracleError(rc, err.value.decode())
        return [(pairs[2 * i], pairs[2 * i + 1]) for i in range(n.value)]

    def polish(self, assembly, sams, fraction_invalid=0.2, fraction_valid=0.5, max_errors=10, min_depth=5,
               careful=False, debug=False):
        res = PolishResult()
        err = C.create_string_buffer(1024)
        arr = (C.c_char_p * max(1, len(sams)))(*[str(s).encode() for s in sams])
        rc = self.lib.orc_polish(str(assembly).encode(), arr, len(sams), fraction_invalid, fraction_valid, max_errors,
                                 min_depth, int(careful), int(debug), C.byref(res), err, 1024)
        if rc:
            raise OracleError(rc, err.value.decode())
        n = res.n_contigs
        out = dict(fasta=C.string_at(res.fasta, res.fasta_len),
                   debug_tsv=C.string_at(res.debug_tsv, res.debug_len) if debug else b"",
                   alignment_total=res.alignment_total, used_total=res.used_total,
                   changed=[res.changed[i] for i in range(n)], zero_depth=[res.zero_depth[i] for i in range(n)],
                   new_length=[res.new_length[i] for i in range(n)],
                   total_depth=[res.total_depth[i] for i in range(n)],
                   secs=dict(load=res.secs_load, parse=res.secs_parse, scatter=res.secs_scatter, vote=res.secs_vote))
        self.lib.orc_polish_result_free(C.byref(res))
        return out

    def filter(self, in1, in2, orientation="auto", low=0.1, high=99.9, out1="<out1>", out2="<out2>"):
        res = FilterResult()
        err = C.create_string_buffer(1024)
        rc = self.lib.orc_filter(str(in1).encode(), str(in2).encode(), str(out1).encode(), str(out2).encode(),
                                 orientation.encode(), low, high, C.byref(res), err, 1024)
        if rc:
            raise OracleError(rc, err.value.decode())
        out = dict(out1=C.string_at(res.out1, res.out1_len), out2=C.string_at(res.out2, res.out2_len), low=res.low,
                   high=res.high, orientation=ORIENT[res.orientation] if res.orientation >= 0 else None,
                   pairs=list(res.pairs), before_count=res.before_count, after_count=res.after_count)
        self.lib.orc_filter_result_free(C.byref(res))
        return out


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = Oracle(build())
    return _cached
