"""Small random FASTA + SAM cases for oracle-vs-GPU parity (adversarial by construction).

Reads are built by walking the draft and applying per-position "truth edits" (substitutions, deletions,
insertions, N bases), so that many positions really change under the vote; on top of that the generator throws in
everything the reference's text/grouping code has to cope with: multi-mapped groups with SEQ="*" secondaries on
both strands (non-dyadic k), soft clips and NM > max (discarded), ZP:Z:fail tags, unaligned records inside a
group, =/X CIGARs, D directly followed by I, homopolymer tails, lower-case SEQ, header / empty lines, CRLF,
alignments touching contig ends, IUPAC and '-' characters in the draft, several SAM files.
"""
import random

COMP = {"A": "T", "T": "A", "G": "C", "C": "G", "N": "N", "R": "Y", "Y": "R", "S": "S", "W": "W", "K": "M", "M": "K",
        "B": "V", "V": "B", "D": "H", "H": "D", ".": ".", "-": "-", "?": "?"}


def revcomp(s):
    return "".join(COMP.get(c, "N") for c in reversed(s))


def merge_ops(ops):
    out = []
    for op, n in ops:
        if n == 0:
            continue
        if out and out[-1][0] == op:
            out[-1][1] += n
        else:
            out.append([op, n])
    return out


def cigar_str(ops):
    return "".join(f"{n}{op}" for op, n in ops)


class Case:
    def __init__(self, fasta_text, sam_texts, opts):
        self.fasta_text = fasta_text
        self.sam_texts = sam_texts
        self.opts = opts

    def write(self, d):
        fa = d / "asm.fasta"
        fa.write_bytes(self.fasta_text.encode("latin-1"))
        sams = []
        for i, t in enumerate(self.sam_texts):
            p = d / f"reads_{i + 1}.sam"
            p.write_bytes(t.encode("latin-1"))
            sams.append(p)
        return fa, sams


def make_case(seed, n_contigs=None, contig_len=(120, 600), depth=(20, 60), read_len=(40, 150), exotic=0.0,
              iupac_draft=0.02, n_files=None, multimap=0.25, opts=None):
    rng = random.Random(seed)
    n_contigs = n_contigs or rng.randint(1, 3)
    n_files = n_files or rng.randint(1, 2)
    contigs = []
    for c in range(n_contigs):
        L = rng.randint(*contig_len)
        seq = []
        while len(seq) < L:
            if rng.random() < 0.08:                       # homopolymer stretches
                seq.extend(rng.choice("ACGT") * rng.randint(3, 9))
            else:
                seq.append(rng.choice("ACGT"))
        seq = seq[:L]
        for i in range(L):
            r = rng.random()
            if r < iupac_draft * 0.5:
                seq[i] = rng.choice("NRYSWKMBVDH")
            elif r < iupac_draft * 0.6 and exotic:
                seq[i] = "-"
            elif r < iupac_draft * 0.7 and exotic:
                seq[i] = rng.choice("XU*")
        name = f"ctg{c + 1}"
        desc = rng.choice(["", "len=%d" % L, "some description  here"])
        # truth edits
        edits = {}
        for i in range(5, L - 5):
            r = rng.random()
            if r < 0.012:
                edits[i] = ("sub", rng.choice("ACGT"))
            elif r < 0.018:
                edits[i] = ("del",)
            elif r < 0.024:
                edits[i] = ("ins", "".join(rng.choice("ACGT") for _ in range(rng.choice([1, 1, 1, 2, 3]))))
            elif r < 0.027:
                edits[i] = ("sub", "N")
            elif r < 0.029:
                edits[i] = ("delins", rng.choice("ACGT"))          # D directly followed by I
            elif r < 0.030 and exotic:
                edits[i] = ("sub", rng.choice("X.-?="))
        contigs.append(dict(name=name, desc=desc, seq="".join(seq), edits=edits))

    def fasta_text():
        out = []
        for c in contigs:
            h = ">" + c["name"] + ((" " + c["desc"]) if c["desc"] else "")
            out.append(h)
            s = c["seq"]
            if rng.random() < 0.5:
                s = "".join(ch.lower() if rng.random() < 0.1 else ch for ch in s)
            w = rng.choice([60, 80, 10 ** 9])
            out.extend(s[i:i + w] for i in range(0, len(s), w))
            if rng.random() < 0.3:
                out.append("")
        return "\n".join(out) + "\n"

    def build_read(c, start, want_len, use_eqx):
        """Walk contig c from `start`; returns (ops, read_seq, nm, ref_span)."""
        seq, edits = c["seq"], c["edits"]
        ops, read, nm = [], [], 0
        p = start
        while len(read) < want_len and p < len(seq):
            e = edits.get(p) if rng.random() < 0.9 else None
            base = seq[p]
            if e is None and rng.random() < 0.004:                   # sequencing error
                e = ("sub", rng.choice("ACGT"))
            if e is None:
                ops.append(("=" if use_eqx else "M", 1)); read.append(base)
            elif e[0] == "sub":
                ops.append(("X" if use_eqx and e[1] != base else ("=" if use_eqx else "M"), 1)); read.append(e[1])
                nm += e[1] != base
            elif e[0] == "del":
                if not ops:                                           # cannot start with D: treat as match
                    ops.append(("=" if use_eqx else "M", 1)); read.append(base)
                else:
                    ops.append(("D", 1)); nm += 1
            elif e[0] == "ins":
                ops.append(("=" if use_eqx else "M", 1)); read.append(base)
                ops.append(("I", len(e[1]))); read.extend(e[1]); nm += len(e[1])
            elif e[0] == "delins":
                if not ops:
                    ops.append(("=" if use_eqx else "M", 1)); read.append(base)
                else:
                    ops.append(("D", 1)); ops.append(("I", 1)); read.append(e[1]); nm += 2
            p += 1
        ops = merge_ops(ops)
        # do not end on D (make it end on a match by dropping trailing D / I)
        while ops and ops[-1][0] in "DI":
            op, n = ops.pop()
            if op == "I":
                del read[-n:]
            else:
                p -= n
        return ops, "".join(read), nm, p - start

    sam_texts = []
    rid = 0
    for f in range(n_files):
        lines = []
        if rng.random() < 0.8:
            lines.append("@HD\tVN:1.6\tSO:unsorted")
            for c in contigs:
                lines.append(f"@SQ\tSN:{c['name']}\tLN:{len(c['seq'])}")
        total_bp = sum(len(c["seq"]) for c in contigs)
        n_reads = max(3, int(total_bp * rng.uniform(*depth) / ((read_len[0] + read_len[1]) / 2) / n_files))
        for _ in range(n_reads):
            rid += 1
            qname = f"read{rid}"
            c = rng.choice(contigs)
            want = rng.randint(*read_len)
            L = len(c["seq"])
            start = rng.choice([0, max(0, L - want), rng.randint(0, max(0, L - 20))]) if rng.random() < 0.2 else rng.randint(0, max(0, L - 20))
            use_eqx = rng.random() < 0.15
            ops, seq, nm, span = build_read(c, start, want, use_eqx)
            if not seq:
                continue
            if rng.random() < 0.10:                                   # homopolymer tail
                t = rng.choice("ACGT") * rng.randint(2, 6)
                k = min(len(t), max(0, L - (start + span)))
                if k:
                    seq += t[:k]; ops = merge_ops(ops + [["=" if use_eqx else "M", k]]); span += k
                    nm += sum(1 for i in range(k) if c["seq"][start + span - k + i] != t[i])
            flag = rng.choice([0, 0, 16])
            r = rng.random()
            if r < 0.04:                                              # soft clip at an end -> discarded
                n = rng.randint(1, 5)
                if rng.random() < 0.5:
                    ops = [["S", n]] + ops; seq = "".join(rng.choice("ACGT") for _ in range(n)) + seq
                else:
                    ops = ops + [["S", n]]; seq = seq + "".join(rng.choice("ACGT") for _ in range(n))
            elif r < 0.07:
                nm += 11                                              # NM > max_errors
            tags = [f"NM:i:{nm}"]
            if rng.random() < 0.5:
                tags = [f"AS:i:{rng.randint(0, 150)}"] + tags + [f"XS:i:{rng.randint(0, 150)}"]
            if rng.random() < 0.03:
                tags.append(rng.choice(["ZP:Z:fail", "zp:z:FAIL"]))
            seq_out = seq.lower() if rng.random() < 0.05 else seq
            qual = "I" * len(seq)
            recs = [(flag, c["name"], start + 1, cigar_str(ops), seq_out, qual, tags)]
            if rng.random() < multimap:                               # secondaries, SEQ="*"
                for _ in range(rng.choice([1, 1, 2, 2, 3, 4, 6])):
                    c2 = rng.choice(contigs)
                    L2 = len(c2["seq"])
                    rl = len(seq)
                    kind = rng.random()
                    if kind < 0.6 and L2 > rl:
                        cg = f"{rl}M"; st = rng.randint(0, L2 - rl); nm2 = rng.randint(0, 6)
                    elif kind < 0.8 and L2 > rl + 2 and rl > 8:
                        a = rng.randint(2, rl - 3)
                        cg = f"{a}M1D{rl - a}M"; st = rng.randint(0, L2 - rl - 1); nm2 = rng.randint(1, 6)
                    elif kind < 0.9 and rl > 8:
                        a = rng.randint(2, rl - 4)
                        cg = f"{a}M1I{rl - a - 1}M"; st = rng.randint(0, max(0, L2 - rl)); nm2 = rng.randint(1, 6)
                        if st + rl - 1 > L2:
                            continue
                    else:
                        n = rng.randint(1, min(5, rl - 1))
                        if L2 < rl:
                            continue
                        cg = f"{n}S{rl - n}M"; st = rng.randint(0, L2 - (rl - n)); nm2 = rng.randint(0, 3)
                    fl2 = 256 | rng.choice([0, 16])
                    t2 = [f"NM:i:{nm2}"]
                    if rng.random() < 0.05:
                        t2.append("ZP:Z:fail")
                    if rng.random() < 0.15:                           # secondary that carries its own SEQ
                        s2 = "".join(rng.choice("ACGT") for _ in range(rl)); q2 = "I" * rl
                    else:
                        s2, q2 = "*", "*"
                    recs.append((fl2, c2["name"], st + 1, cg, s2, q2, t2))
                if rng.random() < 0.3:                                # source sequence not on the first record
                    rng.shuffle(recs)
            for (fl, rn, pos, cg, s, q, tg) in recs:
                lines.append("\t".join([qname, str(fl), rn, str(pos), "60", cg, "*", "0", "0", s, q] + tg))
                if rng.random() < 0.03:                               # unaligned record inside the group
                    lines.append("\t".join([qname, "4", "*", "0", "0", "*", "*", "0", "0", "ACGT", "IIII"]))
            if rng.random() < 0.02:
                lines.append("")
        eol = "\r\n" if rng.random() < 0.15 else "\n"
        sam_texts.append(eol.join(lines) + (eol if rng.random() < 0.9 else ""))

    o = dict(fraction_invalid=0.2, fraction_valid=0.5, max_errors=10, min_depth=5, careful=False)
    r = rng.random()
    if r < 0.15:
        o.update(careful=True)
    elif r < 0.3:
        o.update(min_depth=rng.choice([0, 1, 2, 12]))
    elif r < 0.45:
        o.update(fraction_invalid=rng.choice([0.01, 0.1, 0.3]), fraction_valid=rng.choice([0.4, 0.6, 0.9]))
    elif r < 0.55:
        o.update(max_errors=rng.choice([0, 2, 30]))
    if opts:
        o.update(opts)
    return Case(fasta_text(), sam_texts, o)
