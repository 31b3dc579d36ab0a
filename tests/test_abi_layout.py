"""The boundary documentation is checked against the boundary: every `#[repr(C)]` struct INTEGRATION.md shows a Rust host
must have the size, field names and field offsets of its pp_abi.h counterpart (gcc's sizeof / offsetof).  CPU only."""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pp_abi.h")
DOC = os.path.join(ROOT, "INTEGRATION.md")

RUST_TO_C = {"PpAlignments": "pp_alignments", "PpContigs": "pp_contigs", "PpPolishParams": "pp_polish_params", "PpTiming": "pp_timing",
             "PpPolishResult": "pp_polish_result", "PpFilterMate": "pp_filter_mate", "PpFilterParams": "pp_filter_params",
             "PpFilterResult": "pp_filter_result", "PpTokStats": "pp_tok_stats", "PpDebugPos": "pp_debug_pos", "PpDebugNode": "pp_debug_node"}
PRIM = {"u8": 1, "i8": 1, "u16": 2, "i16": 2, "u32": 4, "i32": 4, "f32": 4, "u64": 8, "i64": 8, "f64": 8, "usize": 8, "isize": 8}


def c_structs():
    """{struct name: [field names]} of every `typedef struct { ... } name;` in the header."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const uint32_t *a, *b" / "uint64_t pairs[4]" / "double x, y"
            first, *rest = decl.split(",")
            names = [re.sub(r"\[.*", "", first.split()[-1]).lstrip("*")] + [re.sub(r"\[.*", "", r.strip()).lstrip("*") for r in rest]
            fields += names
        out[name] = fields
    return out


def c_layout(tmp_path):
    structs = c_structs()
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "pp_abi.h"', "int main(void) {", 'printf("{");']
    first = True
    for name, fields in structs.items():
        src.append('printf("%s\\"%s\\": {\\"size\\": %%zu, \\"fields\\": {", sizeof(%s));' % ("" if first else ", ", name, name))
        first = False
        for i, f in enumerate(fields):
            src.append('printf("%s\\"%s\\": %%zu", offsetof(%s, %s));' % ("" if i == 0 else ", ", f, name, f))
        src.append('printf("}}");')
    src += ['printf("}\\n");', "return 0; }"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    return json.loads(subprocess.check_output([str(exe)]).decode())


def rust_structs():
    text = open(DOC).read()
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\]\s*pub struct (\w+)\s*\{(.*?)\}", text, flags=re.S):
        fields = []
        depth, cur = 0, ""
        for ch in body:                       # split on commas outside [..]
            if ch == "[":
                depth += 1
            elif ch == "]":
                depth -= 1
            if ch == "," and depth == 0:
                fields.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            fields.append(cur)
        parsed = []
        for f in fields:
            m = re.match(r"\s*pub\s+(\w+)\s*:\s*(.+?)\s*$", f, flags=re.S)
            assert m, f
            parsed.append((m.group(1), m.group(2)))
        out[name] = parsed
    return out


def rust_size_align(ty, structs):
    ty = ty.strip()
    if ty.startswith("*"):
        return 8, 8
    m = re.match(r"\[\s*(.+?)\s*;\s*(\d+)\s*\]$", ty)
    if m:
        s, a = rust_size_align(m.group(1), structs)
        return s * int(m.group(2)), a
    if ty in PRIM:
        return PRIM[ty], PRIM[ty]
    return rust_layout(ty, structs)[:2]


def rust_layout(name, structs):
    off, align, offsets = 0, 1, {}
    for fname, ty in structs[name]:
        s, a = rust_size_align(ty, structs)
        off = (off + a - 1) // a * a
        offsets[fname] = off
        off += s
        align = max(align, a)
    return (off + align - 1) // align * align, align, offsets


def test_documented_rust_structs_match_the_header(tmp_path):
    c = c_layout(tmp_path)
    r = rust_structs()
    assert set(RUST_TO_C) <= set(r), "INTEGRATION.md must define " + ", ".join(sorted(set(RUST_TO_C) - set(r)))
    for rname, cname in RUST_TO_C.items():
        size, _, offsets = rust_layout(rname, r)
        assert cname in c, cname
        assert size == c[cname]["size"], (rname, size, c[cname]["size"])
        assert list(offsets) == list(c[cname]["fields"]), (rname, list(offsets), list(c[cname]["fields"]))
        assert offsets == c[cname]["fields"], (rname, offsets, c[cname]["fields"])


def test_python_mirror_matches_the_header(tmp_path):
    """api.py's ctypes structures are the same boundary seen from Python."""
    import ctypes as C
    from polypolish_b200 import api
    c = c_layout(tmp_path)
    pairs = {"pp_alignments": api.Alignments, "pp_contigs": api.Contigs, "pp_polish_params": api.PolishParams, "pp_timing": api.Timing,
             "pp_polish_result": api.PolishResult, "pp_filter_mate": api.FilterMate, "pp_filter_params": api.FilterParams,
             "pp_filter_result": api.FilterResult, "pp_tok_stats": api.TokStats, "pp_synth_params": api.SynthParams}
    for cname, cls in pairs.items():
        assert C.sizeof(cls) == c[cname]["size"], cname
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == c[cname]["fields"][fname], (cname, fname)
