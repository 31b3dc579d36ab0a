// tok_strip.h — host side of the SAM upload for `polish`: QUAL (column 11) is 45 % of a bwa-mem line and nothing on the
// polish path reads it (alignment.rs:49-98 takes columns 1-4, 6, 10 and the tags), so the reader threads replace it by "*"
// while they stage the text for the PCIe copy.  Pure byte work on whole lines; shared with tests/tok_harness.cpp.
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace tok {

// Position just after the 10th '\t' of src[p, le), or 0 when the line has fewer than 10 tabs.  (SSE2: sixteen bytes per step;
// the ten columns in front of QUAL are short, so ten memchr calls cost more than this one pass.)
inline size_t after_tenth_tab(const uint8_t* src, size_t p, size_t le) {
    int tabs = 0;
    size_t t = p;
#if defined(__SSE2__)
    const __m128i tabv = _mm_set1_epi8('\t');
    while (t + 16 <= le) {
        unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(src + t)), tabv));
        const int c = __builtin_popcount(m);
        if (tabs + c >= 10) {
            for (int skip = 10 - tabs - 1; skip > 0; --skip) m &= m - 1;      // drop the tabs before the wanted one
            return t + (size_t)__builtin_ctz(m) + 1;
        }
        tabs += c;
        t += 16;
    }
#endif
    for (; t < le; ++t)
        if (src[t] == '\t' && ++tabs == 10) return t + 1;
    return 0;
}

// src[0, len) holds whole lines (every line ends with '\n', except possibly the last).  Writes the same lines to dst with
// the QUAL field of every record line that has one replaced by "*"; header lines ('@'), lines with fewer than 11
// columns and lines whose QUAL is shorter than 2 bytes are copied verbatim.  Returns the bytes written (<= len); the number
// of lines, the order of the lines and every byte outside a QUAL field are unchanged.  dst may not overlap src.
inline size_t strip_qual_lines(const uint8_t* src, size_t len, uint8_t* dst) {
    size_t o = 0, p = 0;
    while (p < len) {
        const uint8_t* nl = (const uint8_t*)memchr(src + p, '\n', len - p);
        const size_t le = nl ? (size_t)(nl - src) : len;          // line = [p, le), newline (if any) at le
        const size_t next = nl ? le + 1 : len;
        size_t q0 = 0, q1 = 0;
        bool strip = false;
        if (le > p && src[p] != '@') {
            const size_t t = after_tenth_tab(src, p, le);
            if (t) {
                q0 = t;
                const uint8_t* tb = (const uint8_t*)memchr(src + q0, '\t', le - q0);
                q1 = tb ? (size_t)(tb - src) : le;
                if (!tb && nl && q1 > q0 && src[q1 - 1] == '\r') q1--;      // the '\r' of a CRLF line end is not QUAL
                strip = q1 - q0 >= 2;
            }
        }
        if (strip) {
            memcpy(dst + o, src + p, q0 - p);
            o += q0 - p;
            dst[o++] = '*';
            memcpy(dst + o, src + q1, next - q1);
            o += next - q1;
        } else {
            memcpy(dst + o, src + p, next - p);
            o += next - p;
        }
        p = next;
    }
    return o;
}


// One slice of the stripping upload.  The file is cut into nominal slices of S bytes; slice k = [o, e) owns the LINES that start
// in it, i.e. the bytes [a, b) with a = 1 + the first '\n' at or after o - 1 (0 for the first slice) and b likewise for e (n for
// the last slice).  raw holds the file bytes [rd0, rd1) with rd0 = o - 1 (0 for k = 0) and rd1 = min(n, e + look - 1).
struct SliceOut {
    uint64_t a = 0, b = 0;      // owned byte range of the file
    uint64_t c = 0;             // bytes staged in `pin`: they go to offset a; [a + c, b) becomes filler
    int status = 0;             // 0 staged, 1 nothing owned (a line longer than the slice runs through it), 3 no line end within `look`
    bool ends_file = false;     // b == n
    uint8_t last = '\n';        // last byte of the file (valid when ends_file)
};

inline SliceOut strip_slice(const uint8_t* raw, uint64_t rd0, uint64_t rd1, uint64_t k, uint64_t e, uint64_t n, uint8_t* pin) {
    SliceOut r;
    r.a = 0; r.b = n;
    if (k) {
        const uint8_t* q = (const uint8_t*)memchr(raw, '\n', (size_t)(rd1 - rd0));
        if (!q) { r.status = rd1 == n ? 1 : 3; return r; }       // no line starts between here and the end of the file: nothing owned
        r.a = rd0 + (uint64_t)(q - raw) + 1;
    }
    if (e < n) {
        const uint8_t* q = (const uint8_t*)memchr(raw + (e - 1 - rd0), '\n', (size_t)(rd1 - (e - 1)));
        if (q) r.b = rd0 + (uint64_t)(q - raw) + 1;
        else if (rd1 != n) { r.status = 3; return r; }           // (else: the unterminated last line of the file, b = n)
    }
    if (r.a >= r.b) { r.status = 1; return r; }
    const uint8_t* region = raw + (r.a - rd0);
    const uint64_t rlen = r.b - r.a;
    r.ends_file = r.b == n;
    if (r.ends_file) r.last = region[rlen - 1];
    if (r.ends_file && region[rlen - 1] != '\n') { memcpy(pin, region, (size_t)rlen); r.c = rlen; }   // unterminated last line: verbatim
    else r.c = strip_qual_lines(region, (size_t)rlen, pin);
    return r;
}

// What the device buffer holds for [a, b) after the slice's copies: the staged bytes, then a filler that every parser of the
// polish path skips - nothing (gap 0), an empty line (gap 1), or '@' + blanks + '\n'.  (The CUDA code issues the same
// bytes as one copy, one memset and one 1-byte copy; this host version is what tests/tok_harness.cpp checks.)
inline void apply_slice(const SliceOut& r, const uint8_t* pin, uint8_t* text) {
    memcpy(text + r.a, pin, (size_t)r.c);
    const uint64_t gap = (r.b - r.a) - r.c;
    if (gap == 1) text[r.a + r.c] = '\n';
    else if (gap >= 2) {
        text[r.a + r.c] = '@';
        memset(text + r.a + r.c + 1, ' ', (size_t)(gap - 2));
        text[r.b - 1] = '\n';
    }
}

}  // namespace tok
