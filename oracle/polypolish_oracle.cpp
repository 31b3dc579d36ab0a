// polypolish_oracle.cpp — CPU restatement of the Polypolish v0.6.1 alignment-pileup-and-vote path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  This file is the parity oracle and the timed CPU baseline.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or run it.  Nothing under polypolish_b200/ links or calls it; the product path
// fails loudly when the CUDA library is missing.
//
// PARITY STATUS: "parity unpinned" end to end.  The reference is Rust (Cargo.toml:1-19, 78 crates.io
// dependencies) and no Rust toolchain or prebuilt binary exists in this environment, so the reference
// itself cannot be run to generate golden SAM->FASTA vectors.  The restatement is pinned against every
// known-answer vector carried by the reference's own 21 unit tests (tests/test_oracle_kat.py,
// SURVEY.md Appendix B); everything those tests do not cover follows the Rust source line by line,
// each function below citing the file:line it restates.
//
// The restatement is deliberately structure-faithful (one heap string per SAM line, a vector of
// fields, an expanded CIGAR string, an array-of-structs pileup with a string->count hash map, a
// sequential f64 depth sum, single thread) so that its wall time is a fair stand-in for the
// reference's own CPU path.  It is reported as "C++ restatement of the reference, 1 thread",
// never as the Rust binary.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, no fast-math: IEEE double semantics matter).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>
#include <zlib.h>

namespace {

// misc.rs:29-33 quit_with_error -> exit(1).  Here: exception carried to the API boundary.
struct OracleError : std::runtime_error {
    int code;
    OracleError(const std::string& m, int c = 1) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void quit_with_error(const std::string& text) { throw OracleError(text, 1); }
// Rust panics (unwrap on None/Err, index out of bounds) -> exit code 101 in the reference.
[[noreturn]] void rust_panic(const std::string& text) { throw OracleError("panic: " + text, 101); }

using Clock = std::chrono::steady_clock;
inline double secs_since(Clock::time_point t0) {
    return std::chrono::duration<double>(Clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------------
// misc.rs
// ---------------------------------------------------------------------------------------------

// misc.rs:208-215 bankers_rounding.  `float as u32` saturates in Rust (NaN -> 0).
uint32_t bankers_rounding(double x) {
    uint32_t rounded_down;
    if (!(x == x)) rounded_down = 0;
    else if (x <= 0.0) rounded_down = 0;
    else if (x >= 4294967295.0) rounded_down = 4294967295u;
    else rounded_down = (uint32_t)x;
    double fract = x - std::trunc(x);           // f64::fract
    if (fract < 0.5) return rounded_down;
    if (fract > 0.5) return rounded_down + 1;
    return rounded_down + (rounded_down & 1u);
}

// misc.rs:170-182 complement_base
char complement_base(char b) {
    switch (b) {
        case 'A': return 'T'; case 'T': return 'A'; case 'G': return 'C'; case 'C': return 'G';
        case 'a': return 't'; case 't': return 'a'; case 'g': return 'c'; case 'c': return 'g';
        case 'N': return 'N'; case 'n': return 'n';
        case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W';
        case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
        case 'r': return 'y'; case 'y': return 'r'; case 's': return 's'; case 'w': return 'w';
        case 'k': return 'm'; case 'm': return 'k';
        case 'b': return 'v'; case 'v': return 'b'; case 'd': return 'h'; case 'h': return 'd';
        case '.': return '.'; case '-': return '-'; case '?': return '?';
        default: return 'N';
    }
}

// misc.rs:185-191 reverse_complement (ASCII; the reference iterates chars, SAM SEQ is ASCII)
std::string reverse_complement(const std::string& seq) {
    std::string out;
    out.reserve(seq.size());
    for (size_t i = seq.size(); i-- > 0;) out.push_back(complement_base(seq[i]));
    return out;
}

void make_ascii_uppercase(std::string& s) {
    for (char& c : s) if (c >= 'a' && c <= 'z') c = (char)(c - 32);
}

// Rust str::lines(): split on '\n', strip one trailing '\r'; a final unterminated line is a line.
template <class F> void for_each_line(const std::string& data, F&& f) {
    size_t pos = 0, n = data.size();
    while (pos < n) {
        size_t nl = data.find('\n', pos);
        size_t end = (nl == std::string::npos) ? n : nl;
        size_t e2 = end;
        if (nl != std::string::npos && e2 > pos && data[e2 - 1] == '\r') e2--;
        f(std::string(data, pos, e2 - pos));    // one heap String per line, like BufRead::lines()
        if (nl == std::string::npos) break;
        pos = nl + 1;
    }
}

bool read_whole_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string buf;
    char tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
    fclose(f);
    out.swap(buf);
    return true;
}

bool read_whole_gz(const std::string& path, std::string& out) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) return false;
    std::string buf;
    char tmp[1 << 16];
    int n;
    while ((n = gzread(g, tmp, sizeof tmp)) > 0) buf.append(tmp, (size_t)n);
    bool ok = (n == 0);
    gzclose(g);
    if (!ok) return false;
    out.swap(buf);
    return true;
}

// Length in bytes of the Unicode White_Space char starting at s[i] (char::is_whitespace), 0 if none.
size_t unicode_ws_len(const std::string& s, size_t i) {
    unsigned char c = (unsigned char)s[i];
    if (c == ' ' || (c >= 0x09 && c <= 0x0D)) return 1;
    if (c == 0xC2 && i + 1 < s.size()) {
        unsigned char d = (unsigned char)s[i + 1];
        if (d == 0x85 || d == 0xA0) return 2;
    }
    if (i + 2 < s.size()) {
        unsigned char d = (unsigned char)s[i + 1], e = (unsigned char)s[i + 2];
        if (c == 0xE1 && d == 0x9A && e == 0x80) return 3;                       // U+1680
        if (c == 0xE2 && d == 0x80 && ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF)) return 3;
        if (c == 0xE2 && d == 0x81 && e == 0x9F) return 3;                       // U+205F
        if (c == 0xE3 && d == 0x80 && e == 0x80) return 3;                       // U+3000
    }
    return 0;
}

struct FastaRec { std::string name, description, sequence; };

// misc.rs:102-167 load_fasta_not_gzipped / load_fasta_gzipped (same body over a different reader)
std::vector<FastaRec> load_fasta_text(const std::string& data, const std::string& filename) {
    std::vector<FastaRec> recs;
    std::string name, description, sequence;
    for_each_line(data, [&](std::string text) {
        if (text.size() == 0) return;
        if (text[0] == '>') {
            if (name.size() > 0) {
                make_ascii_uppercase(sequence);
                recs.push_back({name, description, sequence});
                sequence.clear();
            }
            // text[1..].splitn(2, char::is_whitespace)
            std::string rest = text.substr(1);
            size_t i = 0;
            size_t wl = 0;
            while (i < rest.size() && (wl = unicode_ws_len(rest, i)) == 0) i++;
            if (i < rest.size()) { name = rest.substr(0, i); description = rest.substr(i + wl); }
            else { name = rest; description.clear(); }
        } else {
            if (name.size() == 0) quit_with_error("\"" + filename + "\" is not correctly formatted");
            sequence += text;
        }
    });
    if (name.size() > 0) {
        make_ascii_uppercase(sequence);
        recs.push_back({name, description, sequence});
    }
    return recs;
}

// misc.rs:38-99 load_fasta + is_file_gzipped + check_load_fasta
std::vector<FastaRec> load_fasta(const std::string& filename) {
    FILE* f = fopen(filename.c_str(), "rb");
    if (!f) quit_with_error("unable to open \"" + filename + "\"");
    unsigned char b[2];
    size_t got = fread(b, 1, 2, f);
    fclose(f);
    if (got != 2) quit_with_error("\"" + filename + "\" is too small");
    bool gz = (b[0] == 31 && b[1] == 139);
    std::string data;
    bool ok = gz ? read_whole_gz(filename, data) : read_whole_file(filename, data);
    if (!ok) quit_with_error("unable to load \"" + filename + "\"");
    std::vector<FastaRec> recs = load_fasta_text(data, filename);
    if (recs.size() == 0) quit_with_error("\"" + filename + "\" contains no sequences");
    for (auto& r : recs) {
        if (r.name.size() == 0) quit_with_error("\"" + filename + "\" has an unnamed sequence");
        if (r.sequence.size() == 0) quit_with_error("\"" + filename + "\" has an empty sequence");
    }
    std::unordered_set<std::string> set;
    for (auto& r : recs) set.insert(r.name);
    if (set.size() < recs.size()) quit_with_error("\"" + filename + "\" has a duplicated name");
    return recs;
}

// ---------------------------------------------------------------------------------------------
// alignment.rs
// ---------------------------------------------------------------------------------------------

inline bool is_cigar_op(char c) {
    return c == 'M' || c == 'I' || c == 'D' || c == 'N' || c == 'S' || c == 'H' || c == 'P' ||
           c == '=' || c == 'X';
}

// Leftmost non-overlapping matches of the regex \d+[MIDNSHP=X] (alignment.rs:27-29), as
// regex::find_iter yields them (ASCII digits).  Each match = (start, end).
std::vector<std::pair<size_t, size_t>> cigar_regex_find_iter(const std::string& s) {
    std::vector<std::pair<size_t, size_t>> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        if (s[i] >= '0' && s[i] <= '9') {
            size_t j = i;
            while (j < n && s[j] >= '0' && s[j] <= '9') j++;
            if (j < n && is_cigar_op(s[j])) { out.push_back({i, j + 1}); i = j + 1; }
            else i = j;      // digits not followed by an op letter: no match can start inside them
        } else i++;
    }
    return out;
}

// Rust "123".parse::<u32/usize>(): optional '+', then >=1 ASCII digits, overflow is an error.
bool rust_parse_uint(const std::string& s, uint64_t maxv, uint64_t& out) {
    size_t i = 0;
    if (i < s.size() && s[i] == '+') i++;
    if (i >= s.size()) return false;
    uint64_t v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        uint64_t d = (uint64_t)(s[i] - '0');
        if (v > (maxv - d) / 10) return false;
        v = v * 10 + d;
    }
    out = v;
    return true;
}

// alignment.rs:325-346 get_expanded_cigar
bool get_expanded_cigar(const std::string& cigar, size_t read_seq_len, std::string& expanded) {
    expanded.clear();
    if (cigar == "*") return true;
    expanded.reserve(read_seq_len);
    size_t total_len = 0;
    for (auto m : cigar_regex_find_iter(cigar)) {
        uint64_t num;
        if (!rust_parse_uint(cigar.substr(m.first, m.second - 1 - m.first), 0xFFFFFFFFull, num))
            rust_panic("CIGAR length does not fit u32");
        char letter = cigar[m.second - 1];
        for (uint64_t k = 0; k < num; ++k) expanded.push_back(letter);
        total_len += m.second - m.first;
    }
    if (cigar.size() != total_len) return false;
    return true;
}

bool eq_ignore_ascii_case(const std::string& a, const char* b) {
    size_t n = strlen(b);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i) {
        char x = a[i], y = b[i];
        if (x >= 'A' && x <= 'Z') x = (char)(x + 32);
        if (y >= 'A' && y <= 'Z') y = (char)(y + 32);
        if (x != y) return false;
    }
    return true;
}

std::vector<std::string> split_tabs(const std::string& line) {
    std::vector<std::string> parts;
    size_t pos = 0;
    for (;;) {
        size_t t = line.find('\t', pos);
        if (t == std::string::npos) { parts.emplace_back(line, pos); break; }
        parts.emplace_back(line, pos, t - pos);
        pos = t + 1;
    }
    return parts;
}

// alignment.rs:33-43
struct Alignment {
    std::string read_name, ref_name;
    uint32_t sam_flags = 0;
    size_t ref_start = 0;
    std::string cigar, expanded_cigar, read_seq;
    uint32_t mismatches = 0;
    bool pass_qc = true;

    bool is_aligned() const { return (sam_flags & 4) == 0; }               // alignment.rs:130-132
    bool is_on_forward_strand() const { return (sam_flags & 16) == 0; }    // alignment.rs:151-153
    int get_strand() const { return is_on_forward_strand() ? 1 : -1; }     // alignment.rs:134-136

    // alignment.rs:138-149
    size_t get_ref_end() const {
        size_t ref_end = ref_start;
        for (auto m : cigar_regex_find_iter(cigar)) {
            uint64_t num;
            if (!rust_parse_uint(cigar.substr(m.first, m.second - 1 - m.first), ~0ull, num))
                rust_panic("CIGAR length does not fit usize");
            char letter = cigar[m.second - 1];
            if (letter == 'M' || letter == 'D' || letter == 'N' || letter == '=' || letter == 'X')
                ref_end += num;
        }
        return ref_end;
    }

    // alignment.rs:155-159
    bool starts_and_ends_with_match() const {
        if (expanded_cigar.empty()) rust_panic("empty expanded CIGAR for an aligned record (read " + read_name + ")");
        char f = expanded_cigar.front(), l = expanded_cigar.back();
        return (f == 'M' || f == '=') && (l == 'M' || l == '=');
    }

    // alignment.rs:161-167
    void add_read_seq(const std::string& seq, int strand) {
        if (get_strand() == strand) read_seq = seq;
        else read_seq = reverse_complement(seq);
    }
};

// alignment.rs:49-98 Alignment::new.  Returns "" on success, else the Err string.
const char* alignment_new(const std::string& sam_line, Alignment& a) {
    std::vector<std::string> parts = split_tabs(sam_line);
    if (parts.size() < 11) return "too few columns";
    uint64_t v;
    if (!rust_parse_uint(parts[1], 0xFFFFFFFFull, v)) rust_panic("invalid FLAG field \"" + parts[1] + "\"");
    uint32_t sam_flags = (uint32_t)v;
    if (!rust_parse_uint(parts[3], ~0ull, v)) rust_panic("invalid POS field \"" + parts[3] + "\"");
    size_t ref_start = (size_t)v;
    if (ref_start > 0) ref_start -= 1;
    const std::string& cigar = parts[5];
    const std::string& read_seq = parts[9];

    uint32_t mismatches = 0xFFFFFFFFu;
    bool pass_qc = true;
    for (size_t i = 11; i < parts.size(); ++i) {
        const std::string& p = parts[i];
        if (p.compare(0, 5, "NM:i:") == 0) {
            uint64_t nm;
            if (!rust_parse_uint(p.substr(5), 0xFFFFFFFFull, nm)) rust_panic("invalid NM tag \"" + p + "\"");
            mismatches = (uint32_t)nm;
        }
        if (eq_ignore_ascii_case(p, "ZP:Z:fail")) pass_qc = false;
    }
    if (mismatches == 0xFFFFFFFFu && (sam_flags & 4) == 0) return "missing NM tag";
    std::string expanded;
    if (!get_expanded_cigar(cigar, read_seq.size(), expanded))
        quit_with_error("encountered an invalid CIGAR string for read " + parts[0] + ": \"" + cigar + "\"");

    a.read_name = parts[0];
    a.ref_name = parts[2];
    a.sam_flags = sam_flags;
    a.ref_start = ref_start;
    a.cigar = cigar;
    a.expanded_cigar = std::move(expanded);
    a.read_seq = read_seq;
    make_ascii_uppercase(a.read_seq);
    a.mismatches = mismatches;
    a.pass_qc = pass_qc;
    return "";
}

// alignment.rs:102-128 Alignment::new_quick
const char* alignment_new_quick(const std::string& sam_line, Alignment& a) {
    std::vector<std::string> parts = split_tabs(sam_line);
    if (parts.size() < 11) return "too few columns";
    uint64_t v;
    if (!rust_parse_uint(parts[1], 0xFFFFFFFFull, v)) rust_panic("invalid FLAG field \"" + parts[1] + "\"");
    uint32_t sam_flags = (uint32_t)v;
    if (!rust_parse_uint(parts[3], ~0ull, v)) rust_panic("invalid POS field \"" + parts[3] + "\"");
    size_t ref_start = (size_t)v;
    if (ref_start > 0) ref_start -= 1;
    a.read_name = parts[0];
    a.ref_name = parts[2];
    a.sam_flags = sam_flags;
    a.ref_start = ref_start;
    a.cigar = parts[5];
    a.expanded_cigar.clear();
    a.read_seq.clear();
    a.mismatches = 0;
    a.pass_qc = true;
    return "";
}

// alignment.rs:364-378 trim_bases_for_homopolymers
void trim_bases_for_homopolymers(std::vector<std::pair<size_t, size_t>>& read_bases, const std::string& read_seq) {
    if (read_bases.empty()) rust_panic("trim on an empty read_bases vector");
    auto last = read_bases.back();
    std::string last_base = read_seq.substr(last.first, last.second - last.first);
    while (read_bases.size() > 0) {
        auto cur = read_bases.back();
        if (read_seq.compare(cur.first, cur.second - cur.first, last_base) != 0) break;
        read_bases.pop_back();
    }
    if (read_bases.size() > 0) read_bases.pop_back();
}

// alignment.rs:175-201 get_read_bases_for_each_target_base
std::vector<std::pair<size_t, size_t>> get_read_bases_for_each_target_base(const Alignment& a) {
    size_t i = 0;
    std::vector<std::pair<size_t, size_t>> read_bases;
    read_bases.reserve(a.expanded_cigar.size());
    for (char c : a.expanded_cigar) {
        if (c == 'M' || c == '=' || c == 'X') {
            read_bases.push_back({i, i + 1});
            i += 1;
        } else if (c == 'I') {
            if (read_bases.empty()) rust_panic("insertion before any reference-consuming CIGAR op");
            read_bases.back().second = i + 1;
            i += 1;
        } else if (c == 'D') {
            read_bases.push_back({i, i});
        } else {
            quit_with_error("unexpected character (other than M, =, X, I or D) in CIGAR string for read " +
                            a.read_name + ": \"" + a.cigar + "\" - did you use BWA MEM to generate your alignments?");
        }
    }
    if (i != a.read_seq.size())
        quit_with_error("CIGAR string for read " + a.read_name + " does not match read sequence");
    trim_bases_for_homopolymers(read_bases, a.read_seq);
    return read_bases;
}

// ---------------------------------------------------------------------------------------------
// pileup.rs
// ---------------------------------------------------------------------------------------------

enum BaseStatus { DepthTooLow, NoValidOptions, MultipleValidOptions, TooClose, OriginalBaseKept, Changed };

const char* status_str(BaseStatus s) {     // pileup.rs:156-163
    switch (s) {
        case OriginalBaseKept: return "kept";
        case Changed: return "changed";
        case DepthTooLow: return "low_depth";
        case NoValidOptions: return "none";
        case MultipleValidOptions: return "multiple";
        case TooClose: return "too_close";
    }
    return "?";
}

// pileup.rs:29-41
struct PileupBase {
    char original;
    double depth = 0.0;
    uint32_t count_a = 0, count_c = 0, count_g = 0, count_t = 0;
    std::unordered_map<std::string, uint32_t> counts;
    explicit PileupBase(char o) : original(o) {}

    // pileup.rs:56-65
    void add_seq(const std::string& seq, double depth_contribution) {
        if (seq == "A") count_a += 1;
        else if (seq == "C") count_c += 1;
        else if (seq == "G") count_g += 1;
        else if (seq == "T") count_t += 1;
        else counts[seq] += 1;
        depth += depth_contribution;
    }

    // pileup.rs:137-148
    std::string get_count_str() const {
        std::vector<std::string> v;
        if (count_a > 0) v.push_back("Ax" + std::to_string(count_a));
        if (count_c > 0) v.push_back("Cx" + std::to_string(count_c));
        if (count_g > 0) v.push_back("Gx" + std::to_string(count_g));
        if (count_t > 0) v.push_back("Tx" + std::to_string(count_t));
        for (auto& kv : counts) v.push_back(kv.first + "x" + std::to_string(kv.second));
        std::sort(v.begin(), v.end());
        std::string out;
        for (size_t i = 0; i < v.size(); ++i) { if (i) out += ","; out += v[i]; }
        return out;
    }

    // pileup.rs:67-134 (+ :150-166 debug line)
    void get_polished_seq(uint32_t min_depth, double fraction_valid, double fraction_invalid,
                          bool build_debug_line, std::string& new_base, BaseStatus& status,
                          std::string& debug_line) const {
        std::string original(1, this->original);
        uint32_t valid_threshold = std::max(min_depth, bankers_rounding(depth * fraction_valid));
        uint32_t invalid_threshold = bankers_rounding(depth * fraction_invalid);

        std::vector<std::string> valid_seqs, intermediate_seqs;
        if (count_a >= valid_threshold) valid_seqs.push_back("A");
        else if (count_a >= invalid_threshold) intermediate_seqs.push_back("A");
        if (count_c >= valid_threshold) valid_seqs.push_back("C");
        else if (count_c >= invalid_threshold) intermediate_seqs.push_back("C");
        if (count_g >= valid_threshold) valid_seqs.push_back("G");
        else if (count_g >= invalid_threshold) intermediate_seqs.push_back("G");
        if (count_t >= valid_threshold) valid_seqs.push_back("T");
        else if (count_t >= invalid_threshold) intermediate_seqs.push_back("T");
        for (auto& kv : counts) {
            if (kv.second >= valid_threshold) valid_seqs.push_back(kv.first);
            else if (kv.second >= invalid_threshold) intermediate_seqs.push_back(kv.first);
        }

        new_base = original;
        status = OriginalBaseKept;
        if (depth < (double)min_depth) {
            status = DepthTooLow;
        } else if (valid_seqs.size() == 1) {
            if (intermediate_seqs.size() > 0) status = TooClose;
            else {
                new_base = valid_seqs[0];
                if (new_base != original) status = Changed;
            }
        } else if (valid_seqs.size() == 0) {
            status = NoValidOptions;
        } else {
            status = MultipleValidOptions;
        }

        debug_line.clear();
        if (build_debug_line) {
            char buf[64];
            snprintf(buf, sizeof buf, "%.1f", depth);      // Rust {:.1}: both round the exact binary value
            debug_line = original + "\t" + buf + "\t" + std::to_string(invalid_threshold) + "\t" +
                         std::to_string(valid_threshold) + "\t" + get_count_str() + "\t" +
                         status_str(status) + "\t" + new_base;
        }
    }
};

// pileup.rs:173-200
struct Pileup {
    std::vector<PileupBase> bases;
    explicit Pileup(const std::string& seq) {
        bases.reserve(seq.size());
        for (char b : seq) bases.emplace_back(b);
    }
    void add_alignment(const Alignment& a, double depth_contribution) {
        auto read_bases = get_read_bases_for_each_target_base(a);
        size_t i = a.ref_start;
        for (auto se : read_bases) {
            if (i >= bases.size()) rust_panic("alignment of read " + a.read_name + " extends past the end of " + a.ref_name);
            if (se.first == se.second) bases[i].add_seq("-", depth_contribution);
            else bases[i].add_seq(a.read_seq.substr(se.first, se.second - se.first), depth_contribution);
            i += 1;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// alignment.rs: process_sam path
// ---------------------------------------------------------------------------------------------

struct PhaseTimes { double load = 0, parse = 0, scatter = 0, vote = 0; };

// alignment.rs:311-322
void get_read_seq_from_alignments(const std::vector<Alignment>& alignments, std::string& seq, int& strand) {
    for (auto& a : alignments) {
        if (a.read_seq == "*") continue;
        seq = a.read_seq; strand = a.get_strand();
        return;
    }
    if (alignments.empty()) rust_panic("process_one_read on an empty group (SAM file with no aligned records)");
    quit_with_error("no alignments for read " + alignments.front().read_name + " contain sequence");
}

// alignment.rs:275-305
size_t process_one_read(std::vector<Alignment>& alignments, std::unordered_map<std::string, Pileup>& pileups,
                        uint32_t max_errors, bool careful) {
    if (careful && alignments.size() > 1) return 0;
    std::string read_seq; int strand = 1;
    get_read_seq_from_alignments(alignments, read_seq, strand);

    std::vector<Alignment> good;
    for (auto& a : alignments)
        if (a.starts_and_ends_with_match() && a.mismatches <= max_errors && a.pass_qc) good.push_back(std::move(a));
    double depth_contribution = 1.0 / (double)good.size();

    for (auto& a : good) {
        bool needs_length = (a.read_seq == "*");
        if (needs_length) a.add_read_seq(read_seq, strand);
    }
    for (auto& a : good) {
        auto it = pileups.find(a.ref_name);
        if (it == pileups.end()) quit_with_error("query name " + a.ref_name + " in SAM but not in assembly");
        it->second.add_alignment(a, depth_contribution);
    }
    return good.size();
}

// alignment.rs:225-272 add_to_pileup (+ :214-222 process_sam)
void process_sam(const std::string& filename, std::unordered_map<std::string, Pileup>& pileups,
                 uint32_t max_errors, bool careful, size_t& alignment_count, size_t& used_count,
                 size_t& read_count, PhaseTimes& pt) {
    auto t_load = Clock::now();
    std::string data;
    if (!read_whole_file(filename, data)) quit_with_error("unable to load alignments from \"" + filename + "\"");
    pt.load += secs_since(t_load);

    std::string current_read_name;
    std::vector<Alignment> current;
    size_t line_count = 0;
    alignment_count = used_count = read_count = 0;
    double scatter_secs = 0;
    auto t_all = Clock::now();

    for_each_line(data, [&](std::string sam_line) {
        line_count += 1;
        if (sam_line.size() == 0) return;
        if (sam_line[0] == '@') return;
        Alignment a;
        const char* err = alignment_new(sam_line, a);
        if (err[0]) quit_with_error(std::string(err) + " in \"" + filename + "\" (line " + std::to_string(line_count) + ")");
        if (!a.is_aligned()) return;
        alignment_count += 1;
        std::string read_name = a.read_name;
        if (current_read_name.empty() || current_read_name == a.read_name) {
            current.push_back(std::move(a));
        } else {
            auto t0 = Clock::now();
            used_count += process_one_read(current, pileups, max_errors, careful);
            scatter_secs += secs_since(t0);
            read_count += 1;
            current.clear();
            current.push_back(std::move(a));
        }
        current_read_name = read_name;
    });
    if (alignment_count == 0) {
        // The reference panics inside get_read_seq_from_alignments before reaching this message
        // (alignment.rs:265,319 vs :268-270); behaviour on such input is not part of parity.
        quit_with_error("no alignments in \"" + filename + "\"");
    }
    auto t0 = Clock::now();
    used_count += process_one_read(current, pileups, max_errors, careful);
    scatter_secs += secs_since(t0);
    read_count += 1;
    pt.scatter += scatter_secs;
    pt.parse += secs_since(t_all) - scatter_secs;
}

// ---------------------------------------------------------------------------------------------
// polish.rs
// ---------------------------------------------------------------------------------------------

struct PolishStats { std::string name; size_t length, new_length, changed, zero_depth; double total_depth; };

struct PolishOutput {
    std::string fasta;       // exactly what the reference prints to stdout
    std::string debug_tsv;   // exactly what --debug writes (only when requested)
    std::vector<PolishStats> stats;
    size_t alignment_total = 0, used_total = 0;
    PhaseTimes pt;
};

// polish.rs:277-287
void check_option_values(double fraction_invalid, double fraction_valid) {
    if (fraction_valid <= 0.0 || fraction_valid >= 1.0) quit_with_error("--fraction_valid must be between 0 and 1 (exclusive)");
    if (fraction_invalid <= 0.0 || fraction_invalid >= 1.0) quit_with_error("--fraction_invalid must be between 0 and 1 (exclusive)");
    if (fraction_invalid >= fraction_valid) quit_with_error("--fraction_invalid must be less than --fraction_valid");
}

bool file_exists(const std::string& p) { FILE* f = fopen(p.c_str(), "rb"); if (!f) return false; fclose(f); return true; }

// polish.rs:26-38 polish (+ :93-203)
void polish(bool debug, double fraction_invalid, double fraction_valid, uint32_t max_errors, uint32_t min_depth,
            bool careful, const std::string& assembly, const std::vector<std::string>& sam, PolishOutput& out) {
    check_option_values(fraction_invalid, fraction_valid);
    if (!file_exists(assembly)) quit_with_error("\"" + assembly + "\" file does not exist");
    for (auto& s : sam) if (!file_exists(s)) quit_with_error("\"" + s + "\" file does not exist");

    // load_assembly polish.rs:93-106
    auto t0 = Clock::now();
    std::vector<FastaRec> fasta = load_fasta(assembly);
    std::vector<std::pair<std::string, std::string>> seq_names;
    std::unordered_map<std::string, Pileup> pileups;
    for (auto& r : fasta) {
        seq_names.push_back({r.name, r.description});
        pileups.emplace(r.name, Pileup(r.sequence));
    }
    out.pt.load += secs_since(t0);

    // load_alignments polish.rs:109-134
    for (auto& s : sam) {
        size_t ac, uc, rc;
        process_sam(s, pileups, max_errors, careful, ac, uc, rc, out.pt);
        out.alignment_total += ac;
        out.used_total += uc;
    }

    // polish_sequences / polish_one_sequence polish.rs:137-193
    t0 = Clock::now();
    if (debug) out.debug_tsv = "name\tpos\tbase\tdepth\tinvalid\tvalid\tpileup\tstatus\tnew_base\n";
    for (auto& nd : seq_names) {
        const Pileup& pileup = pileups.at(nd.first);
        size_t seq_len = pileup.bases.size();
        std::string polished_seq;
        polished_seq.reserve(seq_len);
        double total_depth = 0.0;
        size_t zero_depth_count = 0, changed_count = 0, pos = 0;
        std::string seq, debug_line;
        BaseStatus status;
        for (auto& b : pileup.bases) {
            b.get_polished_seq(min_depth, fraction_valid, fraction_invalid, debug, seq, status, debug_line);
            if (status == Changed) changed_count += 1;
            total_depth += b.depth;
            if (b.depth == 0.0) zero_depth_count += 1;
            if (debug) out.debug_tsv += nd.first + "\t" + std::to_string(pos) + "\t" + debug_line + "\n";
            polished_seq += seq;
            pos += 1;
        }
        // polished_seq.replace("-", "")  polish.rs:188
        polished_seq.erase(std::remove(polished_seq.begin(), polished_seq.end(), '-'), polished_seq.end());
        // print_seq_to_stdout polish.rs:196-203
        out.fasta += ">" + nd.first;
        if (nd.second.size() > 0) out.fasta += " " + nd.second;
        out.fasta += " polypolish\n";
        out.fasta += polished_seq;
        out.fasta += "\n";
        out.stats.push_back({nd.first, seq_len, polished_seq.size(), changed_count, zero_depth_count, total_depth});
    }
    out.pt.vote += secs_since(t0);
}

// ---------------------------------------------------------------------------------------------
// filter.rs
// ---------------------------------------------------------------------------------------------

// filter.rs:189-209
std::string get_orientation(const Alignment& a1, const Alignment& a2) {
    char s1 = a1.is_on_forward_strand() ? 'f' : 'r';
    char s2 = a2.is_on_forward_strand() ? 'f' : 'r';
    size_t p1 = a1.is_on_forward_strand() ? a1.ref_start : a1.get_ref_end();
    size_t p2 = a2.is_on_forward_strand() ? a2.ref_start : a2.get_ref_end();
    if (s1 != s2) {
        if (p1 < p2) return std::string() + s1 + s2;
        return std::string() + s2 + s1;
    }
    if (s1 == 'f') return (p1 < p2) ? "ff" : "rr";
    return (p2 < p1) ? "ff" : "rr";
}

// filter.rs:212-218
uint32_t get_insert_size(const Alignment& a1, const Alignment& a2) {
    size_t p[4] = {a1.ref_start, a1.get_ref_end(), a2.ref_start, a2.get_ref_end()};
    size_t lo = *std::min_element(p, p + 4), hi = *std::max_element(p, p + 4);
    return (uint32_t)(hi - lo);
}

// filter.rs:249-259
uint32_t get_percentile(const std::vector<uint32_t>& sorted_list, double percentile) {
    if (sorted_list.empty()) return 0;
    double fraction = percentile / 100.0;
    double r = std::ceil(fraction * (double)sorted_list.size());
    size_t rank;                                   // `as usize` saturates
    if (!(r == r) || r <= 0.0) rank = 0;
    else if (r >= 18446744073709551615.0) rank = (size_t)-1;
    else rank = (size_t)r;
    if (rank < 1) rank = 1;
    if (rank - 1 < sorted_list.size()) return sorted_list[rank - 1];
    return 0;
}

using InsertSizes = std::unordered_map<std::string, std::vector<uint32_t>>;

// filter.rs:238-246
std::string auto_determine_orientation(const InsertSizes& insert_sizes) {
    size_t max_count = 0;
    for (auto& kv : insert_sizes) max_count = std::max(max_count, kv.second.size());
    std::vector<std::string> orientations;
    for (const char* o : {"fr", "rf", "ff", "rr"}) {
        auto it = insert_sizes.find(o);
        size_t c = (it == insert_sizes.end()) ? 0 : it->second.size();
        if (c == max_count) orientations.push_back(o);
    }
    if (orientations.size() != 1) quit_with_error("could not automatically determine read pair orientation");
    return orientations[0];
}

using AlignmentMap = std::unordered_map<std::string, std::vector<Alignment>>;

// filter.rs:110-145
void load_alignments_one_file(const std::string& sam_filename, AlignmentMap& alignments, const char* suffix) {
    std::string data;
    if (!read_whole_file(sam_filename, data)) quit_with_error("unable to load alignments from \"" + sam_filename + "\"");
    size_t line_count = 0;
    for_each_line(data, [&](std::string sam_line) {
        line_count += 1;
        if (!sam_line.empty() && sam_line[0] == '@') return;
        Alignment a;
        const char* err = alignment_new_quick(sam_line, a);
        if (err[0]) quit_with_error(std::string(err) + " in \"" + sam_filename + "\" (line " + std::to_string(line_count) + ")");
        if (!a.is_aligned()) return;
        a.read_name += suffix;
        std::string key = a.read_name;
        alignments[key].push_back(std::move(a));
    });
    if (alignments.empty()) quit_with_error("no alignments found in \"" + sam_filename + "\"");
}

struct FilterOutput {
    std::string out1, out2;              // exactly what the reference writes to --out1 / --out2
    uint32_t low = 0, high = 0;
    std::string orientation;
    size_t pairs[4] = {0, 0, 0, 0};      // fr, rf, ff, rr
    size_t before_count = 0, after_count = 0;
};

// filter.rs:148-186
void get_insert_size_thresholds(const AlignmentMap& alignments, const std::string& correct_orientation_in,
                                double low_percentile, double high_percentile, FilterOutput& out) {
    InsertSizes insert_sizes;
    for (auto& kv : alignments) {
        const std::string& name_1 = kv.first;
        if (name_1.size() < 2 || name_1.compare(name_1.size() - 2, 2, "_1") != 0 || kv.second.size() != 1) continue;
        std::string name_2 = name_1.substr(0, name_1.size() - 2) + "_2";
        auto it = alignments.find(name_2);
        if (it != alignments.end()) {
            auto& al2 = it->second;
            if (al2.size() == 1 && kv.second[0].ref_name == al2[0].ref_name) {
                std::string o = get_orientation(kv.second[0], al2[0]);
                uint32_t ins = get_insert_size(kv.second[0], al2[0]);
                insert_sizes[o].push_back(ins);
            }
        }
    }
    if (insert_sizes.empty())
        quit_with_error("no one-alignment-per-read pairs available to determine orientation and insert size thresholds");
    const char* names[4] = {"fr", "rf", "ff", "rr"};
    for (int i = 0; i < 4; ++i) { auto it = insert_sizes.find(names[i]); out.pairs[i] = it == insert_sizes.end() ? 0 : it->second.size(); }
    // determine_correct_orientation filter.rs:221-235
    std::string correct = correct_orientation_in == "auto" ? auto_determine_orientation(insert_sizes) : correct_orientation_in;
    std::vector<uint32_t> sizes;
    auto it = insert_sizes.find(correct);
    if (it != insert_sizes.end()) sizes = it->second;
    if (sizes.empty()) quit_with_error("no read pairs available to determine insert size thresholds");
    std::sort(sizes.begin(), sizes.end());
    out.low = get_percentile(sizes, low_percentile);
    out.high = get_percentile(sizes, high_percentile);
    out.orientation = correct;
}

// filter.rs:352-377
bool alignment_pass_qc(const Alignment& a, const std::vector<Alignment>& this_alignments,
                       const std::vector<Alignment>& pair_alignments, uint32_t low, uint32_t high,
                       const std::string& correct_orientation) {
    if (pair_alignments.empty()) return true;
    if (this_alignments.size() == 1) return true;
    for (auto& pa : pair_alignments) {
        bool same_ref = a.ref_name == pa.ref_name;
        uint32_t insert = get_insert_size(a, pa);
        std::string orientation = get_orientation(a, pa);
        if (same_ref && low <= insert && insert <= high && orientation == correct_orientation) return true;
    }
    return false;
}

// filter.rs:296-349
size_t filter_sam(const std::string& in_filename, const AlignmentMap& alignments, uint32_t low, uint32_t high,
                  const std::string& correct_orientation, int read_num, std::string& out) {
    static const std::vector<Alignment> NO_ALIGNMENTS;
    std::string data;
    if (!read_whole_file(in_filename, data)) quit_with_error("unable to write alignments");
    size_t pass_count = 0;
    for_each_line(data, [&](std::string sam_line) {
        if (!sam_line.empty() && sam_line[0] == '@') { out += sam_line; out += "\n"; return; }
        Alignment a;
        const char* err = alignment_new_quick(sam_line, a);
        if (err[0]) rust_panic(std::string("called unwrap on Err(") + err + ")");
        if (!a.is_aligned()) { out += sam_line; out += "\n"; return; }
        std::string this_name = a.read_name + (read_num == 1 ? "_1" : "_2");
        std::string pair_name = a.read_name + (read_num == 1 ? "_2" : "_1");
        const std::vector<Alignment>& this_alignments = alignments.at(this_name);
        auto it = alignments.find(pair_name);
        const std::vector<Alignment>& pair_alignments = it == alignments.end() ? NO_ALIGNMENTS : it->second;
        if (alignment_pass_qc(a, this_alignments, pair_alignments, low, high, correct_orientation)) {
            out += sam_line; out += "\n";
            pass_count += 1;
        } else {
            out += sam_line; out += "\tZP:Z:fail\n";    // split('\t') + push + join('\t') == append
        }
    });
    return pass_count;
}

// filter.rs:26-53
void filter(const std::string& in1, const std::string& in2, const std::string& out1_name, const std::string& out2_name,
            const std::string& orientation, double low, double high, FilterOutput& out) {
    std::unordered_set<std::string> files;
    if (!files.insert(in1).second || !files.insert(in2).second || !files.insert(out1_name).second || !files.insert(out2_name).second)
        quit_with_error("--in1, --in2, --out1 and --out2 must all have unique values");
    if (low <= 0.0 || low >= 50.0) quit_with_error("--low must be greater than 0 and less than 50");
    if (high <= 50.0 || high >= 100.0) quit_with_error("--high must be greater than 50 and less than 100");
    AlignmentMap alignments;
    load_alignments_one_file(in1, alignments, "_1");
    load_alignments_one_file(in2, alignments, "_2");
    for (auto& kv : alignments) out.before_count += kv.second.size();
    get_insert_size_thresholds(alignments, orientation, low, high, out);
    out.after_count += filter_sam(in1, alignments, out.low, out.high, out.orientation, 1, out.out1);
    out.after_count += filter_sam(in2, alignments, out.low, out.high, out.orientation, 2, out.out2);
}

void set_err(char* err, size_t cap, const std::string& m) {
    if (!err || cap == 0) return;
    size_t n = std::min(cap - 1, m.size());
    memcpy(err, m.data(), n);
    err[n] = 0;
}

char* dup_buf(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    return p;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C entry points (ctypes from tests/ and bench.py only)
// ---------------------------------------------------------------------------------------------
extern "C" {

uint32_t orc_bankers_rounding(double x) { return bankers_rounding(x); }

// out must hold n+1 bytes
void orc_reverse_complement(const char* seq, char* out) {
    std::string r = reverse_complement(seq);
    memcpy(out, r.c_str(), r.size() + 1);
}

// returns 0 ok (out = malloc'd expanded string, free with orc_free), 1 = invalid
int orc_get_expanded_cigar(const char* cigar, uint64_t read_seq_len, char** out) {
    std::string e;
    try {
        if (!get_expanded_cigar(cigar, read_seq_len, e)) return 1;
    } catch (const OracleError&) { return 2; }
    *out = dup_buf(e);
    return 0;
}

// Alignment::new on one SAM line -> ref_start, ref_end. 0 ok, 1 Err, 2 quit/panic
int orc_alignment_new(const char* sam_line, uint64_t* ref_start, uint64_t* ref_end, uint32_t* nm, int* pass_qc,
                      char* err, uint64_t errcap) {
    try {
        Alignment a;
        const char* e = alignment_new(sam_line, a);
        if (e[0]) { set_err(err, errcap, e); return 1; }
        *ref_start = a.ref_start; *ref_end = a.get_ref_end(); *nm = a.mismatches; *pass_qc = a.pass_qc;
        return 0;
    } catch (const OracleError& e) { set_err(err, errcap, e.what()); return 2; }
}

// get_orientation on two SAM lines via new_quick -> 2-char string in out[3]
int orc_get_orientation(const char* line1, const char* line2, char* out, uint32_t* insert) {
    try {
        Alignment a, b;
        if (alignment_new_quick(line1, a)[0] || alignment_new_quick(line2, b)[0]) return 1;
        std::string o = get_orientation(a, b);
        out[0] = o[0]; out[1] = o[1]; out[2] = 0;
        *insert = get_insert_size(a, b);
        return 0;
    } catch (const OracleError&) { return 2; }
}

// counts[4] = sizes of fr, rf, ff, rr lists. returns index 0..3, or -1 = could not determine
int orc_auto_determine_orientation(const uint64_t* counts) {
    InsertSizes m;
    const char* names[4] = {"fr", "rf", "ff", "rr"};
    for (int i = 0; i < 4; ++i) if (counts[i]) m[names[i]] = std::vector<uint32_t>(counts[i], 1);
    try {
        std::string o = auto_determine_orientation(m);
        for (int i = 0; i < 4; ++i) if (o == names[i]) return i;
    } catch (const OracleError&) {}
    return -1;
}

uint32_t orc_get_percentile(const uint32_t* sorted, uint64_t n, double p) {
    return get_percentile(std::vector<uint32_t>(sorted, sorted + n), p);
}

// One PileupBase: seqs = n NUL-terminated strings back to back, contribs[n].
// status_out: 0 low_depth 1 none 2 multiple 3 too_close 4 kept 5 changed
int orc_vote(char original, const char* seqs, const double* contribs, uint64_t n, uint32_t min_depth,
             double fraction_valid, double fraction_invalid, char** new_base, int* status_out,
             char** count_str, char** debug_line) {
    PileupBase b(original);
    const char* p = seqs;
    for (uint64_t i = 0; i < n; ++i) { std::string s(p); p += s.size() + 1; b.add_seq(s, contribs[i]); }
    std::string nb, dl; BaseStatus st;
    b.get_polished_seq(min_depth, fraction_valid, fraction_invalid, true, nb, st, dl);
    *new_base = dup_buf(nb); *status_out = (int)st; *count_str = dup_buf(b.get_count_str()); *debug_line = dup_buf(dl);
    return 0;
}

// load_fasta -> "name\tdescription\tsequence\n" per record (test helper)
int orc_load_fasta(const char* path, char** out, char* err, uint64_t errcap) {
    try {
        std::string s;
        for (auto& r : load_fasta(path)) s += r.name + "\t" + r.description + "\t" + r.sequence + "\n";
        *out = dup_buf(s);
        return 0;
    } catch (const OracleError& e) { set_err(err, errcap, e.what()); return e.code; }
}

// Walk + trim for one good alignment: kept entries as (start,end) pairs. Test helper for the
// CIGAR-walk / homopolymer-trim semantics (alignment.rs:175-201, :364-378).
int orc_walk(const char* cigar, const char* read_seq, uint64_t* pairs, uint64_t cap, uint64_t* n_out,
             char* err, uint64_t errcap) {
    try {
        Alignment a; a.read_name = "r"; a.cigar = cigar; a.read_seq = read_seq;
        if (!get_expanded_cigar(a.cigar, a.read_seq.size(), a.expanded_cigar)) { set_err(err, errcap, "invalid CIGAR"); return 1; }
        auto v = get_read_bases_for_each_target_base(a);
        *n_out = v.size();
        for (size_t i = 0; i < v.size() && i < cap; ++i) { pairs[2 * i] = v[i].first; pairs[2 * i + 1] = v[i].second; }
        return 0;
    } catch (const OracleError& e) { set_err(err, errcap, e.what()); return e.code; }
}

struct orc_polish_result {
    char* fasta; uint64_t fasta_len;
    char* debug_tsv; uint64_t debug_len;
    uint64_t alignment_total, used_total;
    uint64_t n_contigs;
    uint64_t* changed; uint64_t* zero_depth; uint64_t* new_length; double* total_depth;
    double secs_load, secs_parse, secs_scatter, secs_vote;
};

int orc_polish(const char* assembly, const char* const* sams, int n_sams, double fraction_invalid,
               double fraction_valid, uint32_t max_errors, uint32_t min_depth, int careful, int debug,
               orc_polish_result* res, char* err, uint64_t errcap) {
    memset(res, 0, sizeof *res);
    try {
        std::vector<std::string> sam;
        for (int i = 0; i < n_sams; ++i) sam.push_back(sams[i]);
        PolishOutput out;
        polish(debug != 0, fraction_invalid, fraction_valid, max_errors, min_depth, careful != 0, assembly, sam, out);
        res->fasta = dup_buf(out.fasta); res->fasta_len = out.fasta.size();
        res->debug_tsv = dup_buf(out.debug_tsv); res->debug_len = out.debug_tsv.size();
        res->alignment_total = out.alignment_total; res->used_total = out.used_total;
        res->n_contigs = out.stats.size();
        res->changed = (uint64_t*)malloc(8 * out.stats.size() + 8);
        res->zero_depth = (uint64_t*)malloc(8 * out.stats.size() + 8);
        res->new_length = (uint64_t*)malloc(8 * out.stats.size() + 8);
        res->total_depth = (double*)malloc(8 * out.stats.size() + 8);
        for (size_t i = 0; i < out.stats.size(); ++i) {
            res->changed[i] = out.stats[i].changed; res->zero_depth[i] = out.stats[i].zero_depth;
            res->new_length[i] = out.stats[i].new_length; res->total_depth[i] = out.stats[i].total_depth;
        }
        res->secs_load = out.pt.load; res->secs_parse = out.pt.parse; res->secs_scatter = out.pt.scatter; res->secs_vote = out.pt.vote;
        return 0;
    } catch (const OracleError& e) { set_err(err, errcap, e.what()); return e.code; }
}

void orc_polish_result_free(orc_polish_result* r) {
    free(r->fasta); free(r->debug_tsv); free(r->changed); free(r->zero_depth); free(r->new_length); free(r->total_depth);
    memset(r, 0, sizeof *r);
}

struct orc_filter_result {
    char* out1; uint64_t out1_len; char* out2; uint64_t out2_len;
    uint32_t low, high; int orientation; uint64_t pairs[4]; uint64_t before_count, after_count;
};

int orc_filter(const char* in1, const char* in2, const char* out1_name, const char* out2_name,
               const char* orientation, double low, double high, orc_filter_result* res, char* err, uint64_t errcap) {
    memset(res, 0, sizeof *res);
    try {
        FilterOutput out;
        filter(in1, in2, out1_name, out2_name, orientation, low, high, out);
        res->out1 = dup_buf(out.out1); res->out1_len = out.out1.size();
        res->out2 = dup_buf(out.out2); res->out2_len = out.out2.size();
        res->low = out.low; res->high = out.high;
        const char* names[4] = {"fr", "rf", "ff", "rr"};
        res->orientation = -1;
        for (int i = 0; i < 4; ++i) { if (out.orientation == names[i]) res->orientation = i; res->pairs[i] = out.pairs[i]; }
        res->before_count = out.before_count; res->after_count = out.after_count;
        return 0;
    } catch (const OracleError& e) { set_err(err, errcap, e.what()); return e.code; }
}

void orc_filter_result_free(orc_filter_result* r) { free(r->out1); free(r->out2); memset(r, 0, sizeof *r); }

void orc_free(void* p) { free(p); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// CLI (polypolish_oracle polish|filter ...), same flags as main.rs:44-109.  Used by the tests to diff
// whole-command output against the GPU CLI and by bench.py --impl reference.
// ---------------------------------------------------------------------------------------------
#ifdef ORACLE_MAIN
static double parse_double_arg(const char* s) { return strtod(s, nullptr); }

int main(int argc, char** argv) {
    try {
        if (argc < 2) { fprintf(stderr, "usage: polypolish_oracle polish|filter ...\n"); return 2; }
        std::string cmd = argv[1];
        if (cmd == "polish") {
            std::string debug_path; double fi = 0.2, fv = 0.5; uint32_t me = 10, md = 5; bool careful = false;
            std::vector<std::string> pos;
            for (int i = 2; i < argc; ++i) {
                std::string a = argv[i];
                auto need = [&](const char* n) { if (i + 1 >= argc) quit_with_error(std::string("missing value for ") + n); return argv[++i]; };
                if (a == "--debug") debug_path = need("--debug");
                else if (a == "-i" || a == "--fraction_invalid") fi = parse_double_arg(need("-i"));
                else if (a == "-v" || a == "--fraction_valid") fv = parse_double_arg(need("-v"));
                else if (a == "-m" || a == "--max_errors") me = (uint32_t)strtoul(need("-m"), nullptr, 10);
                else if (a == "-d" || a == "--min_depth") md = (uint32_t)strtoul(need("-d"), nullptr, 10);
                else if (a == "--careful") careful = true;
                else pos.push_back(a);
            }
            if (pos.empty()) quit_with_error("missing <ASSEMBLY>");
            std::vector<std::string> sam(pos.begin() + 1, pos.end());
            PolishOutput out;
            polish(!debug_path.empty(), fi, fv, me, md, careful, pos[0], sam, out);
            fwrite(out.fasta.data(), 1, out.fasta.size(), stdout);
            if (!debug_path.empty()) { FILE* f = fopen(debug_path.c_str(), "wb"); if (!f) quit_with_error("unable to create \"" + debug_path + "\""); fwrite(out.debug_tsv.data(), 1, out.debug_tsv.size(), f); fclose(f); }
            fprintf(stderr, "oracle phases (s): load %.3f parse %.3f walk+scatter %.3f vote+print %.3f; alignments %zu used %zu\n",
                    out.pt.load, out.pt.parse, out.pt.scatter, out.pt.vote, out.alignment_total, out.used_total);
            for (auto& s : out.stats) fprintf(stderr, "%s: %zu bp -> %zu bp, %zu changed, %zu zero-depth, mean depth %.1f\n", s.name.c_str(), s.length, s.new_length, s.changed, s.zero_depth, s.total_depth / (double)s.length);
            return 0;
        } else if (cmd == "filter") {
            std::string in1, in2, out1, out2, orientation = "auto"; double low = 0.1, high = 99.9;
            for (int i = 2; i < argc; ++i) {
                std::string a = argv[i];
                auto need = [&](const char* n) { if (i + 1 >= argc) quit_with_error(std::string("missing value for ") + n); return argv[++i]; };
                if (a == "--in1") in1 = need("--in1"); else if (a == "--in2") in2 = need("--in2");
                else if (a == "--out1") out1 = need("--out1"); else if (a == "--out2") out2 = need("--out2");
                else if (a == "--orientation") orientation = need("--orientation");
                else if (a == "--low") low = parse_double_arg(need("--low")); else if (a == "--high") high = parse_double_arg(need("--high"));
                else quit_with_error("unexpected argument " + a);
            }
            FilterOutput out;
            filter(in1, in2, out1, out2, orientation, low, high, out);
            FILE* f1 = fopen(out1.c_str(), "wb"); FILE* f2 = fopen(out2.c_str(), "wb");
            if (!f1 || !f2) quit_with_error("unable to write alignments");
            fwrite(out.out1.data(), 1, out.out1.size(), f1); fwrite(out.out2.data(), 1, out.out2.size(), f2);
            fclose(f1); fclose(f2);
            fprintf(stderr, "orientation %s low %u high %u before %zu after %zu\n", out.orientation.c_str(), out.low, out.high, out.before_count, out.after_count);
            return 0;
        }
        fprintf(stderr, "unknown subcommand %s\n", cmd.c_str());
        return 2;
    } catch (const OracleError& e) {
        fprintf(stderr, "\nError: %s\n", e.what());
        return e.code;
    }
}
#endif
