// synth.cpp — deterministic synthetic assembly + multi-mapped paired-read SAM generator.
//
// Measurement/test support, not part of the polishing path: there is no aligner and no dataset in this
// environment, so bench.py and the large parity tests build their inputs here.  A truth genome with planted
// repeat families (copy numbers 7,5,3,2,4 -> non-dyadic k), a draft = truth + planted errors (what polishing
// fixes), 150 bp "fr" read pairs carrying the truth allele plus sequencing errors, and bwa-mem -a style
// records by construction: one SAM per mate, reads in random (not coordinate) order, a read inside a c-copy
// repeat emits c consecutive records (the first with SEQ/QUAL, the rest flag|256 with SEQ="*"), ~0.5 %
// soft-clipped, ~0.5 % NM > 10, ~0.5 % unaligned.  The same records can be written as SAM text or streamed
// straight into the packer (identical text, never materialised on disk).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "pp_internal.h"

namespace {

struct Rng {   // xoshiro256** seeded by splitmix64
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        uint64_t z = seed;
        for (int i = 0; i < 4; ++i) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            s[i] = x ^ (x >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    char base() { return "ACGT"[next() & 3]; }
    char other(char b) { char c; do c = base(); while (c == b); return c; }
};

inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }

struct Copy { uint64_t start, len; uint32_t family; bool rev; int32_t shared = -1; };   // shared: index of a cross-contig family, -1 = this contig only

struct Contig {
    std::string name, truth, draft;
    std::vector<uint32_t> t2d;       // draft index of truth base i (of the next surviving base when deleted)
    std::vector<uint8_t> tdel;       // truth base i is missing from the draft
    std::vector<Copy> copies;        // sorted by start
};

struct Col { uint8_t type; char base; };   // 0 match/mismatch (consumes truth), 1 read lacks the truth base, 2 inserted read base
enum { C_M = 0, C_DEL = 1, C_INS = 2 };

struct Rec { uint32_t pos; std::string cigar; std::string seq; uint32_t nm; bool ok; };

}  // namespace

struct SharedCopy { uint32_t contig; uint64_t start, len; bool rev; };

struct pp_synth {
    pp_synth_params prm;
    std::vector<Contig> contigs;
    std::vector<std::vector<SharedCopy>> shared;   // repeat families whose copies lie on different contigs (config 5: k spans GPUs)
    bool per_pair_rng = false;                     // every pair / mate draws from its own stream (seed, pair index): parts of the data
                                                   // set can be generated without generating the rest
    uint32_t filter_shards = 0, filter_shard = 0;  // > 0: only reads with a record on a contig of `filter_shard` are emitted
    std::vector<uint32_t> shard_of_contig;
    uint64_t total_draft = 0;
    uint64_t n_pairs = 0;
};

namespace {

void plant_repeats(Contig& c, Rng& rng, double frac) {
    const uint64_t T = c.truth.size();
    static const struct { uint32_t len, copies; } fam[] = {{5000, 7}, {1500, 5}, {3000, 3}, {2000, 2}, {1000, 4}};
    double scale = std::min(1.0, std::max(0.1, (double)T / 1e6));
    uint64_t target = (uint64_t)(frac * (double)T), placed = 0;
    uint32_t fid = 0;
    for (int round = 0; placed < target && round < 1000; ++round) {
        for (int f = 0; f < 5 && (placed < target || (round == 0 && f < 3)); ++f) {
            uint64_t len = std::max<uint64_t>(200, (uint64_t)(fam[f].len * scale));
            if (len * fam[f].copies > T / 4) continue;
            std::string seg(len, 'A');
            for (auto& ch : seg) ch = rng.base();
            std::string rc(len, 'A');
            for (uint64_t i = 0; i < len; ++i) rc[i] = comp(seg[len - 1 - i]);
            for (uint32_t k = 0; k < fam[f].copies; ++k) {
                for (int tries = 0; tries < 200; ++tries) {
                    uint64_t st = 200 + rng.below(T - len - 400);
                    bool clash = false;
                    for (auto& o : c.copies) if (st < o.start + o.len + 300 && o.start < st + len + 300) { clash = true; break; }
                    if (clash) continue;
                    bool rev = (k > 0) && (rng.next() & 1);
                    c.truth.replace(st, len, rev ? rc : seg);
                    c.copies.push_back({st, len, fid, rev});
                    placed += len;
                    break;
                }
            }
            fid++;
        }
        if (T < 4000) break;
    }
    std::sort(c.copies.begin(), c.copies.end(), [](const Copy& a, const Copy& b) { return a.start < b.start; });
}

// Repeat families whose copies sit on DIFFERENT contigs (copy numbers 3, 2, 5, ...): reads inside them multi-map across
// contigs, so with contig sharding their k spans GPUs (alignment.rs:283-288).
void plant_shared(pp_synth* S, Rng& rng, double frac) {
    const size_t nc = S->contigs.size();
    if (nc < 2 || frac <= 0) return;
    static const struct { uint32_t len, copies; } fam[] = {{3000, 3}, {2000, 2}, {1500, 5}, {1000, 7}};
    uint64_t total = 0;
    for (auto& c : S->contigs) total += c.truth.size();
    const double scale = std::min(1.0, std::max(0.1, (double)S->contigs[0].truth.size() / 1e6));
    const uint64_t target = (uint64_t)(frac * (double)total);
    uint64_t placed = 0;
    size_t next_contig = 0;
    for (int round = 0; placed < target && round < 100000; ++round) {
        const auto& f = fam[round % 4];
        const uint64_t len = std::max<uint64_t>(200, (uint64_t)(f.len * scale));
        const uint32_t copies = (uint32_t)std::min<size_t>(f.copies, nc);
        std::string seg(len, 'A');
        for (auto& ch : seg) ch = rng.base();
        std::string rc(len, 'A');
        for (uint64_t i = 0; i < len; ++i) rc[i] = comp(seg[len - 1 - i]);
        std::vector<SharedCopy> fc;
        for (uint32_t k = 0; k < copies; ++k) {
            const uint32_t ci = (uint32_t)((next_contig + k) % nc);
            Contig& c = S->contigs[ci];
            const uint64_t T = c.truth.size();
            if (len + 1000 > T) continue;
            for (int tries = 0; tries < 200; ++tries) {
                const uint64_t st = 200 + rng.below(T - len - 400);
                bool clash = false;
                for (auto& o : c.copies) if (st < o.start + o.len + 300 && o.start < st + len + 300) { clash = true; break; }
                if (clash) continue;
                const bool rev = (k > 0) && (rng.next() & 1);
                c.truth.replace(st, len, rev ? rc : seg);
                Copy cp{st, len, 0x40000000u + (uint32_t)S->shared.size(), rev};
                cp.shared = (int32_t)S->shared.size();
                c.copies.push_back(cp);
                fc.push_back({ci, st, len, rev});
                placed += len;
                break;
            }
        }
        next_contig = (next_contig + copies) % nc;
        S->shared.push_back(std::move(fc));
    }
    for (auto& c : S->contigs)
        std::sort(c.copies.begin(), c.copies.end(), [](const Copy& a, const Copy& b) { return a.start < b.start; });
}

void make_draft(Contig& c, Rng& rng, double err) {
    const uint64_t T = c.truth.size();
    c.t2d.assign(T + 1, 0);
    c.tdel.assign(T, 0);
    c.draft.clear();
    c.draft.reserve(T + T / 1000 + 16);
    // choose error sites; half of the indels are moved into a homopolymer run >= 4 when one is near
    std::vector<uint8_t> kind(T, 0);   // 1 sub, 2 draft lacks base, 3 draft has extra base after i
    for (uint64_t i = 50; i + 50 < T; ++i) {
        if (rng.uni() >= err) continue;
        double r = rng.uni();
        uint8_t k = r < 0.5 ? 1 : (r < 0.75 ? 2 : 3);
        uint64_t at = i;
        if (k != 1 && (rng.next() & 1)) {
            for (uint64_t j = i; j + 4 < T - 50 && j < i + 300; ++j)
                if (c.truth[j] == c.truth[j + 1] && c.truth[j] == c.truth[j + 2] && c.truth[j] == c.truth[j + 3]) { at = j + 1; break; }
        }
        kind[at] = k;
    }
    for (uint64_t i = 0; i < T; ++i) {
        c.t2d[i] = (uint32_t)c.draft.size();
        char b = c.truth[i];
        if (kind[i] == 2) { c.tdel[i] = 1; continue; }
        c.draft.push_back(kind[i] == 1 ? rng.other(b) : b);
        if (kind[i] == 3) c.draft.push_back(c.truth[i] == c.truth[i + 1] ? b : rng.base());
    }
    c.t2d[T] = (uint32_t)c.draft.size();
    uint64_t nN = std::max<uint64_t>(1, (uint64_t)((double)c.draft.size() * 2e-6));
    for (uint64_t k = 0; k < nN; ++k) c.draft[100 + rng.below(c.draft.size() - 200)] = 'N';
}

// Aligns read columns, anchored at truth position ts of contig c, against the draft.
void align_cols(const Contig& c, uint64_t ts, const std::vector<Col>& cols, uint32_t clip, Rec& r) {
    struct Op { char op; uint32_t n; };
    std::vector<Op> ops;
    auto push = [&](char op, uint32_t n) { if (!n) return; if (!ops.empty() && ops.back().op == op) ops.back().n += n; else ops.push_back({op, n}); };
    r.seq.clear(); r.nm = 0; r.ok = false; r.pos = 0;
    const uint64_t T = c.truth.size();
    uint64_t i = ts;
    bool have_m = false;
    uint32_t prev_end = 0, emitted = 0;
    for (const Col& col : cols) {
        if (col.type != C_INS && i >= T) return;
        if (emitted < clip) {                                 // soft-clipped prefix
            if (col.type != C_DEL) { push('S', 1); r.seq.push_back(col.base); emitted++; }
            if (col.type != C_INS) i++;
            continue;
        }
        if (col.type == C_INS) {
            if (!have_m) push('S', 1); else { push('I', 1); r.nm++; }
            r.seq.push_back(col.base); emitted++;
            continue;
        }
        if (!c.tdel[i]) {
            uint32_t dp = c.t2d[i];
            if (col.type == C_M) {
                if (!have_m) { r.pos = dp; have_m = true; }
                else if (dp > prev_end) { push('D', dp - prev_end); r.nm += dp - prev_end; }
                push('M', 1);
                r.seq.push_back(col.base); emitted++;
                r.nm += col.base != c.draft[dp];
                prev_end = dp + 1;
            }
        } else if (col.type == C_M) {                          // truth base that the draft lacks -> insertion
            if (!have_m) push('S', 1); else { push('I', 1); r.nm++; }
            r.seq.push_back(col.base); emitted++;
        }
        i++;
    }
    if (!have_m) return;
    // trailing I -> S (bwa clips), never end on D by construction
    if (!ops.empty() && ops.back().op == 'I') { r.nm -= ops.back().n; ops.back().op = 'S'; }
    r.cigar.clear();
    char buf[24];
    for (auto& o : ops) { snprintf(buf, sizeof buf, "%u%c", o.n, o.op); r.cigar += buf; }
    r.ok = true;
}

void revcomp_cols(const std::vector<Col>& in, std::vector<Col>& out) {
    out.assign(in.rbegin(), in.rend());
    for (auto& c : out) if (c.type != C_DEL) c.base = comp(c.base);
}

const Copy* containing_copy(const Contig& c, uint64_t ts, uint64_t tlen) {
    size_t lo = 0, hi = c.copies.size();
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (c.copies[mid].start + c.copies[mid].len <= ts) lo = mid + 1; else hi = mid; }
    if (lo < c.copies.size() && c.copies[lo].start <= ts && ts + tlen <= c.copies[lo].start + c.copies[lo].len) return &c.copies[lo];
    return nullptr;
}

using Sink = std::function<bool(const char*, size_t)>;

// Generates every record of one mate (1 or 2), in pair order, into `sink` (called with chunks of whole lines).
bool generate(const pp_synth* S, int mate, const Sink& sink) {
    const pp_synth_params& P = S->prm;
    Rng rng(P.seed * 0x9E3779B97F4A7C15ull + 12345);
    std::string out;
    out.reserve(1 << 22);
    out += "@HD\tVN:1.6\tSO:unsorted\n";
    for (auto& c : S->contigs) out += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.draft.size()) + "\n";
    std::vector<double> cum;
    double tot = 0;
    for (auto& c : S->contigs) { tot += (double)c.truth.size(); cum.push_back(tot); }
    std::vector<Col> cols[2], rcols;
    Rec rec;
    const uint32_t L = P.read_len;
    const std::string qual(L + 64, 'I');
    char line[256];
    const bool v2 = S->per_pair_rng;
    const bool filt = v2 && S->filter_shards > 0;
    auto mine = [&](size_t contig_index) { return S->shard_of_contig[contig_index] == S->filter_shard; };
    std::string grp;                                             // the records of one read (primary + secondaries)
    for (uint64_t pi = 0; pi < S->n_pairs; ++pi) {
        // contig, fragment
        Rng pair_rng(v2 ? (P.seed * 0xD1342543DE82EF95ull + pi * 0x9E3779B97F4A7C15ull + 77) : 0);
        Rng& prng = v2 ? pair_rng : rng;
        double x = prng.uni() * tot;
        size_t ci = 0;
        while (ci + 1 < cum.size() && x >= cum[ci]) ci++;
        const Contig& c = S->contigs[ci];
        const uint64_t T = c.truth.size();
        double isz_d = P.insert_mean + P.insert_sd * prng.normal();
        uint64_t isz = (uint64_t)std::min(700.0, std::max(200.0, isz_d));
        if (isz < L + 10) isz = L + 10;
        if (isz + 2 * L + 64 > T) isz = T > 3 * (uint64_t)L + 128 ? T - 2 * L - 64 : L + 10;
        uint64_t fs = prng.below(T - isz - L - 32);
        bool first_is_left = prng.next() & 1;
        // legacy streams: both mates are always generated so that the RNG stream is identical for the mate 1 and mate 2 passes;
        // per-pair streams: every mate has its own, the other one is simply not generated
        for (int m = 0; m < 2; ++m) {
            const bool left = (m == 0);                      // m = 0: leftmost (forward) read, m = 1: rightmost (reverse)
            const int this_mate = (left == first_is_left) ? 1 : 2;
            if (v2 && this_mate != mate) continue;
            const uint64_t ts = left ? fs : fs + isz - L;
            if (filt && !mine(ci)) {
                // can this read have a record on one of my contigs?  Only through a cross-contig repeat family whose copy
                // contains the read's start (a superset of the exact rule below)
                const Copy* near = containing_copy(c, ts, 1);
                bool maybe = false;
                if (near && near->shared >= 0)
                    for (auto& o : S->shared[(size_t)near->shared]) maybe |= mine(o.contig);
                if (!maybe) continue;
            }
            Rng mate_rng(v2 ? (P.seed * 0xA0761D6478BD642Full + (pi * 2 + (uint64_t)m) * 0xE7037ED1A0B428DBull + 5) : 0);
            Rng& mrng = v2 ? mate_rng : rng;
            std::vector<Col>& cl = cols[m];
            cl.clear();
            uint32_t nb = 0;
            uint64_t tlen = 0;
            double special = mrng.uni();
            const bool unaligned = special < P.unaligned_rate;
            const bool clipped = !unaligned && special < P.unaligned_rate + P.clip_rate;
            const bool highnm = !unaligned && !clipped && special < P.unaligned_rate + P.clip_rate + P.highnm_rate;
            while (nb < L && ts + tlen < T) {
                double r = mrng.uni();
                char b = c.truth[ts + tlen];
                if (r < P.seq_indel_rate * 0.5) { cl.push_back({C_DEL, 0}); tlen++; continue; }
                if (r < P.seq_indel_rate * 0.5 + P.seq_sub_rate) b = mrng.other(b);
                cl.push_back({C_M, b}); nb++; tlen++;
                if (nb < L && mrng.uni() < P.seq_indel_rate * 0.5) { cl.push_back({C_INS, mrng.base()}); nb++; }
            }
            while (!cl.empty() && cl.back().type == C_DEL) { cl.pop_back(); tlen--; }
            if (highnm) for (int k = 0; k < 14; ++k) { Col& q = cl[mrng.below(cl.size())]; if (q.type == C_M) q.base = mrng.other(q.base); }
            const uint32_t clipn = clipped ? 3 + (uint32_t)mrng.below(8) : 0;
            const bool clip_left = mrng.next() & 1;
            if (this_mate != mate) continue;
            const bool rev = !left;
            int n = snprintf(line, sizeof line, "r%llu", (unsigned long long)pi);
            const std::string qname(line, n);
            if (unaligned) {
                if (filt && !mine(ci)) continue;
                std::string seq;
                for (auto& q : cl) if (q.type != C_DEL) seq.push_back(q.base);
                out += qname + "\t4\t*\t0\t0\t*\t*\t0\t0\t" + seq + "\t" + qual.substr(0, seq.size()) + "\n";
                continue;
            }
            // primary
            if (clipped && !clip_left) {
                // soft clip at the right end: align a shortened column list, then append the clipped bases
                std::vector<Col> shortc(cl);
                std::string tail;
                uint32_t cut = 0;
                while (cut < clipn && !shortc.empty()) {
                    Col q = shortc.back(); shortc.pop_back();
                    if (q.type != C_DEL) { tail.insert(tail.begin(), q.base); cut++; }
                }
                while (!shortc.empty() && shortc.back().type != C_M) {
                    Col q = shortc.back(); shortc.pop_back();
                    if (q.type == C_INS) { tail.insert(tail.begin(), q.base); cut++; }
                }
                align_cols(c, ts, shortc, 0, rec);
                if (rec.ok && rec.cigar.back() == 'S') rec.ok = false;
                if (rec.ok) { rec.cigar += std::to_string(cut) + "S"; rec.seq += tail; }
            } else {
                align_cols(c, ts, cl, clipped ? clipn : 0, rec);
            }
            if (!rec.ok) continue;
            bool touches = !filt || mine(ci);
            grp.clear();
            const Copy* cp = containing_copy(c, ts, tlen);
            struct Other { const Contig* c; uint32_t contig; uint64_t start, len; bool rev; };
            std::vector<Other> others;
            if (cp && !clipped) {
                if (cp->shared < 0) { for (auto& o : c.copies) if (o.family == cp->family && &o != cp) others.push_back({&c, (uint32_t)ci, o.start, o.len, o.rev}); }
                else for (auto& o : S->shared[(size_t)cp->shared])
                    if (o.contig != ci || o.start != cp->start) others.push_back({&S->contigs[o.contig], o.contig, o.start, o.len, o.rev});
            }
            n = snprintf(line, sizeof line, "\t%d\t%s\t%u\t%d\t", rev ? 16 : 0, c.name.c_str(), rec.pos + 1, others.empty() ? 60 : 0);
            grp += qname; grp.append(line, n); grp += rec.cigar; grp += "\t*\t0\t0\t"; grp += rec.seq; grp += '\t';
            grp.append(qual.data(), rec.seq.size());
            n = snprintf(line, sizeof line, "\tNM:i:%u\tAS:i:%d\tXS:i:%d\n", rec.nm, (int)rec.seq.size() - 5 * (int)rec.nm, others.empty() ? 0 : (int)rec.seq.size() - 5 * (int)rec.nm);
            grp.append(line, n);
            // secondaries: the same read against the other copies of the repeat family
            const uint64_t off = cp ? ts - cp->start : 0;
            for (const Other& o : others) {
                const bool flip = o.rev != cp->rev;
                uint64_t ts2;
                const std::vector<Col>* c2 = &cl;
                if (!flip) ts2 = o.start + off;
                else { ts2 = o.start + (o.len - off - tlen); revcomp_cols(cl, rcols); c2 = &rcols; }
                align_cols(*o.c, ts2, *c2, 0, rec);
                if (!rec.ok) continue;
                if (filt && mine(o.contig)) touches = true;
                const bool rev2 = rev != flip;
                n = snprintf(line, sizeof line, "\t%d\t%s\t%u\t0\t", 256 | (rev2 ? 16 : 0), o.c->name.c_str(), rec.pos + 1);
                grp += qname; grp.append(line, n); grp += rec.cigar;
                n = snprintf(line, sizeof line, "\t*\t0\t0\t*\t*\tNM:i:%u\tAS:i:%d\n", rec.nm, (int)rec.seq.size() - 5 * (int)rec.nm);
                grp.append(line, n);
            }
            if (touches) out += grp;
        }
        if (out.size() > (1u << 22)) { if (!sink(out.data(), out.size())) return false; out.clear(); }
    }
    if (!out.empty() && !sink(out.data(), out.size())) return false;
    return true;
}

}  // namespace

// cross_contig_fraction = 0 gives exactly pp_synth_create's data (every contig keeps its own RNG stream).
extern "C" pp_synth* pp_synth_create_shared(const pp_synth_params* prm, double cross_contig_fraction) {
    if (!prm || prm->n_contigs == 0 || prm->contig_len < 2000 || prm->read_len < 30 || prm->read_len > 1000) return nullptr;
    pp_synth* S = new pp_synth();
    S->prm = *prm;
    std::vector<Rng> rngs;
    for (uint32_t ci = 0; ci < prm->n_contigs; ++ci) {
        rngs.emplace_back(prm->seed * 1000003ull + 0xB2000001ull + ci);
        Contig c;
        c.name = "contig_" + std::to_string(ci + 1);
        c.truth.resize(prm->contig_len);
        for (auto& ch : c.truth) ch = rngs[ci].base();
        plant_repeats(c, rngs[ci], prm->repeat_fraction);
        S->contigs.push_back(std::move(c));
    }
    if (cross_contig_fraction > 0) {
        Rng srng(prm->seed * 7919ull + 0xC5C5C5C5ull);
        plant_shared(S, srng, cross_contig_fraction);
        S->per_pair_rng = true;
    }
    for (uint32_t ci = 0; ci < prm->n_contigs; ++ci) {
        make_draft(S->contigs[ci], rngs[ci], prm->draft_error_rate);
        S->total_draft += S->contigs[ci].draft.size();
    }
    S->n_pairs = (uint64_t)(prm->depth * (double)S->total_draft / (2.0 * prm->read_len));
    if (S->n_pairs == 0) S->n_pairs = 1;
    return S;
}

extern "C" pp_synth* pp_synth_create(const pp_synth_params* prm) { return pp_synth_create_shared(prm, 0.0); }

// Only for data sets made by pp_synth_create_shared with a cross-contig fraction > 0 (per-pair random streams): from now on
// pp_synth_write_sam / pp_synth_feed_pack emit only the reads that have a record on a contig of `shard` - exactly the reads the
// contig sharder would hand that shard (pp_shards_build_assigned with the same assignment), without generating the others.
extern "C" int pp_synth_set_shard_filter(pp_synth* s, uint32_t n_shards, uint32_t shard, const uint32_t* shard_of_contig) {
    if (!s || !s->per_pair_rng || (n_shards && shard >= n_shards)) return PP_ERR_ARG;
    s->filter_shards = n_shards;
    s->filter_shard = shard;
    s->shard_of_contig.resize(s->contigs.size());
    for (size_t i = 0; i < s->contigs.size(); ++i) s->shard_of_contig[i] = n_shards ? (shard_of_contig ? shard_of_contig[i] : (uint32_t)(i % n_shards)) : 0u;
    return PP_OK;
}

extern "C" void pp_synth_free(pp_synth* s) { delete s; }
extern "C" uint64_t pp_synth_total_bp(const pp_synth* s) { return s ? s->total_draft : 0; }
extern "C" uint64_t pp_synth_n_pairs(const pp_synth* s) { return s ? s->n_pairs : 0; }

extern "C" int pp_synth_write_fasta(const pp_synth* s, const char* path, int truth) {
    FILE* f = fopen(path, "wb");
    if (!f) return PP_ERR_IO;
    for (auto& c : s->contigs) {
        const std::string& q = truth ? c.truth : c.draft;
        fprintf(f, ">%s synthetic len=%zu\n", c.name.c_str(), q.size());
        for (size_t i = 0; i < q.size(); i += 80) { fwrite(q.data() + i, 1, std::min<size_t>(80, q.size() - i), f); fputc('\n', f); }
    }
    return fclose(f) == 0 ? PP_OK : PP_ERR_IO;
}

extern "C" pp_fasta* pp_synth_fasta(const pp_synth* s) {
    pp_fasta* fa = new pp_fasta();
    for (auto& c : s->contigs) {
        fa->index.emplace(c.name, (uint32_t)fa->names.size());
        fa->names.push_back(c.name);
        fa->descriptions.push_back("synthetic len=" + std::to_string(c.draft.size()));
        fa->off.push_back(fa->bases.size());
        fa->bases += c.draft;
    }
    fa->off.push_back(fa->bases.size());
    return fa;
}

extern "C" int pp_synth_write_sam(const pp_synth* s, int mate, const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) return PP_ERR_IO;
    bool ok = generate(s, mate, [&](const char* p, size_t n) { return fwrite(p, 1, n, f) == n; });
    return (fclose(f) == 0 && ok) ? PP_OK : PP_ERR_IO;
}

extern "C" int pp_synth_feed_pack(const pp_synth* s, int mate, pp_pack* pack) {
    std::string name = "synthetic_" + std::to_string(mate) + ".sam";
    int rc = pp_pack_stream_begin(pack, name.c_str());
    if (rc) return rc;
    int frc = PP_OK;
    generate(s, mate, [&](const char* p, size_t n) { frc = pp_pack_stream_feed(pack, p, n); return frc == PP_OK; });
    if (frc) return frc;
    return pp_pack_stream_end(pack);
}
