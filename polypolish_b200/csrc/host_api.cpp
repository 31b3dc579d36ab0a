// host_api.cpp — whole-command drivers above the C-ABI compute calls.
//
// Mirrors the reference's command drivers (same option checks, same error text, same stdout bytes):
//   polish::polish            /root/reference/src/polish.rs:26-38   (+ :93-134 loading, :137-203 output)
//   filter::filter            /root/reference/src/filter.rs:26-37   (+ :273-349 SAM re-streaming)
// Text (FASTA/SAM) is handled here on the host; all per-alignment / per-position work is behind
// pp_polish() / pp_filter() on the device.  There is no CPU fallback for that work.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pp_internal.h"

namespace {

std::string fmt_thousands(uint64_t v) {      // num_format Locale::en
    std::string s = std::to_string(v), o;
    int n = (int)s.size();
    for (int i = 0; i < n; ++i) {
        o += s[i];
        if ((n - 1 - i) % 3 == 0 && i != n - 1) o += ',';
    }
    return o;
}

}  // namespace

extern "C" void pp_free(void* p) { free(p); }

extern "C" int pp_polish_files(pp_ctx* ctx, const char* assembly, const char* const* sams, int n_sams,
                               const pp_polish_params* prm, const char* debug_path, char** out_fasta,
                               uint64_t* out_len, int verbose) {
    if (!ctx) return PP_ERR_ARG;
    if (!assembly || !prm || !out_fasta || !out_len || n_sams < 0) return pp_ctx_fail(ctx, PP_ERR_ARG, "pp_polish_files: bad arguments");
    *out_fasta = nullptr;
    *out_len = 0;
    // check_option_values polish.rs:277-287
    if (!(prm->fraction_valid > 0.0 && prm->fraction_valid < 1.0)) return pp_ctx_fail(ctx, PP_ERR_INPUT, "--fraction_valid must be between 0 and 1 (exclusive)");
    if (!(prm->fraction_invalid > 0.0 && prm->fraction_invalid < 1.0)) return pp_ctx_fail(ctx, PP_ERR_INPUT, "--fraction_invalid must be between 0 and 1 (exclusive)");
    if (prm->fraction_invalid >= prm->fraction_valid) return pp_ctx_fail(ctx, PP_ERR_INPUT, "--fraction_invalid must be less than --fraction_valid");
    // check_inputs_exist polish.rs:269-274
    if (!pp::file_exists(assembly)) return pp_ctx_fail(ctx, PP_ERR_INPUT, ("\"" + std::string(assembly) + "\" file does not exist").c_str());
    for (int i = 0; i < n_sams; ++i)
        if (!pp::file_exists(sams[i])) return pp_ctx_fail(ctx, PP_ERR_INPUT, ("\"" + std::string(sams[i]) + "\" file does not exist").c_str());
    if (debug_path && debug_path[0])
        return pp_ctx_fail(ctx, PP_ERR_ARG, "--debug (per-base TSV, polish.rs:230-266) is not implemented in this build");

    char ebuf[1024];
    pp_fasta* fa = pp_fasta_load(assembly, ebuf, sizeof ebuf);
    if (!fa) return pp_ctx_fail(ctx, PP_ERR_INPUT, ebuf);
    pp_contigs contigs;
    pp_fasta_view(fa, &contigs);
    if (verbose) {
        fprintf(stderr, "Loading assembly\n");
        for (uint32_t i = 0; i < contigs.n_contigs; ++i)
            fprintf(stderr, "%s (%s bp)\n", pp_fasta_name(fa, i), fmt_thousands(contigs.off[i + 1] - contigs.off[i]).c_str());
        fprintf(stderr, "\nLoading alignments\n");
    }

    pp_pack* pk = pp_pack_create(fa, prm->careful);
    int rc = PP_OK;
    for (int i = 0; i < n_sams && rc == PP_OK; ++i) {
        rc = pp_pack_add_sam_file(pk, sams[i]);
        if (rc == PP_OK && verbose) {
            uint64_t na = 0, nr = 0;
            pp_pack_file_stats(pk, (uint32_t)i, &na, &nr);
            fprintf(stderr, "%s: %s alignments from %s reads\n", sams[i], fmt_thousands(na).c_str(), fmt_thousands(nr).c_str());
        }
    }
    pp_alignments alns;
    if (rc == PP_OK) rc = pp_pack_finish(pk, &alns);
    if (rc != PP_OK) {
        rc = pp_ctx_fail(ctx, rc, pp_pack_error(pk));
        pp_pack_free(pk);
        pp_fasta_free(fa);
        return rc;
    }

    const uint64_t G = contigs.off[contigs.n_contigs];
    std::vector<uint64_t> out_off(contigs.n_contigs + 1), changed(contigs.n_contigs), zero(contigs.n_contigs);
    std::vector<uint8_t> bases;
    pp_polish_result res;
    memset(&res, 0, sizeof res);
    // Output is at most G + inserted bases; start with G + 1 MiB and retry once with the exact size.
    uint64_t cap = G + (1u << 20);
    for (int attempt = 0; attempt < 2; ++attempt) {
        bases.resize(cap);
        res.out_off = out_off.data();
        res.out_bases = bases.data();
        res.out_cap = cap;
        res.changed = changed.data();
        res.zero_depth = zero.data();
        rc = pp_polish(ctx, &contigs, &alns, prm, &res);
        if (rc == PP_ERR_ARG && res.out_len > cap) { cap = res.out_len; continue; }
        break;
    }
    if (rc != PP_OK) {
        if (rc == PP_ERR_INPUT && res.error_aln >= 0) {
            // re-word device-detected errors with the names the reference prints (alignment.rs:190-198,298-300)
            std::string m = pp_last_error(ctx);
            const char* rn = pp_pack_read_name(pk, (uint64_t)res.error_aln);
            if (m.rfind("query name", 0) == 0)
                m = "query name " + std::string(pp_pack_unknown_ref(pk, (uint64_t)res.error_aln)) + " in SAM but not in assembly";
            else if (m.rfind("CIGAR string does not", 0) == 0)
                m = "CIGAR string for read " + std::string(rn) + " does not match read sequence";
            else if (m.rfind("unexpected character", 0) == 0) {
                char cg[4096];
                pp_pack_cigar_string(pk, (uint64_t)res.error_aln, cg, sizeof cg);
                m = "unexpected character (other than M, =, X, I or D) in CIGAR string for read " + std::string(rn) +
                    ": \"" + cg + "\" - did you use BWA MEM to generate your alignments?";
            }
            else
                m += " (read " + std::string(rn) + ")";
            rc = pp_ctx_fail(ctx, rc, m.c_str());
        }
        pp_pack_free(pk);
        pp_fasta_free(fa);
        return rc;
    }
    if (verbose) {
        fprintf(stderr, "\nFiltering for high-quality end-to-end alignments%s:\n", prm->careful ? " from reads with only one alignment" : "");
        fprintf(stderr, "  %s alignments kept\n", fmt_thousands(res.n_aln_used).c_str());
        fprintf(stderr, "  %s alignments discarded\n\n", fmt_thousands(alns.n_aln - res.n_aln_used).c_str());
    }

    // print_seq_to_stdout polish.rs:196-203
    std::string out;
    out.reserve(res.out_len + 128 * (size_t)contigs.n_contigs);
    for (uint32_t i = 0; i < contigs.n_contigs; ++i) {
        out += '>';
        out += pp_fasta_name(fa, i);
        const char* d = pp_fasta_description(fa, i);
        if (d[0]) { out += ' '; out += d; }
        out += " polypolish\n";
        out.append((const char*)bases.data() + out_off[i], out_off[i + 1] - out_off[i]);
        out += '\n';
        if (verbose) {
            uint64_t len = contigs.off[i + 1] - contigs.off[i];
            fprintf(stderr, "Polishing %s (%s bp):\n", pp_fasta_name(fa, i), fmt_thousands(len).c_str());
            fprintf(stderr, "  %s bp %s a depth of zero (%.4f%% coverage)\n", fmt_thousands(zero[i]).c_str(), zero[i] == 1 ? "has" : "have",
                    100.0 * (double)(len - zero[i]) / (double)len);
            fprintf(stderr, "  %s %s changed (%.4f%% of total positions)\n\n", fmt_thousands(changed[i]).c_str(),
                    changed[i] == 1 ? "position" : "positions", 100.0 * (double)changed[i] / (double)len);
        }
    }
    if (verbose) {
        fprintf(stderr, "device path: %.3f ms total (h2d %.3f, classify %.3f, scatter %.3f, fix-up %.3f, vote %.3f, d2h %.3f), %u kernels\n",
                res.timing.total_ms, res.timing.stage_ms[6], res.timing.stage_ms[1], res.timing.stage_ms[2], res.timing.stage_ms[3],
                res.timing.stage_ms[5], res.timing.stage_ms[7], res.timing.launches);
    }
    char* buf = (char*)malloc(out.size() + 1);
    if (!buf) { pp_pack_free(pk); pp_fasta_free(fa); return pp_ctx_fail(ctx, PP_ERR_NOMEM, "out of memory"); }
    memcpy(buf, out.data(), out.size());
    buf[out.size()] = 0;
    *out_fasta = buf;
    *out_len = out.size();
    pp_pack_free(pk);
    pp_fasta_free(fa);
    return PP_OK;
}
