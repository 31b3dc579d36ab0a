// shard.cpp — contig sharding of a packed polish job across GPUs: no collective on the data path.
//
// Whole contigs go to shards (longest-processing-time bin packing on aligned bases); every shard receives, in SAM
// order, the alignments that land on its contigs.  Two things span contigs and are settled here, before the device
// ever sees a shard:
//   * k, the number of good alignments of a read (alignment.rs:283-288): a read multi-mapped to contigs on two
//     shards still has ONE k.  Each shard therefore also receives the read's foreign alignments as GHOSTS
//     (PP_FLAG_GHOST): they take part in goodness / k / --careful exactly like any other record, but add nothing to
//     the pileup and are not counted as used.
//   * the sequence of SEQ="*" records (alignment.rs:290-295): the source sequence is copied into every shard that
//     needs it.
// Alignments whose RNAME is not in the assembly stay in shard 0 so that the reference's error (alignment.rs:298-300)
// is still raised, once.
#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
#include <utility>
#include <vector>

#include "pp_internal.h"

struct pp_shards {
    struct Shard {
        std::vector<uint32_t> contig_map;            // original index of each local contig
        std::vector<uint64_t> off;
        std::vector<uint8_t> bases;
        std::vector<uint32_t> contig, ref_start, read_id, seq_off, cigar_off, nm, cigar_ops;
        std::vector<uint16_t> seq_len, n_cigar;
        std::vector<uint8_t> flags;
        pp::AlignedBytes seq_pool;
        uint64_t seq_blocks = 0, n_reads = 0, n_home = 0;
    };
    std::vector<Shard> shards;
    uint32_t seq_bits = 4;
};

extern "C" pp_shards* pp_shards_build(const pp_contigs* c, const pp_alignments* a, uint32_t n_shards) {
    return pp_shards_build_assigned(c, a, n_shards, nullptr, -1);
}

extern "C" pp_shards* pp_shards_build_assigned(const pp_contigs* c, const pp_alignments* a, uint32_t n_shards,
                                               const uint32_t* assigned, int32_t only_shard) {
    if (!c || !a || n_shards == 0 || only_shard >= (int32_t)n_shards) return nullptr;
    if (a->seq_bits != 4 && a->seq_bits != 8) return nullptr;      // (the 2-bit wire format is made per shard, after sharding)
    if (assigned) for (uint32_t i = 0; i < c->n_contigs; ++i) if (assigned[i] >= n_shards) return nullptr;
    const uint32_t nc = c->n_contigs;
    pp_shards* S = new pp_shards();
    S->seq_bits = a->seq_bits;
    S->shards.resize(n_shards);
    const size_t block_bytes = a->seq_bits == 4 ? PP_SEQ_BLOCK / 2 : PP_SEQ_BLOCK;
    // weights: aligned bases + contig length
    std::vector<uint64_t> weight(nc);
    for (uint32_t i = 0; i < nc; ++i) weight[i] = c->off[i + 1] - c->off[i];
    for (uint64_t i = 0; i < a->n_aln; ++i) if (a->contig[i] != PP_CONTIG_UNKNOWN && a->contig[i] < nc) weight[a->contig[i]] += a->seq_len[i];
    std::vector<uint32_t> order(nc);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return weight[x] > weight[y]; });
    std::vector<uint64_t> load(n_shards, 0);
    std::vector<uint32_t> shard_of(nc), local_of(nc);
    for (uint32_t ci : order) {
        uint32_t best = assigned ? assigned[ci] : (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
        shard_of[ci] = best;
        load[best] += weight[ci];
    }
    for (uint32_t ci = 0; ci < nc; ++ci) {                    // contigs keep their input order inside a shard
        auto& sh = S->shards[shard_of[ci]];
        local_of[ci] = (uint32_t)sh.contig_map.size();
        sh.contig_map.push_back(ci);
        sh.off.push_back(sh.bases.size());
        sh.bases.insert(sh.bases.end(), c->bases + c->off[ci], c->bases + c->off[ci + 1]);
    }
    for (auto& sh : S->shards) sh.off.push_back(sh.bases.size());

    // Every shard is built independently of the others (it only reads the inputs), so each gets its own thread: one pass over
    // the read groups to size its arrays, one to fill them.
    auto shard_of_rec = [&](uint64_t i) -> uint32_t {
        const uint32_t ci = a->contig[i];
        return (ci == PP_CONTIG_UNKNOWN || ci >= nc) ? 0u : shard_of[ci];
    };
    auto build_one = [&](uint32_t s) {
        auto& sh = S->shards[s];
        uint64_t n_rec = 0, n_ops = 0, n_blk = 0;
        for (uint64_t g0 = 0; g0 < a->n_aln;) {
            uint64_t g1 = g0 + 1;
            while (g1 < a->n_aln && a->read_id[g1] == a->read_id[g0]) g1++;
            bool touched = false;
            for (uint64_t i = g0; i < g1 && !touched; ++i) touched = shard_of_rec(i) == s;
            if (touched) {
                n_rec += g1 - g0;
                for (uint64_t i = g0; i < g1; ++i) {
                    n_ops += a->n_cigar[i];
                    if (shard_of_rec(i) == s && !(a->flags[i] & PP_FLAG_NOSEQ)) n_blk += ((size_t)a->seq_len[i] + PP_SEQ_BLOCK - 1) / PP_SEQ_BLOCK;
                }
            }
            g0 = g1;
        }
        sh.contig.reserve(n_rec); sh.ref_start.reserve(n_rec); sh.read_id.reserve(n_rec); sh.seq_off.reserve(n_rec);
        sh.seq_len.reserve(n_rec); sh.cigar_off.reserve(n_rec); sh.n_cigar.reserve(n_rec); sh.nm.reserve(n_rec); sh.flags.reserve(n_rec);
        sh.cigar_ops.reserve(n_ops);
        sh.seq_pool.reserve(n_blk * block_bytes + 64);        // an upper bound: SEQ="*" records share their group's copy
        auto copy_seq = [&](uint32_t off_blk, uint32_t len) -> uint32_t {
            const size_t blocks = ((size_t)len + PP_SEQ_BLOCK - 1) / PP_SEQ_BLOCK;
            const uint32_t at = (uint32_t)sh.seq_blocks;
            const size_t base = sh.seq_blocks * block_bytes;
            sh.seq_pool.resize_zero(base + blocks * block_bytes);
            if (blocks) memcpy(sh.seq_pool.p + base, a->seq_pool + (size_t)off_blk * block_bytes, blocks * block_bytes);
            sh.seq_blocks += blocks;
            return at;
        };
        std::vector<std::pair<uint32_t, uint32_t>> seq_at;      // of the current group: original seq_off -> shard seq_off
        for (uint64_t g0 = 0; g0 < a->n_aln;) {
            uint64_t g1 = g0 + 1;
            while (g1 < a->n_aln && a->read_id[g1] == a->read_id[g0]) g1++;
            bool touched = false;
            for (uint64_t i = g0; i < g1 && !touched; ++i) touched = shard_of_rec(i) == s;
            if (touched) {
                seq_at.clear();
                const uint32_t rid = (uint32_t)sh.n_reads++;
                for (uint64_t i = g0; i < g1; ++i) {
                    const uint32_t ci = a->contig[i];
                    const bool unknown = (ci == PP_CONTIG_UNKNOWN || ci >= nc);
                    const bool home = shard_of_rec(i) == s;
                    uint8_t fl = a->flags[i];
                    uint32_t soff = 0;
                    if (home) {
                        sh.n_home++;
                        if (!(fl & PP_FLAG_NOSEQ)) {
                            const uint32_t key = a->seq_off[i];
                            const std::pair<uint32_t, uint32_t>* hit = nullptr;
                            for (const auto& kv : seq_at) if (kv.first == key) { hit = &kv; break; }
                            if (hit && (fl & PP_FLAG_SEQSTAR)) soff = hit->second;
                            else {
                                soff = copy_seq(key, a->seq_len[i]);
                                if (!hit) seq_at.emplace_back(key, soff);
                            }
                        }
                    } else {
                        fl |= PP_FLAG_GHOST;
                    }
                    sh.contig.push_back(home ? (unknown ? PP_CONTIG_UNKNOWN : local_of[ci]) : 0u);
                    sh.ref_start.push_back(a->ref_start[i]);
                    sh.read_id.push_back(rid);
                    sh.seq_off.push_back(soff);
                    sh.seq_len.push_back(a->seq_len[i]);
                    sh.cigar_off.push_back((uint32_t)sh.cigar_ops.size());
                    sh.n_cigar.push_back(a->n_cigar[i]);
                    sh.nm.push_back(a->nm[i]);
                    sh.flags.push_back(fl);
                    sh.cigar_ops.insert(sh.cigar_ops.end(), a->cigar_ops + a->cigar_off[i], a->cigar_ops + a->cigar_off[i] + a->n_cigar[i]);
                }
            }
            g0 = g1;
        }
    };
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned n_threads = std::min<unsigned>(n_shards, hw);
        std::atomic<uint32_t> next{0};
        auto worker = [&]() { for (uint32_t s; (s = next.fetch_add(1)) < n_shards;) if (only_shard < 0 || (uint32_t)only_shard == s) build_one(s); };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(worker);
        worker();
        for (auto& t : th) t.join();
    }
    return S;
}

extern "C" void pp_shards_free(pp_shards* s) { delete s; }

extern "C" int pp_shards_get(const pp_shards* S, uint32_t i, pp_contigs* c, pp_alignments* a, const uint32_t** contig_map,
                             uint64_t* n_home) {
    if (!S || i >= S->shards.size() || !c || !a) return PP_ERR_ARG;
    const auto& sh = S->shards[i];
    c->n_contigs = (uint32_t)sh.contig_map.size();
    c->off = sh.off.data();
    c->bases = sh.bases.data();
    memset(a, 0, sizeof *a);
    a->n_aln = sh.contig.size();
    a->n_reads = sh.n_reads;
    a->contig = sh.contig.data(); a->ref_start = sh.ref_start.data(); a->read_id = sh.read_id.data();
    a->seq_off = sh.seq_off.data(); a->seq_len = sh.seq_len.data(); a->cigar_off = sh.cigar_off.data();
    a->n_cigar = sh.n_cigar.data(); a->nm = sh.nm.data(); a->flags = sh.flags.data();
    a->n_cigar_ops = sh.cigar_ops.size(); a->cigar_ops = sh.cigar_ops.data();
    a->seq_bits = S->seq_bits; a->seq_pool_bytes = sh.seq_pool.n; a->seq_pool = sh.seq_pool.p;
    if (contig_map) *contig_map = sh.contig_map.data();
    if (n_home) *n_home = sh.n_home;
    return PP_OK;
}
