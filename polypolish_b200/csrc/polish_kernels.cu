// polish_kernels.cu — the polish hot path as sm_100a kernels + the C-ABI entry points that drive them.
//
// Replaces, on the device (reference = /root/reference/src):
//   process_one_read            alignment.rs:275-305  -> k_scatter stage 1 (goodness, k = #good per read group)
//   get_read_bases_for_each_target_base + trim_bases_for_homopolymers
//                               alignment.rs:175-201, 364-378 -> k_scatter stage 2 (CIGAR walk, right-end trim)
//   Pileup::add_alignment / PileupBase::add_seq   pileup.rs:189-200, 56-65 -> k_scatter (+ k_collect / k_depth_fixup)
//   PileupBase::get_polished_seq + bankers_rounding pileup.rs:67-134, misc.rs:208-215 -> k_vote
//   polish_one_sequence's join + replace("-","")  polish.rs:185-188 -> k_compact
//
// Design (DESIGN.md has the derivation): the reference adds one counter per aligned base (5e8 increments for
// 5 Mbp x 100x), which on any GPU is bound by atomic throughput, not by HBM.  Here the pileup of a position is
// kept in a form that needs ~3 atomics per ALIGNMENT instead of ~150:
//   * cover[p]  = number of good alignments whose kept entries include p  -> interval add (+v at start,
//     -v at end) into a difference array, prefix-summed by k_diff_sums + a device scan (chunk level, side stream) and k_vote;
//   * explicit[p][allele] = entries whose allele differs from the draft base -> one atomic per mismatch
//     (~0.3 % of bases); count[draft base] = cover - sum(explicit);
//   * alleles other than A,C,G,T,"-" (N / IUPAC bases, insertions) -> one node per distinct (position, allele) in a
//     per-position chain with an exact count (the reference's HashMap<String,u32>, pileup.rs:40,62);
//   * depth: where every covering alignment has k == 1 the f64 depth equals cover exactly; positions covered
//     by a multi-mapped read (k != 1) get the reference's sequential f64 sum re-done in SAM order by
//     k_collect_count / k_collect (pairs in alignment order) -> stable sort by tile -> k_fix_runs -> k_depth_fixup
//     (ordered walk over the alignments binned to that 128-position tile).
// The whole call is one stream of kernels with no host round trip in the middle.  All of it is integer / byte work
// bounded by HBM bandwidth: no tensor cores.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "nib_utils.h"
#include "pp_internal.h"
#include "pp_ctx.cuh"

#define PP_TILE_SHIFT 7              // depth fix-up tile = 128 positions
#define PP_TILE (1u << PP_TILE_SHIFT)
#define SC_THREADS 128               // scatter CTA: one alignment per thread
#define SC_SEQ_BYTES 12288           // smem window for the block's slice of the sequence pool (4-bit mode)
#define SC_GROUP_SCAN_LIMIT 8192     // alignments of one read group a thread will scan outside its block for k
#define CL_THREADS 256               // collect CTA
#define CL_ITEMS 8
#define CL_CHUNK (CL_THREADS * CL_ITEMS)
#define VT_THREADS 256
#define VT_ITEMS 8
#define VT_CHUNK (VT_THREADS * VT_ITEMS)
#define NONE32 0xFFFFFFFFu

enum : unsigned {
    ERR_UNKNOWN_CONTIG = 1, ERR_SEQ_MISMATCH = 2, ERR_BAD_OP = 3, ERR_OOB = 4, ERR_NOSEQ = 5
};
enum : unsigned { FL_NODE_OVF = 1, FL_FIX_OVF = 2, FL_COUNTER_OVF = 4, FL_OUT_OVF = 8, FL_BIGGROUP = 16 };

struct DevStatus {
    unsigned long long err;          // min over (aln << 8 | code); ~0 = none
    unsigned long long n_used;       // good alignments
    unsigned long long out_len;      // polished bases
    unsigned long long fix_count;    // (alignment, flagged tile) pairs found by k_collect
    unsigned int node_count;         // other-allele nodes allocated
    unsigned int flags;
    unsigned int n_fix_tiles, pad1, pad2, pad3;   // flagged tiles that have a run in the sorted list (k_fix_runs)
};

struct DevParams {                   // pp_polish_params, device resident (refreshed by a memcpy before each call)
    double fv, fi;
    uint32_t min_depth, max_errors;
    int careful;
    int pad;
};

// One distinct other allele at one position (the reference's HashMap<String,u32> entry, pileup.rs:40,62).
struct OthNode {
    unsigned long long sig;          // allele signature (see make_sig)
    unsigned long long val;          // where to read the allele: aln << 32 | start << 16 | len
    uint32_t count;
    uint32_t next;                   // next node of the same position, NONE32 = end
};

struct DevData {                     // everything the kernels read, by value
    // alignments
    unsigned long long n_aln;
    const uint32_t *contig, *ref_start, *read_id, *seq_off, *cigar_off, *nm, *cigar_ops;
    const uint16_t *seq_len, *n_cigar;
    const uint8_t *flags, *seq_pool;
    // assembly
    const uint8_t* draft;            // ASCII
    const unsigned long long* contig_off;
    uint32_t n_contigs;
    uint32_t G;                      // total positions
    uint32_t n_tiles;
    // work
    uint32_t* k;                     // [n_reads] good alignments per read (only in the global-k fallback mode)
    unsigned long long* draft_nib;   // 4-bit draft codes, 16 per word
    unsigned long long* diff;        // [G+1] lo32 cover, hi32 covering alignments with k != 1
    unsigned long long* ex;          // [G] explicit A,C,G,T counts, 16 bits each
    uint32_t* delother;              // [G] lo16 "-" count, hi16 other-allele entry count
    uint32_t* oth_head;              // [G] 1 + index of the first OthNode of the position, 0 = none
    OthNode* nodes;
    uint32_t node_cap;
    uint32_t* tileflag;              // bitmap, tiles that hold k != 1 coverage
    unsigned long long* rec_gn;      // [n_aln] gstart << 32 | kept entries (0 = contributes nothing)
    uint32_t* rec_k;                 // [n_aln] k of the alignment's read group
    double* depth_fix;               // [n_tiles * 128] ordered depth, valid for flagged tiles
    uint32_t *fix_key, *fix_val;     // (tile + 1, alignment) pairs of flagged tiles; 0 keys = unused slots
    uint32_t fix_cap;
    const DevParams* prm;
    DevStatus* st;
};

// ------------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void report_error(DevStatus* st, unsigned long long aln, unsigned code) {
    atomicMin(&st->err, (aln << 8) | code);
}

__constant__ uint8_t c_comp[256];      // misc.rs:170-182 complement_base on upper-cased bytes
__constant__ char c_nib2asc[16] = {'=', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};

__device__ __forceinline__ uint32_t brev4(uint32_t c) {   // complement of a BAM nibble = 4-bit reversal
    return __brev(c) >> 28;
}

// Sequence access policies.  sym = 4-bit code (SEQ4) or upper-cased ASCII byte (SEQ8).
template <int BITS> struct Seq;
template <> struct Seq<4> {
    static __device__ __forceinline__ uint32_t read_sym(const uint8_t* pool, uint32_t off_blk, uint32_t len, bool rc, uint32_t i) {
        uint32_t j = rc ? (len - 1 - i) : i;
        uint32_t b = pool[(size_t)off_blk * (PP_SEQ_BLOCK / 2) + (j >> 1)];
        uint32_t c = (b >> ((j & 1) * 4)) & 15u;
        return rc ? brev4(c) : c;
    }
    static __device__ __forceinline__ uint32_t draft_sym(const DevData& d, uint32_t pos) {
        return (uint32_t)(d.draft_nib[pos >> 4] >> ((pos & 15) * 4)) & 15u;
    }
    static __device__ __forceinline__ bool is_del(uint32_t) { return false; }
    static __device__ __forceinline__ int acgt(uint32_t s) { return s == 1 ? 0 : s == 2 ? 1 : s == 4 ? 2 : s == 8 ? 3 : -1; }
    static __device__ __forceinline__ uint8_t ascii(uint32_t s) { return (uint8_t)c_nib2asc[s & 15]; }
};
template <> struct Seq<8> {
    static __device__ __forceinline__ uint32_t read_sym(const uint8_t* pool, uint32_t off_blk, uint32_t len, bool rc, uint32_t i) {
        uint32_t j = rc ? (len - 1 - i) : i;
        uint32_t b = pool[(size_t)off_blk * PP_SEQ_BLOCK + j];
        return rc ? c_comp[b] : b;
    }
    static __device__ __forceinline__ uint32_t draft_sym(const DevData& d, uint32_t pos) {
        uint32_t b = d.draft[pos];
        return b == '-' ? 0u : b;          // a '-' in the draft never "matches" a read base: both count as "-"
    }
    static __device__ __forceinline__ bool is_del(uint32_t s) { return s == '-'; }   // the 1-char string "-"
    static __device__ __forceinline__ int acgt(uint32_t s) { return s == 'A' ? 0 : s == 'C' ? 1 : s == 'G' ? 2 : s == 'T' ? 3 : -1; }
    static __device__ __forceinline__ uint8_t ascii(uint32_t s) { return (uint8_t)s; }
};

// Allele signature of an "other" allele (anything that is not a single A/C/G/T or "-"): the whole string for
// short alleles (4-bit: <= 15 bases, 8-bit: <= 7 bytes) with the length in the low field, else length field 0
// and a hash of the content (equality then falls back to comparing the sequences themselves).
template <int BITS> __device__ __forceinline__ bool sig_exact(unsigned long long sig) {
    return BITS == 4 ? (sig & 15ull) != 0 : (sig & 255ull) != 0;
}
template <int BITS>
__device__ __forceinline__ unsigned long long make_sig(const uint8_t* pool, uint32_t off_blk, uint32_t slen, bool rc,
                                                        uint32_t start, uint32_t len) {
    const uint32_t maxlen = BITS == 4 ? 15 : 7;
    if (len <= maxlen) {
        unsigned long long sig = len;
        for (uint32_t i = 0; i < len; ++i)
            sig |= (unsigned long long)Seq<BITS>::read_sym(pool, off_blk, slen, rc, start + i) << ((BITS == 4 ? 4 : 8) * (i + 1));
        return sig;
    }
    unsigned long long h = 0xcbf29ce484222325ull;
    for (uint32_t i = 0; i < len; ++i) { h ^= Seq<BITS>::read_sym(pool, off_blk, slen, rc, start + i); h *= 0x100000001b3ull; }
    h ^= len;
    return h << (BITS == 4 ? 4 : 8);
}
__device__ __forceinline__ uint32_t mix32(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)x;
}

// ------------------------------------------------------------------------------------------------------
// k_draft_nib: ASCII draft -> 4-bit codes (0 = not one of the 15 letters: never equals a read code)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t asc2nib(uint32_t c) {
    switch (c) {
        case 'A': return 1; case 'C': return 2; case 'M': return 3; case 'G': return 4; case 'R': return 5;
        case 'S': return 6; case 'V': return 7; case 'T': return 8; case 'W': return 9; case 'Y': return 10;
        case 'H': return 11; case 'K': return 12; case 'D': return 13; case 'B': return 14; case 'N': return 15;
        default: return 0;
    }
}

__global__ void __launch_bounds__(256) k_draft_nib(const uint8_t* __restrict__ draft, uint32_t G,
                                                   unsigned long long* __restrict__ nib, uint32_t n_words) {
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        size_t base = (size_t)w * 16;
        unsigned long long v = 0;
        if (base + 16 <= G) {
            uint4 q = *reinterpret_cast<const uint4*>(draft + base);   // draft is 16 B aligned
            uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) v |= (unsigned long long)asc2nib((ws[i >> 2] >> ((i & 3) * 8)) & 255u) << (4 * i);
        } else {
            for (int i = 0; i < 16; ++i)
                if (base + i < G) v |= (unsigned long long)asc2nib(draft[base + i]) << (4 * i);
        }
        nib[w] = v;
    }
}

// ------------------------------------------------------------------------------------------------------
// Goodness (alignment.rs:283-287) and --careful (:277-279) of one alignment.  `multi` = its read group has more
// than one aligned record.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool group_is_multi(const DevData& d, unsigned long long a, uint32_t rid) {
    return (a > 0 && d.read_id[a - 1] == rid) || (a + 1 < d.n_aln && d.read_id[a + 1] == rid);
}
__device__ __forceinline__ bool alignment_is_good(const DevData& d, unsigned long long a, bool multi, uint32_t co, uint32_t nc, uint8_t fl) {
    if (nc == 0) { report_error(d.st, a, ERR_BAD_OP); return false; }       // the packer never emits this
    const uint32_t f = d.cigar_ops[co] & 15u, l = d.cigar_ops[co + nc - 1] & 15u;
    return (f == PP_OP_M || f == PP_OP_EQ) && (l == PP_OP_M || l == PP_OP_EQ) && d.nm[a] <= d.prm->max_errors &&
           !(fl & PP_FLAG_ZPFAIL) && !(d.prm->careful && multi);
}

// k_classify_multi: FALLBACK pre-pass, only launched when a read group was too large for k_scatter's in-kernel scan
// (FL_BIGGROUP): k = #good of every multi-record group into a global array (alignment.rs:288).
__global__ void __launch_bounds__(256) k_classify_multi(DevData d) {
    for (unsigned long long a = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; a < d.n_aln;
         a += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t rid = d.read_id[a];
        if (!group_is_multi(d, a, rid)) continue;
        if (alignment_is_good(d, a, true, d.cigar_off[a], d.n_cigar[a], d.flags[a])) atomicAdd(&d.k[rid], 1u);
    }
}

// ------------------------------------------------------------------------------------------------------
// Other alleles: find-or-insert into the per-position chain.  Nodes are written completely, fenced, then linked
// with a CAS on the head; readers load head and node fields through L2 (ld.cg), so a linked node is always whole.
// ------------------------------------------------------------------------------------------------------
struct SeqRef {                      // where an alignment's bases live, for comparing long alleles
    const uint8_t* seq_pool;
    const uint32_t* seq_off;
    const uint16_t* seq_len;
    const uint8_t* flags;
};

template <int BITS>
__device__ __forceinline__ uint32_t allele_sym(const SeqRef& r, unsigned long long val, uint32_t t) {
    const uint32_t aln = (uint32_t)(val >> 32), start = (uint32_t)(val >> 16) & 0xFFFFu;
    return Seq<BITS>::read_sym(r.seq_pool, r.seq_off[aln], r.seq_len[aln], r.flags[aln] & PP_FLAG_RC, start + t);
}
template <int BITS>
__device__ bool allele_equal(const SeqRef& r, unsigned long long va, unsigned long long vb) {
    const uint32_t la = (uint32_t)va & 0xFFFFu, lb = (uint32_t)vb & 0xFFFFu;
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; ++i)
        if (allele_sym<BITS>(r, va, i) != allele_sym<BITS>(r, vb, i)) return false;
    return true;
}

template <int BITS>
__device__ void other_insert(const DevData& d, uint32_t pos, unsigned long long val, unsigned long long sig) {
    atomicAdd(&d.delother[pos], 1u << 16);                     // entries (not distinct alleles): feeds `matched`
    const SeqRef sr{d.seq_pool, d.seq_off, d.seq_len, d.flags};
    uint32_t mine = NONE32;
    uint32_t h = __ldcg(&d.oth_head[pos]);                      // 1 + node index, 0 = empty
    uint32_t stop = 0;
    for (;;) {
        for (uint32_t n = h; n != stop;) {
            const OthNode* nd = &d.nodes[n - 1];
            const unsigned long long nsig = __ldcg(&nd->sig);
            if (nsig == sig && (sig_exact<BITS>(sig) || allele_equal<BITS>(sr, __ldcg(&nd->val), val))) {
                atomicAdd(&d.nodes[n - 1].count, 1u);
                return;                                         // (a node allocated on an earlier round stays unlinked)
            }
            const uint32_t nx = __ldcg(&nd->next);
            n = (nx == NONE32) ? 0 : nx + 1;
        }
        if (mine == NONE32) {
            mine = atomicAdd(&d.st->node_count, 1u);
            if (mine >= d.node_cap) { atomicOr(&d.st->flags, FL_NODE_OVF); return; }
            d.nodes[mine].sig = sig;
            d.nodes[mine].val = val;
            d.nodes[mine].count = 1;
        }
        d.nodes[mine].next = (h == 0) ? NONE32 : h - 1;
        __threadfence();
        const uint32_t old = atomicCAS(&d.oth_head[pos], h, mine + 1);
        if (old == h) return;
        stop = h;                                               // someone linked new nodes in front: look only at those
        h = old;
    }
}

// ------------------------------------------------------------------------------------------------------
// k_scatter
// ------------------------------------------------------------------------------------------------------
struct ScatterShared {
    uint32_t rid[SC_THREADS];
    uint8_t good[SC_THREADS];
    uint32_t n_good;
    uint32_t seq_lo, seq_hi;                      // byte range of the sequence pool used by this block's good alignments
    __align__(16) uint8_t seq[SC_SEQ_BYTES + 64]; // staged [seq_lo, seq_lo + SC_SEQ_BYTES) (+ slack for the 3-word window reads)
};

template <int BITS> struct Scatter {
    const DevData& d;

    __device__ __forceinline__ void push_other(uint32_t pos, unsigned long long aln, uint32_t start, uint32_t len, unsigned long long sig) {
        other_insert<BITS>(d, pos, (aln << 32) | ((unsigned long long)(start & 0xFFFFu) << 16) | (len & 0xFFFFu), sig);
    }
    // one single-base entry at reference position pos carrying read symbol s (read index ri)
    __device__ __forceinline__ void count_base(uint32_t pos, uint32_t s, unsigned long long aln, uint32_t ri) {
        if (Seq<BITS>::is_del(s)) { atomicAdd(&d.delother[pos], 1u); return; }
        uint32_t ds = Seq<BITS>::draft_sym(d, pos);
        if (s == ds) return;                                   // counted implicitly: cover - explicit
        int c = Seq<BITS>::acgt(s);
        if (c >= 0) atomicAdd(&d.ex[pos], 1ull << (16 * c));
        else push_other(pos, aln, ri, 1, 1ull | ((unsigned long long)s << (BITS == 4 ? 4 : 8)));
    }
    // interval add for an alignment that keeps entries [gstart, gstart + nkept); alignments of reads with k != 1 also
    // mark the 128-position tiles they cover (those get the ordered depth sum)
    __device__ __forceinline__ void add_interval(uint32_t gstart, uint32_t nkept, bool multi) {
        if (nkept == 0) return;
        const unsigned long long v = 1ull | (multi ? (1ull << 32) : 0ull);
        atomicAdd(&d.diff[gstart], v);
        atomicAdd(&d.diff[gstart + nkept], 0ull - v);
        if (multi) {
            const uint32_t t0 = gstart >> PP_TILE_SHIFT, t1 = (gstart + nkept - 1) >> PP_TILE_SHIFT;
            for (uint32_t t = t0; t <= t1; ++t) atomicOr(&d.tileflag[t >> 5], 1u << (t & 31));
        }
    }
    // 4-bit only: `vc` (<= 32) single-base entries whose read codes are the low nibbles of r0:r1, at reference
    // positions pos0.. ; ri0 = read index of the first one.  One explicit count per base that differs from the draft.
    __device__ __forceinline__ void scan_mismatches(unsigned long long r0, unsigned long long r1, uint32_t vc, uint32_t pos0,
                                                    unsigned long long aln, uint32_t ri0) {
        unsigned long long d0, d1;
        load_nib32(d.draft_nib, pos0, d0, d1);
        unsigned long long m0, m1;
        mismatch_masks(r0, r1, d0, d1, vc, m0, m1);
        while (m0) {
            const uint32_t j = (uint32_t)(__ffsll((long long)m0) - 1) >> 2;
            m0 &= m0 - 1;
            const uint32_t code = (uint32_t)(r0 >> (4 * j)) & 15u;
            const int c = Seq<4>::acgt(code);
            if (c >= 0) atomicAdd(&d.ex[pos0 + j], 1ull << (16 * c));
            else push_other(pos0 + j, aln, ri0 + j, 1, 1ull | ((unsigned long long)code << 4));
        }
        while (m1) {
            const uint32_t j = (uint32_t)(__ffsll((long long)m1) - 1) >> 2;
            m1 &= m1 - 1;
            const uint32_t code = (uint32_t)(r1 >> (4 * j)) & 15u;
            const int c = Seq<4>::acgt(code);
            if (c >= 0) atomicAdd(&d.ex[pos0 + 16 + j], 1ull << (16 * c));
            else push_other(pos0 + 16 + j, aln, ri0 + 16 + j, 1, 1ull | ((unsigned long long)code << 4));
        }
    }
};

// One THREAD per alignment (consecutive alignments on consecutive threads): every per-alignment array is read
// coalesced, there is no intra-warp cooperation to fall out of step, and simple and indel-bearing alignments run the
// same code with slightly different trip counts.  The block's slice of the 4-bit sequence pool is staged in shared
// memory by a coalesced cooperative copy; the draft is compared 32 bases at a time on 128-bit words.
// GLOBALK = false: k of a multi-record group is counted right here (its records are consecutive alignments);
// GLOBALK = true : k comes from k_classify_multi (fallback for huge groups).
template <int BITS, bool GLOBALK>
__global__ void __launch_bounds__(SC_THREADS) k_scatter(DevData d) {
    __shared__ ScatterShared sh;
    Scatter<BITS> S{d};
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const unsigned long long n_blocks = (d.n_aln + SC_THREADS - 1) / SC_THREADS;
    unsigned long long used = 0;

    for (unsigned long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        if (tid == 0) { sh.seq_lo = 0xFFFFFFFFu; sh.seq_hi = 0; }
        // ---- stage 1: metadata + goodness (alignment.rs:283-287), one alignment per thread
        const unsigned long long blk0 = blk * SC_THREADS;
        const unsigned long long aln = blk0 + tid;
        const uint32_t nvalid = (uint32_t)min((unsigned long long)SC_THREADS, d.n_aln - blk0);
        bool good = false, grp = false;
        uint32_t rid = 0, cigoff = 0, ncig = 0;
        uint8_t fl = 0;
        if (aln < d.n_aln) {
            rid = d.read_id[aln];
            grp = group_is_multi(d, aln, rid);
            cigoff = d.cigar_off[aln];
            ncig = d.n_cigar[aln];
            fl = d.flags[aln];
            good = alignment_is_good(d, aln, grp, cigoff, ncig, fl);
        }
        sh.rid[tid] = rid;
        sh.good[tid] = good ? 1 : 0;
        __syncthreads();
        // k = number of good alignments of the read group (alignment.rs:288)
        uint32_t k = 1;
        if (good && grp) {
            if (GLOBALK) k = d.k[rid];
            else {
                uint32_t count = 1, steps = 0;
                int i = (int)tid;
                while (i > 0 && sh.rid[i - 1] == rid) { --i; count += sh.good[i]; }
                if (i == 0) {
                    for (unsigned long long a2 = blk0; a2 > 0 && d.read_id[a2 - 1] == rid;) {
                        --a2;
                        count += alignment_is_good(d, a2, true, d.cigar_off[a2], d.n_cigar[a2], d.flags[a2]) ? 1 : 0;
                        if (++steps > SC_GROUP_SCAN_LIMIT) { atomicOr(&d.st->flags, FL_BIGGROUP); break; }
                    }
                }
                i = (int)tid;
                while (i + 1 < (int)nvalid && sh.rid[i + 1] == rid) { ++i; count += sh.good[i]; }
                if (i == (int)nvalid - 1) {
                    for (unsigned long long a2 = blk0 + nvalid - 1; a2 + 1 < d.n_aln && d.read_id[a2 + 1] == rid;) {
                        ++a2;
                        count += alignment_is_good(d, a2, true, d.cigar_off[a2], d.n_cigar[a2], d.flags[a2]) ? 1 : 0;
                        if (++steps > SC_GROUP_SCAN_LIMIT) { atomicOr(&d.st->flags, FL_BIGGROUP); break; }
                    }
                }
                k = count;
            }
        }
        bool rc = false;
        uint32_t gstart = 0, cend = 0, seqoff = 0, len = 0;
        if (fl & PP_FLAG_GHOST) good = false;                   // another shard scatters it; it only counted towards k
        if (good) {
            used++;
            good = false;
            const uint32_t c = d.contig[aln];
            if (c == PP_CONTIG_UNKNOWN) report_error(d.st, aln, ERR_UNKNOWN_CONTIG);
            else if (fl & PP_FLAG_NOSEQ) report_error(d.st, aln, ERR_NOSEQ);
            else {
                const unsigned long long gs = d.contig_off[c] + d.ref_start[aln];
                const unsigned long long ce = d.contig_off[c + 1];
                if (gs >= ce) report_error(d.st, aln, ERR_OOB);
                else {
                    good = true;
                    gstart = (uint32_t)gs; cend = (uint32_t)ce;
                    seqoff = d.seq_off[aln]; len = d.seq_len[aln];
                    rc = fl & PP_FLAG_RC;
                    if (BITS == 4) {
                        const unsigned long long b0 = (unsigned long long)seqoff * 16, b1 = b0 + (((unsigned long long)len + 31) / 32) * 16;
                        if (b1 < 0xFFFFFFFFull) { atomicMin(&sh.seq_lo, (uint32_t)b0); atomicMax(&sh.seq_hi, (uint32_t)b1); }
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.draft_nib + (gs >> 4)));
                    }
                }
            }
        }
        __syncthreads();
        // ---- stage 1b: coalesced copy of the block's sequence bytes into shared memory
        const uint32_t seq_lo = sh.seq_lo;
        const uint32_t seq_hi = (sh.seq_hi > seq_lo) ? (uint32_t)min((unsigned long long)sh.seq_hi, (unsigned long long)seq_lo + SC_SEQ_BYTES) : seq_lo;
        if (BITS == 4) {
            for (unsigned long long off = (unsigned long long)seq_lo + tid * 16; off < seq_hi; off += SC_THREADS * 16)
                *reinterpret_cast<uint4*>(sh.seq + (off - seq_lo)) = __ldg(reinterpret_cast<const uint4*>(d.seq_pool + off));
            __syncthreads();
        }

        // ---- stage 2: the CIGAR walk of this thread's alignment (alignment.rs:175-201, 364-378; pileup.rs:189-200)
        uint32_t nkept = 0;
        if (good) do {
            const uint8_t* seqp = d.seq_pool + (size_t)seqoff * (BITS == 4 ? 16 : 32);
            if (BITS == 4) {
                const unsigned long long b0 = (unsigned long long)seqoff * 16, b1 = b0 + (((unsigned long long)len + 31) / 32) * 16;
                if (b0 >= seq_lo && b1 <= seq_hi) seqp = sh.seq + (b0 - seq_lo);
            }
            const uint32_t* ops = d.cigar_ops + cigoff;
            // pass 1: validate ops, E = entries, R = read bases consumed
            unsigned long long E = 0, R = 0;
            bool bad = false;
            for (uint32_t p = 0; p < ncig; ++p) {
                const uint32_t op = ops[p], o = op & 15u, l = op >> 4;
                if (o == PP_OP_M || o == PP_OP_EQ || o == PP_OP_X) { E += l; R += l; }
                else if (o == PP_OP_I) R += l;
                else if (o == PP_OP_D) E += l;
                else bad = true;                                   // alignment.rs:187-193
            }
            if (bad) { report_error(d.st, aln, ERR_BAD_OP); break; }
            if (R != len) { report_error(d.st, aln, ERR_SEQ_MISMATCH); break; }   // :195-198
            // pass 2: trim.  Walk entries from the right; stop at the first entry that is not the single base `last`.
            const uint32_t last = Seq<BITS>::read_sym(seqp, 0, len, rc, len - 1);
            unsigned long long run = 0;
            {
                uint32_t ri = len;            // read index just past the current entry's M-part
                uint32_t pend = 0;            // inserted bases that extend the entry being visited
                bool stop = false;
                for (int p = (int)ncig - 1; p >= 0 && !stop; --p) {
                    const uint32_t op = ops[p], o = op & 15u, l = op >> 4;
                    if (o == PP_OP_I) { pend += l; ri -= l; continue; }
                    if (o == PP_OP_D) {
                        // entries (ri, ri + pend): only the rightmost can carry pend; equal to `last` iff pend == 1 and base == last
                        for (uint32_t t = 0; t < l; ++t) {
                            if (pend == 1 && Seq<BITS>::read_sym(seqp, 0, len, rc, ri) == last) { run++; pend = 0; }
                            else { stop = true; break; }
                        }
                        continue;
                    }
                    for (uint32_t t = 0; t < l; ++t) {                // M / = / X
                        if (pend == 0 && Seq<BITS>::read_sym(seqp, 0, len, rc, ri - 1) == last) { run++; ri--; }
                        else { stop = true; break; }
                    }
                }
            }
            const unsigned long long nk64 = (E - run >= 1) ? (E - run - 1) : 0;
            if ((unsigned long long)gstart + nk64 > cend) { report_error(d.st, aln, ERR_OOB); break; }
            nkept = (uint32_t)nk64;
            S.add_interval(gstart, nkept, k != 1);
            // pass 3: emit entries e < nkept
            uint32_t e = 0, ri = 0;
            for (uint32_t p = 0; p < ncig && e < nkept; ++p) {
                const uint32_t op = ops[p], o = op & 15u, l = op >> 4;
                if (o == PP_OP_I) { ri += l; continue; }
                uint32_t ins = 0;                                     // inserted bases right after this op
                for (uint32_t q = p + 1; q < ncig && (ops[q] & 15u) == PP_OP_I; ++q) ins += ops[q] >> 4;
                if (o == PP_OP_D) {
                    const uint32_t plain = min(ins ? l - 1 : l, nkept - e);   // the last "-" entry absorbs a following insertion
                    for (uint32_t t = 0; t < plain; ++t) atomicAdd(&d.delother[gstart + e + t], 1u);
                    if (ins && e + l - 1 < nkept) {
                        const uint32_t pos = gstart + e + l - 1;
                        if (ins == 1) S.count_base(pos, Seq<BITS>::read_sym(seqp, 0, len, rc, ri), aln, ri);
                        else S.push_other(pos, aln, ri, ins, make_sig<BITS>(seqp, 0, len, rc, ri, ins));
                    }
                    e += l;
                    continue;
                }
                const uint32_t plain = min(ins ? l - 1 : l, nkept - e);   // single-base entries of this M / = / X run that are kept
                if (BITS == 4) {
                    for (uint32_t c0 = 0; c0 < plain; c0 += 32) {
                        unsigned long long r0, r1;
                        load_read32(reinterpret_cast<const unsigned long long*>(seqp), len, rc, ri + c0, r0, r1);
                        S.scan_mismatches(r0, r1, min(plain - c0, 32u), gstart + e + c0, aln, ri + c0);
                    }
                } else {
                    for (uint32_t t = 0; t < plain; ++t)
                        S.count_base(gstart + e + t, Seq<BITS>::read_sym(seqp, 0, len, rc, ri + t), aln, ri + t);
                }
                if (ins && e + l - 1 < nkept)
                    S.push_other(gstart + e + l - 1, aln, ri + l - 1, 1 + ins, make_sig<BITS>(seqp, 0, len, rc, ri + l - 1, 1 + ins));
                e += l;
                ri += l;
            }
        } while (false);
        // what k_collect / k_depth_fixup need to know about this alignment
        if (aln < d.n_aln) {
            d.rec_gn[aln] = nkept ? (((unsigned long long)gstart << 32) | nkept) : 0ull;
            d.rec_k[aln] = k;
        }
        __syncthreads();
    }
    // good alignments (alignment.rs:304): block reduce, one atomic per CTA
    if (tid == 0) sh.n_good = 0;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) used += __shfl_down_sync(0xffffffffu, used, o);
    if (lane == 0 && used) atomicAdd(&sh.n_good, (uint32_t)used);
    __syncthreads();
    if (tid == 0 && sh.n_good) atomicAdd(&d.st->n_used, (unsigned long long)sh.n_good);
}

// block-wide exclusive scan of one u64 per thread (VT_THREADS threads); returns exclusive prefix, total in *total
__device__ __forceinline__ unsigned long long block_exscan(unsigned long long v, unsigned long long* s_warp, unsigned long long* total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = (lane < VT_THREADS / 32) ? s_warp[lane] : 0;
        unsigned long long winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if ((int)lane >= o) winc += t;
        }
        if (lane < VT_THREADS / 32) s_warp[lane] = winc - w;      // exclusive per-warp offsets
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    unsigned long long r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

// misc.rs:208-215 bankers_rounding on a non-negative finite double (depth * fraction)
__device__ __forceinline__ uint32_t bankers_rounding(double x) {
    uint32_t rd;
    if (!(x == x) || x <= 0.0) rd = 0;
    else if (x >= 4294967295.0) rd = 4294967295u;
    else rd = (uint32_t)x;                                  // truncation
    const double fr = __dsub_rn(x, trunc(x));
    if (fr < 0.5) return rd;
    if (fr > 0.5) return rd + 1;
    return rd + (rd & 1u);
}


// ------------------------------------------------------------------------------------------------------
// k_collect: every (alignment, flagged tile) pair, written IN ALIGNMENT ORDER by an order-preserving compaction
// (block scan + decoupled look-back).  A stable sort by tile then leaves each tile's list in SAM order, which is
// the order the reference adds depth contributions in (pileup.rs:64).
// ------------------------------------------------------------------------------------------------------
struct CollectParams { uint32_t n_chunks; unsigned long long* chunk_off; };   // chunk_off[c]: pairs before chunk c (after the scan)

// Pairs of one thread's CL_ITEMS consecutive alignments: count, or write them at o.
template <bool WRITE>
__device__ __forceinline__ unsigned long long collect_items(const DevData& d, unsigned long long a0, unsigned long long o) {
    unsigned long long gn[CL_ITEMS];
    {
        const ulonglong2* q = reinterpret_cast<const ulonglong2*>(d.rec_gn + a0);     // rec_gn is padded to whole chunks
#pragma unroll
        for (int i = 0; i < CL_ITEMS / 2; ++i) { const ulonglong2 v = q[i]; gn[2 * i] = v.x; gn[2 * i + 1] = v.y; }
    }
    unsigned long long cnt = 0;
#pragma unroll
    for (int i = 0; i < CL_ITEMS; ++i) {
        if (a0 + i >= d.n_aln) continue;
        const uint32_t nk = (uint32_t)gn[i];
        if (!nk) continue;
        const uint32_t gs = (uint32_t)(gn[i] >> 32);
        for (uint32_t t = gs >> PP_TILE_SHIFT; t <= ((gs + nk - 1) >> PP_TILE_SHIFT); ++t)
            if ((d.tileflag[t >> 5] >> (t & 31)) & 1u) {
                if (WRITE) {
                    if (o < d.fix_cap) { d.fix_key[o] = t + 1; d.fix_val[o] = (uint32_t)(a0 + i); }
                    o++;
                }
                cnt++;
            }
    }
    return cnt;
}

// Pass 1: pairs per chunk of CL_CHUNK alignments (then one small device scan).  Pass 2: the pairs, written in alignment order.
// (A single pass with a decoupled look-back spent most of its time waiting for the prefix wavefront: 79 us.)
__global__ void __launch_bounds__(CL_THREADS) k_collect_count(DevData d, CollectParams cp) {
    __shared__ unsigned long long s_warp[CL_THREADS / 32];
    __shared__ unsigned long long s_total;
    const uint32_t chunk = blockIdx.x, tid = threadIdx.x;
    const unsigned long long cnt = collect_items<false>(d, (unsigned long long)chunk * CL_CHUNK + tid * CL_ITEMS, 0);
    block_exscan(cnt, s_warp, &s_total);
    if (tid == 0) cp.chunk_off[chunk] = s_total;
}

__global__ void __launch_bounds__(CL_THREADS) k_collect(DevData d, CollectParams cp) {
    __shared__ unsigned long long s_warp[CL_THREADS / 32];
    __shared__ unsigned long long s_total;
    const uint32_t chunk = blockIdx.x, tid = threadIdx.x;
    const unsigned long long a0 = (unsigned long long)chunk * CL_CHUNK + tid * CL_ITEMS;
    const unsigned long long cnt = collect_items<false>(d, a0, 0);
    const unsigned long long excl = block_exscan(cnt, s_warp, &s_total);
    const unsigned long long o = cp.chunk_off[chunk] + excl;
    if (chunk == cp.n_chunks - 1 && tid == CL_THREADS - 1) {
        d.st->fix_count = o + cnt;
        if (o + cnt > d.fix_cap) atomicOr(&d.st->flags, FL_FIX_OVF);
    }
    if (cnt) collect_items<true>(d, a0, o);
}

// ------------------------------------------------------------------------------------------------------
// k_depth_fixup: the reference's sequential f64 depth sum (pileup.rs:64, alignment.rs:288) for every
// position of a flagged tile, in SAM order.  keys = tile + 1 sorted ascending (unused slots are 0 and sort first).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

#define FX_WARPS 4
#define FX_BATCH 128                 // list entries staged per round (4 per lane, their gathers in flight together)

// Where each flagged tile's run starts and ends in the sorted list, and the list of tiles that have one: one pass over the
// sorted keys instead of a ticket per tile and two binary searches per flagged tile.
__global__ void __launch_bounds__(256) k_fix_runs(DevData d, const uint32_t* __restrict__ keys, uint32_t* __restrict__ run_lo, uint32_t* __restrict__ run_hi,
                                                  uint32_t* __restrict__ tiles) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.fix_cap) return;
    const uint32_t k = keys[i];
    if (!k) return;
    if (i == 0 || keys[i - 1] != k) {
        run_lo[k - 1] = i;
        tiles[atomicAdd(&d.st->n_fix_tiles, 1u)] = k - 1;
    }
    if (i + 1 == d.fix_cap || keys[i + 1] != k) run_hi[k - 1] = i + 1;
}

__global__ void __launch_bounds__(FX_WARPS * 32) k_depth_fixup(DevData d, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ run_lo,
                                                               const uint32_t* __restrict__ run_hi, const uint32_t* __restrict__ tiles) {
    // One WARP per flagged tile, four consecutive positions per lane (four independent dependent-add chains); the
    // tile's list is staged FX_BATCH entries at a time in the warp's own shared-memory slice.
    __shared__ uint2 s_rng[FX_WARPS][FX_BATCH];        // (start, length) of each staged entry
    __shared__ double s_inv[FX_WARPS][FX_BATCH];       // 1.0 / k
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n_fix = d.st->n_fix_tiles;
    for (uint32_t w = blockIdx.x * FX_WARPS + wib; w < n_fix; w += gridDim.x * FX_WARPS) {
        const uint32_t tile = tiles[w];
        const uint32_t lo = run_lo[tile], hi = run_hi[tile];
        const uint32_t p = tile * PP_TILE + lane * 4;
        double dep0 = 0.0, dep1 = 0.0, dep2 = 0.0, dep3 = 0.0;
        for (uint32_t base = lo; base < hi; base += FX_BATCH) {
            uint32_t aln[FX_BATCH / 32];
            unsigned long long gn[FX_BATCH / 32];
            uint32_t kk[FX_BATCH / 32];
#pragma unroll
            for (int q = 0; q < FX_BATCH / 32; ++q) { const uint32_t i = base + q * 32 + lane; aln[q] = (i < hi) ? vals[i] : 0u; }
#pragma unroll
            for (int q = 0; q < FX_BATCH / 32; ++q) { const uint32_t i = base + q * 32 + lane; gn[q] = (i < hi) ? d.rec_gn[aln[q]] : 0ull; kk[q] = (i < hi) ? d.rec_k[aln[q]] : 1u; }
#pragma unroll
            for (int q = 0; q < FX_BATCH / 32; ++q) {
                s_rng[wib][q * 32 + lane] = make_uint2((uint32_t)(gn[q] >> 32), (uint32_t)gn[q]);
                s_inv[wib][q * 32 + lane] = __ddiv_rn(1.0, (double)kk[q]);      // 1.0 / good_alignments.len() as f64
            }
            __syncwarp();
            const uint32_t cnt = min((uint32_t)FX_BATCH, hi - base);
#pragma unroll 4
            for (uint32_t j = 0; j < cnt; ++j) {
                const uint2 r = s_rng[wib][j];
                const double inv = s_inv[wib][j];
                const uint32_t off = p - r.x;                       // position p + q is covered iff (off + q) < length (unsigned)
                if (off < r.y) dep0 = __dadd_rn(dep0, inv);
                if (off + 1u < r.y) dep1 = __dadd_rn(dep1, inv);
                if (off + 2u < r.y) dep2 = __dadd_rn(dep2, inv);
                if (off + 3u < r.y) dep3 = __dadd_rn(dep3, inv);
            }
            __syncwarp();
        }
        *reinterpret_cast<double4*>(d.depth_fix + p) = make_double4(dep0, dep1, dep2, dep3);
    }
}

// ------------------------------------------------------------------------------------------------------
// k_vote: prefix sum of the difference array (decoupled look-back) and the per-position vote.
// ------------------------------------------------------------------------------------------------------
struct VoteParams {
    uint32_t n_chunks;
    uint8_t* out;
    unsigned long long out_cap;
    unsigned long long* out_off;     // [n_contigs+1]
    unsigned long long *changed, *zero_depth;   // [n_contigs]
    double* total_depth;             // [n_contigs] sum of the per-position depths (polish.rs:177; the log's mean read depth)
    const unsigned long long* chunk_pre;   // [n_chunks] sum of diff[] before the chunk (k_diff_sums + device scan, side stream)
    // per-position verdicts handed from k_vote to k_compact
    uint16_t* res;                    // [padG] low byte = character, high byte = output length (255: see rec_at)
    uint32_t* rec_at;                 // [G] other-allele node to emit at a position (only where res says so)
    long long* chunk_delta;           // [n_chunks] sum(output length) - positions of the chunk
    pp_debug_pos* dbg;                // [G] per-position debug records, or nullptr
};

// What the other-allele slow path needs, passed by value so that the kernel parameter structs are never
// spilled to local memory for a call.
struct OthCtx {
    const OthNode* nodes;
    const uint32_t* head;
    SeqRef sr;
};

struct Tally { uint32_t nvalid, ninter; int which; uint32_t rec; };   // which: 0..3 ACGT, 4 "-", 5 draft's own non-ACGT base, 6 other node

__device__ __forceinline__ void tally(Tally& t, uint32_t c, uint32_t vt, uint32_t it, int which, uint32_t rec) {
    if (c >= vt) { t.nvalid++; t.which = which; t.rec = rec; }
    else if (c >= it) t.ninter++;
}

// Other alleles at `pos` (pileup.rs:102-109): one chain node per distinct allele, count already exact.
__device__ __noinline__ Tally tally_others(OthCtx oc, uint32_t pos, uint32_t vt, uint32_t it, Tally t) {
    for (uint32_t n = oc.head[pos]; n != 0;) {
        const OthNode& nd = oc.nodes[n - 1];
        tally(t, nd.count, vt, it, 6, n - 1);
        n = (nd.next == NONE32) ? 0 : nd.next + 1;
    }
    return t;
}

// Character t of other-allele node `rec` (from the exact signature when there is one, else from the read).
template <int BITS>
__device__ __forceinline__ uint8_t other_char(const OthCtx& oc, uint32_t rec, uint32_t t) {
    const unsigned long long sig = oc.nodes[rec].sig;
    if (sig_exact<BITS>(sig)) return Seq<BITS>::ascii((uint32_t)(sig >> ((BITS == 4 ? 4 : 8) * (t + 1))) & (BITS == 4 ? 15u : 255u));
    return Seq<BITS>::ascii(allele_sym<BITS>(oc.sr, oc.nodes[rec].val, t));
}

// Result of one position, packed: bits 0..15 output length, 16..23 output char (when length is 1 and not from a
// multi-base node), bit 24 changed, bit 25 emit from node `rec`.
struct PosOut { uint32_t packed; uint32_t rec; };

// The vote of pileup.rs:67-134 for one covered position.  packed bits 26..28 carry the BaseStatus.
template <int BITS>
__device__ __forceinline__ PosOut vote_position(const OthCtx& oc, const DevParams& prm, uint32_t pos, uint32_t orig, double depth,
                                                uint32_t cA, uint32_t cC, uint32_t cG, uint32_t cT, uint32_t cDel,
                                                uint32_t matched, uint32_t n_other, pp_debug_pos* dbg) {
    const uint32_t vt = max(prm.min_depth, bankers_rounding(__dmul_rn(depth, prm.fv)));
    const uint32_t it = bankers_rounding(__dmul_rn(depth, prm.fi));
    Tally t{0, 0, -1, 0};
    tally(t, cA, vt, it, 0, 0);
    tally(t, cC, vt, it, 1, 0);
    tally(t, cG, vt, it, 2, 0);
    tally(t, cT, vt, it, 3, 0);
    if (cDel) tally(t, cDel, vt, it, 4, 0);                  // "-" exists only if it was seen (a HashMap entry)
    if (matched) tally(t, matched, vt, it, 5, 0);            // the 1-char string of a non-ACGT draft base
    if (n_other) t = tally_others(oc, pos, vt, it, t);
    PosOut o;
    o.rec = 0;
    o.packed = (orig == '-' ? 0u : 1u) | (orig << 16);
    uint32_t status;                                          // 0 low_depth 1 none 2 multiple 3 too_close 4 kept 5 changed
    if (depth < (double)prm.min_depth) status = 0;
    else if (t.nvalid == 0) status = 1;
    else if (t.nvalid > 1) status = 2;
    else if (t.ninter > 0) status = 3;
    else {
        status = 4;
        if (t.which <= 3) {
            const uint32_t nb = (uint32_t)"ACGT"[t.which];
            if (nb != orig) status = 5;
            o.packed = 1u | (nb << 16) | (nb != orig ? 1u << 24 : 0u);
        } else if (t.which == 4) {
            if (orig != '-') status = 5;
            o.packed = 0u | ((uint32_t)'-' << 16) | (orig != '-' ? 1u << 24 : 0u);
        } else if (t.which == 6) {
            const uint32_t len = (uint32_t)oc.nodes[t.rec].val & 0xFFFFu;
            uint32_t n = 0;
            for (uint32_t q = 0; q < len; ++q) n += other_char<BITS>(oc, t.rec, q) != '-';
            // an other-allele string never equals the draft's own 1-char string (those entries are "matched")
            status = 5;
            o.packed = (n & 0xFFFFu) | (1u << 24) | (1u << 25);
            o.rec = t.rec;
        }
    }
    o.packed |= status << 26;
    if (dbg) {
        dbg->depth = depth; dbg->valid_threshold = vt; dbg->invalid_threshold = it;
        dbg->count[0] = cA; dbg->count[1] = cC; dbg->count[2] = cG; dbg->count[3] = cT; dbg->count[4] = cDel; dbg->count[5] = matched;
        dbg->n_other = n_other;
        dbg->new_node = ((o.packed >> 25) & 1u) ? o.rec : 0xFFFFFFFFu;
        dbg->original = (uint8_t)orig; dbg->status = (uint8_t)status;
        dbg->new_char = ((o.packed >> 25) & 1u) ? 0 : (((o.packed & 0xFFFFu) == 0 && orig != '-') ? (uint8_t)'-' : (uint8_t)(o.packed >> 16));
    }
    return o;
}

// Sum of the difference array over each vote chunk.  Depends on the scatter only, so it runs (with a small device scan) on
// the side stream while the main stream does the fix-up; k_vote then starts from a known prefix instead of waiting for a
// look-back wavefront (26 % of its stall samples).
__global__ void __launch_bounds__(VT_THREADS) k_diff_sums(const unsigned long long* __restrict__ diff, unsigned long long* __restrict__ chunk_sum) {
    __shared__ unsigned long long s_warp[VT_THREADS / 32];
    __shared__ unsigned long long s_total;
    const uint32_t p0 = blockIdx.x * VT_CHUNK + threadIdx.x * VT_ITEMS;
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(diff + p0);
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < VT_ITEMS / 2; ++i) { const ulonglong2 v = q[i]; t += v.x + v.y; }
    block_exscan(t, s_warp, &s_total);
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = s_total;
}

template <int BITS>
__global__ void __launch_bounds__(VT_THREADS, 3) k_vote(DevData d, VoteParams vp) {
    __shared__ unsigned long long s_warp[VT_THREADS / 32];
    __shared__ unsigned long long s_total;
    __shared__ long long s_delta[VT_THREADS / 32];
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = blockIdx.x;
    const uint32_t p0 = chunk * VT_CHUNK + tid * VT_ITEMS;
    const DevParams prm = *d.prm;

    // ---- 1. difference array -> cover / multi (arrays are padded to a whole number of chunks)
    unsigned long long dv[VT_ITEMS];
    {
        const ulonglong2* q = reinterpret_cast<const ulonglong2*>(d.diff + p0);
#pragma unroll
        for (int i = 0; i < VT_ITEMS / 2; ++i) { const ulonglong2 v = q[i]; dv[2 * i] = v.x; dv[2 * i + 1] = v.y; }
    }
    unsigned long long tsum = 0;
#pragma unroll
    for (int i = 0; i < VT_ITEMS; ++i) { tsum += dv[i]; dv[i] = tsum; }     // thread-inclusive
    const unsigned long long texcl = block_exscan(tsum, s_warp, &s_total);
    const unsigned long long base = vp.chunk_pre[chunk] + texcl;
    unsigned long long exv[VT_ITEMS];
    uint32_t dlv[VT_ITEMS];
    {
        const ulonglong2* q = reinterpret_cast<const ulonglong2*>(d.ex + p0);
#pragma unroll
        for (int i = 0; i < VT_ITEMS / 2; ++i) { const ulonglong2 v = q[i]; exv[2 * i] = v.x; exv[2 * i + 1] = v.y; }
        const uint4* r = reinterpret_cast<const uint4*>(d.delother + p0);
#pragma unroll
        for (int i = 0; i < VT_ITEMS / 4; ++i) { const uint4 v = r[i]; dlv[4 * i] = v.x; dlv[4 * i + 1] = v.y; dlv[4 * i + 2] = v.z; dlv[4 * i + 3] = v.w; }
    }
    const uint2 dr = *reinterpret_cast<const uint2*>(d.draft + p0);

    // ---- 2. vote
    OthCtx oc;
    oc.nodes = d.nodes; oc.head = d.oth_head;
    oc.sr = SeqRef{d.seq_pool, d.seq_off, d.seq_len, d.flags};
    PosOut po[VT_ITEMS];
    unsigned long long tlen = 0;
    uint32_t n_changed = 0, n_zero = 0;
    double tdepth = 0.0;
    uint32_t ctg = 0;
    if (p0 < d.G) {
        uint32_t lo = 0, hi = d.n_contigs;           // largest c with contig_off[c] <= p0
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (d.contig_off[mid] <= p0) lo = mid; else hi = mid; }
        ctg = lo;
    }
    uint32_t next_start = (ctg + 1 < d.n_contigs) ? (uint32_t)d.contig_off[ctg + 1] : 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < VT_ITEMS; ++i) {
        const uint32_t p = p0 + i;
        po[i].packed = 0; po[i].rec = 0;
        if (p >= d.G) continue;
        while (p >= next_start) {
            if (n_changed) atomicAdd(&vp.changed[ctg], (unsigned long long)n_changed);
            if (n_zero) atomicAdd(&vp.zero_depth[ctg], (unsigned long long)n_zero);
            if (tdepth != 0.0) atomicAdd(&vp.total_depth[ctg], tdepth);
            n_changed = n_zero = 0;
            tdepth = 0.0;
            ctg++;
            next_start = (ctg + 1 < d.n_contigs) ? (uint32_t)d.contig_off[ctg + 1] : 0xFFFFFFFFu;
        }
        const unsigned long long pv = base + dv[i];
        const uint32_t cover = (uint32_t)pv, multi = (uint32_t)(pv >> 32);
        const uint32_t orig = ((i < 4 ? dr.x : dr.y) >> ((i & 3) * 8)) & 255u;
        if (cover == 0) {                                    // depth 0: always the original base
            n_zero++;
            po[i].packed = (orig == '-' ? 0u : 1u) | (orig << 16);
            tlen += po[i].packed & 0xFFFFu;
            if (vp.dbg) {                                    // min_depth > 0: low_depth; min_depth == 0: A,C,G,T all "valid" -> multiple
                pp_debug_pos r;
                memset(&r, 0, sizeof r);
                r.valid_threshold = prm.min_depth; r.new_node = 0xFFFFFFFFu; r.original = (uint8_t)orig;
                r.status = prm.min_depth > 0 ? 0 : 2; r.new_char = (uint8_t)orig;
                vp.dbg[p] = r;
            }
            continue;
        }
        if (cover >= 65536u) atomicOr(&d.st->flags, FL_COUNTER_OVF);
        const double depth = multi ? d.depth_fix[p] : (double)cover;
        tdepth += depth;
        const unsigned long long ex = exv[i];
        if (ex == 0 && dlv[i] == 0 && !vp.dbg) {
            // every covering entry equals the draft base: the only allele with a non-zero count is the draft's own, so
            // whatever the thresholds say (kept, too_close, low_depth, ...) the emitted base is the original
            po[i].packed = (orig == '-' ? 0u : 1u) | (orig << 16);
            tlen += po[i].packed & 0xFFFFu;
            continue;
        }
        uint32_t cA = (uint32_t)ex & 0xFFFFu, cC = (uint32_t)(ex >> 16) & 0xFFFFu, cG = (uint32_t)(ex >> 32) & 0xFFFFu,
                 cT = (uint32_t)(ex >> 48) & 0xFFFFu;
        const uint32_t cDel = dlv[i] & 0xFFFFu, n_other = dlv[i] >> 16;
        uint32_t matched = cover - (cA + cC + cG + cT + cDel + n_other);
        if (orig == 'A') { cA += matched; matched = 0; }
        else if (orig == 'C') { cC += matched; matched = 0; }
        else if (orig == 'G') { cG += matched; matched = 0; }
        else if (orig == 'T') { cT += matched; matched = 0; }
        po[i] = vote_position<BITS>(oc, prm, p, orig, depth, cA, cC, cG, cT, cDel, matched, n_other, vp.dbg ? vp.dbg + p : nullptr);
        n_changed += (po[i].packed >> 24) & 1u;
        tlen += po[i].packed & 0xFFFFu;
    }
    if (n_changed) atomicAdd(&vp.changed[ctg], (unsigned long long)n_changed);
    if (n_zero) atomicAdd(&vp.zero_depth[ctg], (unsigned long long)n_zero);
    if (tdepth != 0.0) atomicAdd(&vp.total_depth[ctg], tdepth);

    // ---- 3. hand the verdicts to k_compact: 2 bytes per position + this chunk's length delta
    {
        uint32_t w[VT_ITEMS / 2];
#pragma unroll
        for (int i = 0; i < VT_ITEMS; ++i) {
            const uint32_t len = po[i].packed & 0xFFFFu;
            uint32_t h;
            if ((po[i].packed >> 25) & 1u) { h = 255u << 8; if (p0 + i < d.G) vp.rec_at[p0 + i] = po[i].rec; }
            else h = (len << 8) | ((po[i].packed >> 16) & 255u);
            if (i & 1) w[i >> 1] |= h << 16; else w[i >> 1] = h;
        }
        *reinterpret_cast<uint4*>(vp.res + p0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    const uint32_t npos = (p0 < d.G) ? min((uint32_t)VT_ITEMS, d.G - p0) : 0u;
    long long delta = (long long)tlen - (long long)npos;
    for (int o = 16; o > 0; o >>= 1) delta += __shfl_down_sync(0xffffffffu, delta, o);
    if ((tid & 31) == 0) s_delta[tid >> 5] = delta;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int i = 0; i < VT_THREADS / 32; ++i) t += s_delta[i];
        vp.chunk_delta[chunk] = t;
    }
}

// ------------------------------------------------------------------------------------------------------
// k_compact: polish.rs:185-188 (push_str of every position's allele, then replace("-", "")).  Chunk c writes its
// characters at c * VT_CHUNK + sum(chunk_delta[0..c)); no inter-CTA dependency.
// ------------------------------------------------------------------------------------------------------
#define CP_STAGE (VT_CHUNK + 2048)
template <int BITS>
__global__ void __launch_bounds__(VT_THREADS) k_compact(DevData d, VoteParams vp) {
    __shared__ unsigned long long s_warp[VT_THREADS / 32];
    __shared__ unsigned long long s_total;
    __shared__ long long s_red[VT_THREADS / 32];
    __shared__ long long s_base;
    __shared__ __align__(16) uint8_t s_out[CP_STAGE];
    const uint32_t tid = threadIdx.x, chunk = blockIdx.x;
    const uint32_t p0 = chunk * VT_CHUNK + tid * VT_ITEMS;
    // base offset of this chunk
    long long acc = 0;
    for (uint32_t j = tid; j < chunk; j += VT_THREADS) acc += vp.chunk_delta[j];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int i = 0; i < VT_THREADS / 32; ++i) t += s_red[i];
        s_base = (long long)chunk * VT_CHUNK + t;
    }
    // verdicts
    const uint4 rv = *reinterpret_cast<const uint4*>(vp.res + p0);
    const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t len[VT_ITEMS];
    unsigned long long tlen = 0;
    OthCtx oc;
    oc.nodes = d.nodes; oc.head = d.oth_head;
    oc.sr = SeqRef{d.seq_pool, d.seq_off, d.seq_len, d.flags};
#pragma unroll
    for (int i = 0; i < VT_ITEMS; ++i) {
        const uint32_t h = (w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
        uint32_t l = h >> 8;
        if (p0 + i >= d.G) l = 0;
        else if (l == 255) {                                      // a multi-base allele: count its non-'-' characters
            const uint32_t rec = vp.rec_at[p0 + i];
            const uint32_t rlen = (uint32_t)oc.nodes[rec].val & 0xFFFFu;
            l = 0;
            for (uint32_t t = 0; t < rlen; ++t) l += other_char<BITS>(oc, rec, t) != '-';
        }
        len[i] = l;
        tlen += l;
    }
    const unsigned long long oexcl = block_exscan(tlen, s_warp, &s_total);
    const unsigned long long total = s_total;
    const unsigned long long base = (unsigned long long)s_base;
    if (chunk == vp.n_chunks - 1 && tid == 0) { vp.out_off[d.n_contigs] = base + total; d.st->out_len = base + total; }
    // out_off of contigs that start inside this thread's positions
    if (p0 < d.G) {
        uint32_t lo = 0, hi = d.n_contigs;           // first c with contig_off[c] >= p0
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (d.contig_off[mid] < p0) lo = mid + 1; else hi = mid; }
        if (lo < d.n_contigs && d.contig_off[lo] < (unsigned long long)p0 + VT_ITEMS) {
            unsigned long long oo = base + oexcl;
            uint32_t c = lo;
#pragma unroll
            for (int i = 0; i < VT_ITEMS; ++i) {
                while (c < d.n_contigs && d.contig_off[c] == p0 + i) { vp.out_off[c] = oo; c++; }
                oo += len[i];
            }
        }
    }
    if (base + total > vp.out_cap) { if (tid == 0) atomicOr(&d.st->flags, FL_OUT_OVF); return; }
    const bool staged = total <= CP_STAGE;
    uint8_t* dst = staged ? s_out : vp.out + base;
    unsigned long long o = oexcl;
#pragma unroll
    for (int i = 0; i < VT_ITEMS; ++i) {
        if (len[i] == 0) continue;
        const uint32_t h = (w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
        if ((h >> 8) != 255) { dst[o++] = (uint8_t)h; continue; }
        const uint32_t rec = vp.rec_at[p0 + i];
        const uint32_t rlen = (uint32_t)oc.nodes[rec].val & 0xFFFFu;
        for (uint32_t t = 0; t < rlen; ++t) {
            const uint8_t ch = other_char<BITS>(oc, rec, t);
            if (ch != '-') dst[o++] = ch;                          // polish.rs:188 replace("-", "")
        }
    }
    if (!staged) return;
    __syncthreads();
    // coalesced copy of the staged bytes: head to a 16-byte boundary, body as uint4, tail
    uint8_t* g = vp.out + base;
    const uint32_t n = (uint32_t)total;
    const uint32_t head = min(n, (uint32_t)((16 - ((size_t)g & 15)) & 15));
    for (uint32_t i = tid; i < head; i += VT_THREADS) g[i] = s_out[i];
    const uint32_t nvec = (n - head) / 16;
    for (uint32_t i = tid; i < nvec; i += VT_THREADS) {
        const uint8_t* sp = s_out + head + i * 16;
        uint32_t x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = sp[4 * k] | (sp[4 * k + 1] << 8) | (sp[4 * k + 2] << 16) | ((uint32_t)sp[4 * k + 3] << 24);
        reinterpret_cast<uint4*>(g + head)[i] = make_uint4(x[0], x[1], x[2], x[3]);
    }
    for (uint32_t i = head + nvec * 16 + tid; i < n; i += VT_THREADS) g[i] = s_out[i];
}

// ------------------------------------------------------------------------------------------------------
// host side: context, buffers, entry points
// ------------------------------------------------------------------------------------------------------
// (DevBuf, the buffer ids and pp_ctx live in pp_ctx.cuh, shared with tok_kernels.cu)

static void init_comp_table(uint8_t* t) {
    for (int i = 0; i < 256; ++i) t[i] = 'N';
    const char* a = "ATGCNRYSWKMBVDH.-?";
    const char* b = "TACGNYRSWMKVBHD.-?";
    for (int i = 0; a[i]; ++i) t[(unsigned char)a[i]] = (uint8_t)b[i];
}

extern "C" const char* pp_version(void) { return "0.6.1-b200"; }

extern "C" void* pp_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void pp_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int pp_create(int device, pp_ctx** out) {
    if (!out) return PP_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n) { cudaGetLastError(); return PP_ERR_CUDA; }   // no CPU fallback
    pp_ctx* ctx = new pp_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) { delete ctx; return PP_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    if (ctx->l2_persist_max) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    for (auto& ev : ctx->ev) if (cudaEventCreate(&ev) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    for (auto& ev : ctx->side_ev) if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    if (cudaHostAlloc((void**)&ctx->h_status, sizeof(DevStatus), cudaHostAllocDefault) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    if (cudaHostAlloc((void**)&ctx->h_params, sizeof(DevParams), cudaHostAllocDefault) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    uint8_t comp[256];
    init_comp_table(comp);
    if (cudaMemcpyToSymbol(c_comp, comp, 256) != cudaSuccess) { delete ctx; return PP_ERR_CUDA; }
    *out = ctx;
    return PP_OK;
}

extern "C" void pp_destroy(pp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pp_tok_release(ctx);
    for (auto& b : ctx->b) b.release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->side_ev) if (ev) cudaEventDestroy(ev);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_params) cudaFreeHost(ctx->h_params);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* pp_last_error(const pp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// internal (host_api.cpp, filter_kernels.cu): shared access to the context
int pp_ctx_fail(pp_ctx* ctx, int code, const char* msg) { ctx->err = msg; return code; }
int pp_ctx_device(pp_ctx* ctx) { return ctx->device; }
cudaStream_t pp_ctx_stream(pp_ctx* ctx) { return ctx->stream; }
cudaEvent_t pp_ctx_event(pp_ctx* ctx, int i) { return ctx->ev[i]; }
void pp_ctx_count_launches(pp_ctx* ctx, uint32_t n) { ctx->launches = n; }
int pp_ctx_fail_cuda(pp_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    ctx->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what + " (" + file + ":" + std::to_string(line) + ")";
    return PP_ERR_CUDA;
}
void* pp_ctx_scratch2(pp_ctx* ctx, size_t bytes) {
    if (ctx->b[B_SCRATCH2].ensure(bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return ctx->b[B_SCRATCH2].p;
}
void* pp_ctx_scratch(pp_ctx* ctx, size_t bytes) {
    if (ctx->b[B_SCRATCH].ensure(bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return ctx->b[B_SCRATCH].p;
}

template <class T>
static int upload(pp_ctx* ctx, int which, const T* src, size_t n, size_t pad_bytes = 64) {
    CK(ctx->b[which].ensure(n * sizeof(T) + pad_bytes));
    if (n) CK(cudaMemcpyAsync(ctx->b[which].p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return PP_OK;
}

// Draft bases + contig offsets onto the device (shared by pp_dataset_upload and the SAM tokeniser).
int pp_ctx_upload_contigs(pp_ctx* ctx, const pp_contigs* c) {
    if (!c || !c->off || !c->bases || c->n_contigs == 0) return ctx->fail(PP_ERR_ARG, "null or empty contigs");
    const uint64_t G = c->off[c->n_contigs];
    if (G == 0 || G >= 0xFFFFFFFFull - 2 * VT_CHUNK) return ctx->fail(PP_ERR_ARG, "total assembly length must be in [1, 2^32-4096)");
    int rc;
    if ((rc = upload(ctx, B_DRAFT, c->bases, G, VT_CHUNK + 256))) return rc;
    if ((rc = upload(ctx, B_CTGOFF, c->off, (size_t)c->n_contigs + 1))) return rc;
    ctx->G = G; ctx->n_contigs = c->n_contigs;
    return PP_OK;
}

// The alignment arrays in ctx->b[B_CONTIG..B_SEQPOOL] become the resident dataset.
int pp_ctx_commit_dataset(pp_ctx* ctx, uint64_t n_aln, uint64_t n_reads, uint64_t n_ops, uint64_t seq_bytes, uint32_t seq_bits) {
    if (n_aln >= 0xFFFFFFFFull - 2 * CL_CHUNK) return ctx->fail(PP_ERR_ARG, "more than 2^32-2048 alignments");
    ctx->n_aln = n_aln; ctx->n_reads = n_reads; ctx->n_ops = n_ops; ctx->seq_bytes = seq_bytes; ctx->seq_bits = seq_bits;
    // first guesses; a call that overflows one of them grows it and repeats itself
    ctx->node_cap = (uint32_t)std::min<uint64_t>(0x7FFFFFFFull, std::max<uint64_t>(1 << 16, n_aln / 8 + ctx->G / 64));
    ctx->fix_cap = (uint32_t)std::min<uint64_t>(0x7FFFFFFFull, std::max<uint64_t>(1 << 16, n_aln / 4));
    ctx->out_cap = ctx->G + ctx->G / 16 + (1u << 20);
    ctx->global_k = false;
    ctx->have_ds = true;
    return PP_OK;
}

extern "C" int pp_dataset_upload(pp_ctx* ctx, const pp_contigs* c, const pp_alignments* a) {
    if (!ctx) return PP_ERR_ARG;
    if (!c || !a || !c->off || !c->bases || c->n_contigs == 0) return ctx->fail(PP_ERR_ARG, "pp_dataset_upload: null or empty contigs");
    if (a->n_aln && (!a->contig || !a->ref_start || !a->read_id || !a->seq_off || !a->seq_len || !a->cigar_off ||
                     !a->n_cigar || !a->nm || !a->flags || !a->cigar_ops))
        return ctx->fail(PP_ERR_ARG, "pp_dataset_upload: null alignment array");
    if (a->seq_bits != 4 && a->seq_bits != 8) return ctx->fail(PP_ERR_ARG, "pp_dataset_upload: seq_bits must be 4 or 8");
    if (a->n_aln >= 0xFFFFFFFFull - 2 * CL_CHUNK) return ctx->fail(PP_ERR_ARG, "pp_dataset_upload: more than 2^32-2048 alignments");
    CK(cudaSetDevice(ctx->device));
    ctx->have_ds = false;
    int rc;
    if ((rc = upload(ctx, B_CONTIG, a->contig, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_REFSTART, a->ref_start, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_READID, a->read_id, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_SEQOFF, a->seq_off, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_SEQLEN, a->seq_len, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_CIGOFF, a->cigar_off, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_NCIG, a->n_cigar, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_NM, a->nm, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_FLAGS, a->flags, a->n_aln))) return rc;
    if ((rc = upload(ctx, B_CIGOPS, a->cigar_ops, a->n_cigar_ops))) return rc;
    if ((rc = upload(ctx, B_SEQPOOL, a->seq_pool, a->seq_pool_bytes, 256))) return rc;
    if ((rc = pp_ctx_upload_contigs(ctx, c))) { ctx->err = "pp_dataset_upload: " + ctx->err; return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    return pp_ctx_commit_dataset(ctx, a->n_aln, a->n_reads, a->n_cigar_ops, a->seq_pool_bytes, a->seq_bits);
}

static const char* err_text(unsigned code) {
    switch (code) {
        case ERR_UNKNOWN_CONTIG: return "query name in SAM but not in assembly";
        case ERR_SEQ_MISMATCH: return "CIGAR string does not match read sequence";
        case ERR_BAD_OP: return "unexpected character (other than M, =, X, I or D) in CIGAR string - did you use BWA MEM to generate your alignments?";
        case ERR_OOB: return "alignment extends past the end of its reference sequence";
        case ERR_NOSEQ: return "no alignments for read contain sequence";
        default: return "unknown device-side error";
    }
}

template <int BITS>
static int run_polish(pp_ctx* ctx, const pp_polish_params* prm, pp_polish_result* res) {
    cudaStream_t s = ctx->stream;
    const uint64_t G = ctx->G, n_aln = ctx->n_aln;
    const uint32_t n_tiles = (uint32_t)((G + PP_TILE - 1) >> PP_TILE_SHIFT);
    const uint32_t n_chunks = (uint32_t)((G + VT_CHUNK - 1) / VT_CHUNK);
    const uint32_t n_cchunks = (uint32_t)((n_aln + CL_CHUNK - 1) / CL_CHUNK);
    const uint32_t nib_words = (uint32_t)((G + 15) / 16);
    const size_t padG = (size_t)n_chunks * VT_CHUNK + 16;      // k_vote reads whole chunks with vector loads
    const size_t padA = (size_t)n_cchunks * CL_CHUNK + 16;     // k_collect likewise
    int tile_bits = 1;
    while ((1ull << tile_bits) < (uint64_t)n_tiles + 2) tile_bits++;

    CK(ctx->b[B_NIB].ensure(((size_t)nib_words + 8) * 8));
    CK(ctx->b[B_DEPTHFIX].ensure(((size_t)n_tiles * PP_TILE + 1) * 8));
    CK(ctx->b[B_RECGN].ensure(padA * 8)); CK(ctx->b[B_RECK].ensure(padA * 4));
    CK(ctx->b[B_OUTOFF].ensure(((size_t)ctx->n_contigs + 1) * 8));
    CK(ctx->b[B_AGG1].ensure((size_t)n_chunks * 8 + 8));       // vote chunks: prefix of the difference array
    CK(ctx->b[B_AGGC].ensure((size_t)n_cchunks * 8 + 8));      // collect chunks: pairs before the chunk
    CK(ctx->b[B_RES].ensure(padG * 2)); CK(ctx->b[B_RECAT].ensure((G + 1) * 4)); CK(ctx->b[B_CHUNKDELTA].ensure((size_t)n_chunks * 8));
    CK(ctx->b[B_PARAMS].ensure(sizeof(DevParams)));

    ctx->h_params->fv = prm->fraction_valid; ctx->h_params->fi = prm->fraction_invalid;
    ctx->h_params->min_depth = prm->min_depth; ctx->h_params->max_errors = prm->max_errors;
    ctx->h_params->careful = prm->careful ? 1 : 0; ctx->h_params->pad = 0;

    for (int attempt = 0; attempt < 8; ++attempt) {
        const uint32_t fix_cap = ctx->fix_cap, node_cap = ctx->node_cap;
        const uint64_t out_cap = ctx->out_cap;
        // everything that must be zero at the start of a call lives in one pool: one memset
        size_t zoff = 0;
        auto carve = [&](size_t bytes) { size_t o = zoff; zoff += (bytes + 255) & ~size_t(255); return o; };
        const size_t o_diff = carve(padG * 8), o_ex = carve(padG * 8), o_del = carve(padG * 4), o_head = carve((G + 1) * 4),
                     o_tile = carve(((size_t)n_tiles / 32 + 2) * 4),
                     o_chg = carve((size_t)ctx->n_contigs * 8), o_zero = carve((size_t)ctx->n_contigs * 8), o_tdep = carve((size_t)ctx->n_contigs * 8), o_status = carve(sizeof(DevStatus)),
                     o_fkey = carve((size_t)fix_cap * 4), o_fval = carve((size_t)fix_cap * 4),
                     o_k = carve(ctx->global_k ? (ctx->n_reads + 1) * 4 : 4);
        CK(ctx->b[B_ZEROPOOL].ensure(zoff));
        uint8_t* zp = ctx->b[B_ZEROPOOL].as<uint8_t>();
        CK(ctx->b[B_NODES].ensure((size_t)node_cap * sizeof(OthNode)));
        CK(ctx->b[B_FIXKEY2].ensure((size_t)fix_cap * 4)); CK(ctx->b[B_FIXVAL2].ensure((size_t)fix_cap * 4));
        CK(ctx->b[B_FIXRUN].ensure((size_t)n_tiles * 12 + 64));
        CK(ctx->b[B_OUT].ensure(out_cap + 64));
        size_t cub_bytes = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                           (uint32_t*)nullptr, (int)fix_cap, 0, tile_bits, s));
        {
            size_t vscan = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, vscan, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n_chunks));
            CK(ctx->b[B_CUBTMP2].ensure(vscan + 256));
        }
        {
            size_t scan_bytes = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)std::max<uint32_t>(n_cchunks, 1)));
            CK(ctx->b[B_CUBTMP].ensure(std::max(cub_bytes, scan_bytes) + 256));
        }

        DevData d;
        d.n_aln = n_aln;
        d.contig = ctx->b[B_CONTIG].as<uint32_t>(); d.ref_start = ctx->b[B_REFSTART].as<uint32_t>();
        d.read_id = ctx->b[B_READID].as<uint32_t>(); d.seq_off = ctx->b[B_SEQOFF].as<uint32_t>();
        d.cigar_off = ctx->b[B_CIGOFF].as<uint32_t>(); d.nm = ctx->b[B_NM].as<uint32_t>();
        d.cigar_ops = ctx->b[B_CIGOPS].as<uint32_t>(); d.seq_len = ctx->b[B_SEQLEN].as<uint16_t>();
        d.n_cigar = ctx->b[B_NCIG].as<uint16_t>(); d.flags = ctx->b[B_FLAGS].as<uint8_t>();
        d.seq_pool = ctx->b[B_SEQPOOL].as<uint8_t>(); d.draft = ctx->b[B_DRAFT].as<uint8_t>();
        d.contig_off = ctx->b[B_CTGOFF].as<unsigned long long>(); d.n_contigs = ctx->n_contigs; d.G = (uint32_t)G; d.n_tiles = n_tiles;
        d.k = (uint32_t*)(zp + o_k);
        d.draft_nib = ctx->b[B_NIB].as<unsigned long long>(); d.diff = (unsigned long long*)(zp + o_diff);
        d.ex = (unsigned long long*)(zp + o_ex); d.delother = (uint32_t*)(zp + o_del);
        d.oth_head = (uint32_t*)(zp + o_head); d.nodes = ctx->b[B_NODES].as<OthNode>(); d.node_cap = node_cap;
        d.tileflag = (uint32_t*)(zp + o_tile);
        d.rec_gn = ctx->b[B_RECGN].as<unsigned long long>(); d.rec_k = ctx->b[B_RECK].as<uint32_t>();
        d.depth_fix = ctx->b[B_DEPTHFIX].as<double>();
        d.fix_key = (uint32_t*)(zp + o_fkey); d.fix_val = (uint32_t*)(zp + o_fval); d.fix_cap = fix_cap;
        d.prm = ctx->b[B_PARAMS].as<DevParams>();
        d.st = (DevStatus*)(zp + o_status);
        ctx->launches = 0;

        // The per-position counters take ~3 atomics per alignment at random positions: keep as much of them as the
        // hardware allows resident in L2 (persisting access-policy window) while the alignment arrays stream through.
        if (ctx->l2_persist_max && ctx->l2_window_max) {
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof av);
            const size_t want = o_del - o_diff;                                   // diff + ex
            const size_t bytes = std::min(want, ctx->l2_window_max);
            av.accessPolicyWindow.base_ptr = zp + o_diff;
            av.accessPolicyWindow.num_bytes = bytes;
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)ctx->l2_persist_max / (double)bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av));
        }
        // ---- stage 0: reset + derived 4-bit draft plane
        CK(cudaEventRecord(ctx->ev[0], s));
        CK(cudaMemcpyAsync(ctx->b[B_PARAMS].p, ctx->h_params, sizeof(DevParams), cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(zp, 0, zoff, s));
        CK(cudaMemsetAsync(&d.st->err, 0xFF, 8, s));
        if (BITS == 4) {
            CK(cudaMemsetAsync(d.draft_nib + nib_words, 0, 64, s));
            k_draft_nib<<<std::min<uint32_t>((nib_words + 255) / 256, ctx->sm_count * 8), 256, 0, s>>>(d.draft, (uint32_t)G, d.draft_nib, nib_words);
            ctx->launches++;
        }
        // ---- stage 1: (fallback only) global k of multi-record groups
        CK(cudaEventRecord(ctx->ev[1], s));
        if (n_aln && ctx->global_k) {
            k_classify_multi<<<(uint32_t)std::min<uint64_t>((n_aln + 255) / 256, (uint64_t)ctx->sm_count * 8), 256, 0, s>>>(d);
            ctx->launches++;
        }
        // ---- stage 2: scatter
        CK(cudaEventRecord(ctx->ev[2], s));
        if (n_aln) {
            // persistent kernel: exactly as many CTAs as fit on the chip at once (no partial second wave)
            int occ = 1;
            if (ctx->global_k) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scatter<BITS, true>, SC_THREADS, 0));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scatter<BITS, false>, SC_THREADS, 0));
            const uint32_t grid = (uint32_t)std::min<uint64_t>((n_aln + SC_THREADS - 1) / SC_THREADS, (uint64_t)ctx->sm_count * std::max(occ, 1));
            if (ctx->global_k) k_scatter<BITS, true><<<grid, SC_THREADS, 0, s>>>(d);
            else k_scatter<BITS, false><<<grid, SC_THREADS, 0, s>>>(d);
            ctx->launches++;
        }
        // ---- stage 3: ordered depth where k != 1 coverage exists: collect -> stable sort by tile -> ordered walk
        CK(cudaEventRecord(ctx->ev[3], s));
        {   // beside it, on the side stream: the vote chunks' prefix of the difference array
            CK(cudaEventRecord(ctx->side_ev[0], s));
            CK(cudaStreamWaitEvent(ctx->side, ctx->side_ev[0], 0));
            unsigned long long* pre = ctx->b[B_AGG1].as<unsigned long long>();
            k_diff_sums<<<n_chunks, VT_THREADS, 0, ctx->side>>>(d.diff, pre);
            size_t tb = ctx->b[B_CUBTMP2].cap;
            CK(cub::DeviceScan::ExclusiveSum(ctx->b[B_CUBTMP2].p, tb, pre, pre, (int)n_chunks, ctx->side));
            CK(cudaEventRecord(ctx->side_ev[1], ctx->side));
            ctx->launches++;
        }
        if (n_aln) {
            CollectParams cp;
            cp.n_chunks = n_cchunks; cp.chunk_off = ctx->b[B_AGGC].as<unsigned long long>();
            k_collect_count<<<n_cchunks, CL_THREADS, 0, s>>>(d, cp);
            {
                size_t tb = ctx->b[B_CUBTMP].cap;
                CK(cub::DeviceScan::ExclusiveSum(ctx->b[B_CUBTMP].p, tb, cp.chunk_off, cp.chunk_off, (int)n_cchunks, s));
            }
            k_collect<<<n_cchunks, CL_THREADS, 0, s>>>(d, cp);
            ctx->launches += 2;          // (the scan and the sort are library launches, not counted)
            CK(cub::DeviceRadixSort::SortPairs(ctx->b[B_CUBTMP].p, cub_bytes, d.fix_key, ctx->b[B_FIXKEY2].as<uint32_t>(), d.fix_val,
                                               ctx->b[B_FIXVAL2].as<uint32_t>(), (int)fix_cap, 0, tile_bits, s));
            uint32_t* runs = ctx->b[B_FIXRUN].as<uint32_t>();
            k_fix_runs<<<(fix_cap + 255) / 256, 256, 0, s>>>(d, ctx->b[B_FIXKEY2].as<uint32_t>(), runs, runs + n_tiles, runs + 2 * (size_t)n_tiles);
            k_depth_fixup<<<std::min<uint32_t>((n_tiles + FX_WARPS - 1) / FX_WARPS, ctx->sm_count * 16), FX_WARPS * 32, 0, s>>>(d, ctx->b[B_FIXVAL2].as<uint32_t>(), runs,
                                                                                                                            runs + n_tiles, runs + 2 * (size_t)n_tiles);
            ctx->launches += 2;
        }
        // ---- stage 5: vote; stage 4: compaction
        CK(cudaEventRecord(ctx->ev[4], s));
        VoteParams vp;
        vp.n_chunks = n_chunks;
        vp.out = ctx->b[B_OUT].as<uint8_t>(); vp.out_cap = out_cap;
        vp.out_off = ctx->b[B_OUTOFF].as<unsigned long long>();
        vp.changed = (unsigned long long*)(zp + o_chg); vp.zero_depth = (unsigned long long*)(zp + o_zero); vp.total_depth = (double*)(zp + o_tdep);
        vp.chunk_pre = ctx->b[B_AGG1].as<unsigned long long>();
        CK(cudaStreamWaitEvent(s, ctx->side_ev[1], 0));
        vp.res = ctx->b[B_RES].as<uint16_t>(); vp.rec_at = ctx->b[B_RECAT].as<uint32_t>();
        vp.chunk_delta = ctx->b[B_CHUNKDELTA].as<long long>();
        vp.dbg = nullptr;
        if (ctx->debug_on) { CK(ctx->b[B_DEBUG].ensure((G + 1) * sizeof(pp_debug_pos))); vp.dbg = ctx->b[B_DEBUG].as<pp_debug_pos>(); }
        k_vote<BITS><<<n_chunks, VT_THREADS, 0, s>>>(d, vp);
        CK(cudaEventRecord(ctx->ev[5], s));
        k_compact<BITS><<<n_chunks, VT_THREADS, 0, s>>>(d, vp);
        ctx->launches += 2;
        CK(cudaEventRecord(ctx->ev[6], s));
        CK(cudaMemcpyAsync(ctx->h_status, d.st, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        const DevStatus hs = *ctx->h_status;
        if (hs.err != ~0ull) {
            res->error_aln = (int64_t)(hs.err >> 8);
            return ctx->fail(PP_ERR_INPUT, std::string(err_text((unsigned)(hs.err & 0xFF))) + " (alignment " + std::to_string(hs.err >> 8) + ")");
        }
        // a side buffer was too small or a read group too large for the in-kernel scan: grow / switch mode and repeat
        bool again = false;
        if (hs.flags & FL_BIGGROUP) { if (ctx->global_k) return ctx->fail(PP_ERR_CUDA, "internal error: FL_BIGGROUP in global-k mode"); ctx->global_k = true; again = true; }
        if (hs.flags & FL_NODE_OVF) { ctx->node_cap = (uint32_t)std::min<uint64_t>(0x7FFFFFFFull, (uint64_t)hs.node_count + hs.node_count / 4 + 1024); again = true; }
        if (hs.flags & FL_FIX_OVF) {
            if (hs.fix_count >= 0x7FFFFFFFull) return ctx->fail(PP_ERR_NOMEM, "too many multi-mapped (alignment, tile) pairs for one call");
            ctx->fix_cap = (uint32_t)(hs.fix_count + hs.fix_count / 8 + 1024); again = true;
        }
        if (!again && (hs.flags & FL_OUT_OVF)) { ctx->out_cap = hs.out_len + 64; again = true; }
        if (again) continue;
        if (hs.flags & FL_COUNTER_OVF)
            return ctx->fail(PP_ERR_INPUT, "a position is covered by 65536 or more alignments: not supported by this build's 16-bit allele counters");

        // the side buffers shrink to what this dataset needs (sort length, zeroing); a later call with other options that
        // needs more overflows once and grows them again
        ctx->fix_cap = std::min<uint32_t>(ctx->fix_cap, (uint32_t)std::max<uint64_t>(1 << 16, hs.fix_count + hs.fix_count / 8 + 1024));

        ctx->have_debug = ctx->debug_on; ctx->last_head = d.oth_head; ctx->last_nodes = std::min(hs.node_count, node_cap);
        res->out_len = hs.out_len;
        res->n_aln_used = hs.n_used;
        res->error_aln = -1;
        memset(&res->timing, 0, sizeof res->timing);
        float ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); res->timing.stage_ms[0] = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); res->timing.stage_ms[1] = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); res->timing.stage_ms[2] = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4])); res->timing.stage_ms[3] = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5])); res->timing.stage_ms[5] = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6])); res->timing.stage_ms[4] = ms;
        CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[6])); res->timing.total_ms = ms;
        res->timing.launches = ctx->launches;

        if (res->out_bases) {
            if (res->out_cap < hs.out_len) return ctx->fail(PP_ERR_ARG, "out_cap too small: need " + std::to_string(hs.out_len) + " bytes");
            CK(cudaEventRecord(ctx->ev[7], s));
            CK(cudaMemcpyAsync(res->out_bases, vp.out, hs.out_len, cudaMemcpyDeviceToHost, s));
            if (res->out_off) CK(cudaMemcpyAsync(res->out_off, vp.out_off, ((size_t)ctx->n_contigs + 1) * 8, cudaMemcpyDeviceToHost, s));
            if (res->changed) CK(cudaMemcpyAsync(res->changed, vp.changed, (size_t)ctx->n_contigs * 8, cudaMemcpyDeviceToHost, s));
            if (res->zero_depth) CK(cudaMemcpyAsync(res->zero_depth, vp.zero_depth, (size_t)ctx->n_contigs * 8, cudaMemcpyDeviceToHost, s));
            if (res->total_depth) CK(cudaMemcpyAsync(res->total_depth, vp.total_depth, (size_t)ctx->n_contigs * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaEventRecord(ctx->ev[8], s));
            CK(cudaStreamSynchronize(s));
            CK(cudaEventElapsedTime(&ms, ctx->ev[7], ctx->ev[8]));
            res->timing.stage_ms[7] = ms;
        }
        return PP_OK;
    }
    return ctx->fail(PP_ERR_NOMEM, "side buffers kept overflowing");
}

static int check_params(pp_ctx* ctx, const pp_polish_params* p) {
    if (!p) return ctx->fail(PP_ERR_ARG, "null params");
    // polish.rs:277-287 (the same text the reference prints)
    if (!(p->fraction_valid > 0.0 && p->fraction_valid < 1.0)) return ctx->fail(PP_ERR_INPUT, "--fraction_valid must be between 0 and 1 (exclusive)");
    if (!(p->fraction_invalid > 0.0 && p->fraction_invalid < 1.0)) return ctx->fail(PP_ERR_INPUT, "--fraction_invalid must be between 0 and 1 (exclusive)");
    if (p->fraction_invalid >= p->fraction_valid) return ctx->fail(PP_ERR_INPUT, "--fraction_invalid must be less than --fraction_valid");
    return PP_OK;
}

extern "C" int pp_polish_resident(pp_ctx* ctx, const pp_polish_params* params, pp_polish_result* result) {
    if (!ctx) return PP_ERR_ARG;
    if (!result) return ctx->fail(PP_ERR_ARG, "null result");
    if (!ctx->have_ds) return ctx->fail(PP_ERR_ARG, "pp_polish_resident: no dataset uploaded");
    int rc = check_params(ctx, params);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    result->error_aln = -1;
    return ctx->seq_bits == 4 ? run_polish<4>(ctx, params, result) : run_polish<8>(ctx, params, result);
}

extern "C" int pp_polish(pp_ctx* ctx, const pp_contigs* contigs, const pp_alignments* alns,
                         const pp_polish_params* params, pp_polish_result* result) {
    if (!ctx) return PP_ERR_ARG;
    if (!result) return ctx->fail(PP_ERR_ARG, "null result");
    int rc = check_params(ctx, params);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev[9], ctx->stream));
    rc = pp_dataset_upload(ctx, contigs, alns);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[10], ctx->stream));
    CK(cudaEventSynchronize(ctx->ev[10]));
    float h2d = 0;
    CK(cudaEventElapsedTime(&h2d, ctx->ev[9], ctx->ev[10]));
    result->error_aln = -1;
    rc = ctx->seq_bits == 4 ? run_polish<4>(ctx, params, result) : run_polish<8>(ctx, params, result);
    if (rc == PP_OK) result->timing.stage_ms[6] = h2d;
    return rc;
}

extern "C" int pp_polish_set_debug(pp_ctx* ctx, int on) {
    if (!ctx) return PP_ERR_ARG;
    ctx->debug_on = on == 1;            // 2 = stop recording but keep the last call's records readable
    if (on == 0) ctx->have_debug = false;
    return PP_OK;
}

extern "C" int pp_polish_debug_fetch(pp_ctx* ctx, uint64_t first_pos, uint64_t n_pos, pp_debug_pos* out) {
    if (!ctx) return PP_ERR_ARG;
    if (!ctx->have_debug) return ctx->fail(PP_ERR_ARG, "pp_polish_debug_fetch: the last polish did not record debug positions (pp_polish_set_debug)");
    if (!out || first_pos + n_pos > ctx->G) return ctx->fail(PP_ERR_ARG, "pp_polish_debug_fetch: range outside the assembly");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->b[B_DEBUG].as<pp_debug_pos>() + first_pos, n_pos * sizeof(pp_debug_pos), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PP_OK;
}

extern "C" int pp_polish_debug_alleles(pp_ctx* ctx, uint32_t* head, pp_debug_node* nodes, uint64_t node_cap, uint64_t* n_nodes) {
    if (!ctx) return PP_ERR_ARG;
    if (!ctx->have_debug || !ctx->last_head) return ctx->fail(PP_ERR_ARG, "pp_polish_debug_alleles: no debug polish on this context");
    if (n_nodes) *n_nodes = ctx->last_nodes;
    if (!head && !nodes && node_cap == 0) return PP_ERR_ARG;       // size query
    if (!head || (ctx->last_nodes && !nodes) || node_cap < ctx->last_nodes) return ctx->fail(PP_ERR_ARG, "pp_polish_debug_alleles: buffers too small");
    static_assert(sizeof(pp_debug_node) == sizeof(OthNode), "debug node layout");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(head, ctx->last_head, ctx->G * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->last_nodes) CK(cudaMemcpyAsync(nodes, ctx->b[B_NODES].p, (size_t)ctx->last_nodes * sizeof(OthNode), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PP_OK;
}
