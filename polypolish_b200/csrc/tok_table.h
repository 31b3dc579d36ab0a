// tok_table.h — host image of what the device tokeniser looks names up in: the contig-name hash table and the BAM nibble
// table, laid out as tok::ContigTable reads them.  Shared by tok_kernels.cu and tests/tok_harness.cpp.
#pragma once
#include <string>
#include <vector>

#include "tok_line.h"

namespace tok {

struct TableImage {
    std::vector<uint8_t> bytes;
    size_t o_hash = 0, o_slot = 0, o_off = 0, o_names = 0, o_nib = 0;
    uint32_t mask = 0;
};

inline uint64_t fnv1a(const std::string& s) {
    uint64_t h = FNV_BASIS;
    for (unsigned char c : s) h = (h ^ c) * FNV_PRIME;
    return h;
}

inline void build_table_image(const std::vector<std::string>& names_in, TableImage& im) {
    const size_t nc = names_in.size();
    size_t cap = 4;
    while (cap < 2 * nc + 2) cap <<= 1;
    im.mask = (uint32_t)(cap - 1);
    size_t names_len = 0;
    for (auto& s : names_in) names_len += s.size();
    auto up = [](size_t v) { return (v + 15) & ~size_t(15); };
    im.o_hash = 0;
    im.o_slot = up(im.o_hash + cap * 8);
    im.o_off = up(im.o_slot + cap * 4);
    im.o_names = up(im.o_off + (nc + 1) * 4);
    im.o_nib = up(im.o_names + names_len + 1);
    im.bytes.assign(im.o_nib + 256, 0);
    uint64_t* hash = (uint64_t*)(im.bytes.data() + im.o_hash);
    uint32_t* slot = (uint32_t*)(im.bytes.data() + im.o_slot);
    uint32_t* off = (uint32_t*)(im.bytes.data() + im.o_off);
    uint8_t* names = im.bytes.data() + im.o_names;
    uint32_t o = 0;
    for (size_t c = 0; c < nc; ++c) {
        off[c] = o;
        memcpy(names + o, names_in[c].data(), names_in[c].size());
        o += (uint32_t)names_in[c].size();
        const uint64_t h = fnv1a(names_in[c]);
        uint32_t sl = (uint32_t)h & im.mask;
        while (slot[sl]) sl = (sl + 1) & im.mask;      // names are unique (misc.rs:66-75)
        slot[sl] = (uint32_t)c + 1;
        hash[sl] = h;
    }
    off[nc] = o;
    uint8_t* nib = im.bytes.data() + im.o_nib;
    const char* codes = "=ACMGRSVTWYHKDBN";
    for (int i = 1; i < 16; ++i) {
        nib[(unsigned char)codes[i]] = (uint8_t)i;
        nib[(unsigned char)(codes[i] + 32)] = (uint8_t)i;   // SEQ is upper-cased (alignment.rs:94)
    }
}

inline ContigTable table_view(const TableImage& im, const uint8_t* base) {
    ContigTable ct;
    ct.hash = (const uint64_t*)(base + im.o_hash);
    ct.slot = (const uint32_t*)(base + im.o_slot);
    ct.name_off = (const uint32_t*)(base + im.o_off);
    ct.names = base + im.o_names;
    ct.mask = im.mask;
    return ct;
}

}  // namespace tok
