// CPU harness for polypolish_b200/csrc/nib_utils.h (compiled by tests/test_nib_utils.py with g++).
#include "../polypolish_b200/csrc/nib_utils.h"
extern "C" {
void h_load_nib32(const unsigned long long* w, uint32_t start, unsigned long long* out) { load_nib32(w, start, out[0], out[1]); }
void h_load_read32(const unsigned long long* w, uint32_t len, int rc, uint32_t ri, unsigned long long* out) { load_read32(w, len, rc != 0, ri, out[0], out[1]); }
void h_mismatch(const unsigned long long* r, const unsigned long long* d, uint32_t vc, unsigned long long* out) { mismatch_masks(r[0], r[1], d[0], d[1], vc, out[0], out[1]); }
}
