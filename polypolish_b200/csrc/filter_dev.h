// filter_dev.h — the device-side interface of the filter proper (filter_kernels.cu), shared with the device SAM path
// (tok_kernels.cu).  Not part of the ABI.
#pragma once
#include <stdint.h>

#include "../../include/pp_abi.h"

struct Mate {
    const uint32_t *name_id, *contig, *ref_start, *ref_end;
    const uint8_t* flags;
    uint32_t* cnt;     // [n_names] aligned records per name
    uint32_t* head;    // [n_names] list head (record index) or 0xFFFFFFFF
    uint32_t* next;    // [n] next record of the same name
    uint8_t* pass;     // [n]
    uint32_t n;
};


int pp_filter_core(pp_ctx* ctx, const Mate in[2], const pp_filter_params* prm, pp_filter_result* res, const uint8_t* d_pass[2],
                   uint64_t n_pass_mate[2]);
