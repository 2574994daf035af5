// bp_pair_params.h -- launch parameters of the paired on-chip kernel family (see bp_pair.cuh).
#pragma once
#include <stdint.h>

namespace bpb {

struct PairParams {
    const uint32_t *tab;  // table blob in global memory, copied verbatim to the start of shared memory
    uint32_t tab_bytes;   // multiple of 16
    uint32_t off_row_deg, off_col_deg, off_col_row, off_row_pos, off_col_pos, off_prior;  // byte offsets in the blob
    uint32_t group_bytes;                                      // per-group area (multiple of 16)
    uint32_t goff_msg, goff_syn, goff_acc, goff_ctl;  // byte offsets inside a group area
    int m, n, M, N;       // M, N: padded row / column counts (table strides, multiples of 32)
    int MW;               // 32-bit words per syndrome ceil(m / 32)
    int msg_slots;        // double2 slots of one group's message array
    int groups, T;        // thread groups per CTA, threads per group (multiple of 32)
    int max_iter;
    double ms_scaling;
    int uniform_prior;
    double prior0;
    const uint32_t *synd_packed;  // [B][mwp]
    int mwp;
    long long batch;
    unsigned long long *counter;
    const uint32_t *index_list;             // second stage: batch indices to decode (null = 0..batch-1)
    const unsigned long long *batch_dev;    // second stage: number of entries of index_list (device value)
    uint8_t *out_dec;     // [B][n]
    uint8_t *out_conv;    // [B] or null
    int32_t *out_iters;   // [B] or null
    double *out_llr;      // [B][n] or null
    int llr_last_only;    // BP+OSD: write posterior LLRs only in iteration max_iter (only non-convergers need them)
};

using PairKernel = void (*)(const PairParams);

// defined in bp_pair_{ms,ps}.cu; nullptr when no degree bucket fits.  cta_threads: 512 (every bucket), 640 or 768
// (regular (3,6) codes only): the CTA size the kernel is compiled for.
PairKernel pick_pair_ms(int max_row_degree, int max_col_degree, bool regular, bool llr, int cta_threads);
PairKernel pick_pair_ps(int max_row_degree, int max_col_degree, bool regular, bool llr, int cta_threads);

}  // namespace bpb
