// bp_smem_params.h -- launch parameters of the on-chip kernel family (see bp_smem.cuh).
#pragma once
#include <stdint.h>

namespace bpb {

struct SmemParams {
    const uint32_t *tab;  // table blob in global memory, copied verbatim to the start of shared memory
    uint32_t tab_bytes;   // multiple of 16
    uint32_t off_row_deg, off_col_deg, off_col_row, off_row_pos, off_col_pos, off_prior;
    uint32_t off_col_self, off_lev_ptr, off_lev_bits;  // serial schedule only (bp_smem_serial.cuh)
    int n_levels;  // byte offsets inside the blob
    uint32_t group_bytes;                                                    // per-group area (multiple of 16)
    uint32_t goff_msg, goff_dec, goff_syn, goff_ctl;                         // byte offsets inside a group area
    int m, n, M, N;       // M, N: padded row / column counts (ELL strides)
    int MW;               // 32-bit words per syndrome, ceil(m / 32)
    int groups, T;        // thread groups per CTA, threads per group
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

using SmemKernel = void (*)(const SmemParams);

// CTA size the kernels are compiled for (__launch_bounds__): 1024 threads (64 registers each) only for the small
// min-sum bucket, 512 (128 registers) otherwise.
inline int smem_cta_threads(int method, int dc, int dv) { return (method == 1 && dc <= 8 && dv <= 4) ? 1024 : 512; }

// defined in bp_smem_{ms,ps}.cu; nullptr when no degree bucket fits
// `regular`: every row has exactly max_row_degree entries and every column exactly max_col_degree
SmemKernel pick_smem_ms(int max_row_degree, int max_col_degree, bool regular, bool llr);
SmemKernel pick_smem_ps(int max_row_degree, int max_col_degree, bool regular, bool llr);
// serial schedule (bp_smem_serial.cuh), always 512-thread CTAs
SmemKernel pick_smem_serial_ms(int max_row_degree, int max_col_degree, bool regular, bool llr);
SmemKernel pick_smem_serial_ps(int max_row_degree, int max_col_degree, bool regular, bool llr);

}  // namespace bpb
