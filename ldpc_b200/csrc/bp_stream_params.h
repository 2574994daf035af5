// bp_stream_params.h -- launch parameters of the HBM-streaming kernel family (see bp_stream.cuh).
#pragma once
#include <stdint.h>

namespace bpb {

struct StreamParams {
    // graph blob: [row_ptr m+1][col_idx nnz][col_ptr n+1][csc2csr nnz][row_idx nnz][pad][prior n doubles]
    const uint32_t *blob;
    uint32_t blob_words;  // 32-bit words in the blob
    uint32_t prior_off;   // word offset of prior[] (even)
    int m, n, nnz;
    int mwp;              // 32-bit words per packed syndrome row (multiple of 4)
    int m_pad, n_pad;     // multiples of 32
    int max_iter;
    double ms_scaling;
    int uniform_prior;
    double prior0;
    const uint32_t *synd_packed;  // [B][mwp]
    long long batch;
    unsigned long long *counter;
    double *msg;          // [warps][nnz][32]
    uint32_t *dec_w;      // [warps][n_pad]
    uint32_t *syn_w_g;    // [warps][m_pad] (used when the syndrome words do not fit in shared memory)
    double *llr_tile;     // [warps][n][32] (only when LLRs are requested)
    uint8_t *out_dec;     // [B][n]
    uint8_t *out_conv;    // [B] or null
    int32_t *out_iters;   // [B] or null
    double *out_llr;      // [B][n] or null
    int llr_last_only;    // BP+OSD: copy posterior LLRs out only for syndromes that did not converge
    const uint32_t *order;  // serial schedule: levelised batches of SerialBatch bits, 0xffffffff = padding
    int order_len;
    int iter_cap;                       // hand-off threshold (>= max_iter disables the second stage)
    unsigned long long *handoff_count;  // number of syndromes handed to the second stage
    uint32_t *handoff_list;             // their batch indices
    unsigned long long *iter_total;     // sum of iterations executed by this launch
    int serial_no_init;   // regular-code serial program carries the visited flags: no message initialisation in memory
    int compact_num, compact_den;  // compact when sectors needed * num <= sectors occupied * den (default 2 / 1)
    int no_compaction;    // tuning / testing: keep live lanes where they are during the ramp-down
    int smem_graph, smem_syn;
    uint32_t smem_syn_off;  // word offset of the per-warp syndrome words in dynamic shared memory
};

using StreamKernel = void (*)(const StreamParams);

constexpr uint32_t kProgRowMask = 0x000fffffu;  // serial program word: row index in the low 20 bits

// Bits of one level the serial-schedule kernel keeps in flight together (host pads the schedule to this).
#ifndef BPB_SERIAL_SB_UNI
#define BPB_SERIAL_SB_UNI 2  // measured: 2 bits x 16 warps/SM beats 4 bits x 8 warps/SM by 6 % at n = 10^4
#endif
template <int DC, int DV, bool UNI> struct SerialBatch { static constexpr int v = UNI ? BPB_SERIAL_SB_UNI : ((DC <= 8 && DV <= 4) ? 2 : 1); };
inline int serial_batch(int dc, int dv, bool regular) {
    if (dc > 32 || dv > 16) return 1;  // generic kernel: plain order
    if (regular && dc == 6 && dv == 3) return SerialBatch<6, 3, true>::v;
    if (dc <= 8 && dv <= 4) return SerialBatch<8, 4, false>::v;
    return 1;
}

// defined in bp_stream_{ms,ps}_{parallel,serial}.cu; nullptr when no degree bucket fits
StreamKernel pick_stream_ms_parallel(int max_row_degree, int max_col_degree, bool regular, bool llr);
StreamKernel pick_stream_ps_parallel(int max_row_degree, int max_col_degree, bool regular, bool llr);
StreamKernel pick_stream_ms_serial(int max_row_degree, int max_col_degree, bool regular, bool llr);
StreamKernel pick_stream_ps_serial(int max_row_degree, int max_col_degree, bool regular, bool llr);

}  // namespace bpb
