// bp_edge_params.h -- launch parameters of the edge-parallel kernel family (see bp_edge.cuh).
#pragma once
#include <stdint.h>

namespace bpb {

struct EdgeParams {
    // graph tables in global memory (the streaming family's blob): sorted CSR, CSC as a permutation of CSR positions
    const uint32_t *row_ptr, *col_idx, *col_ptr, *csc2csr, *row_idx;
    const double *prior;
    int m, n, nnz;
    int G, GV;            // lanes per row / per column: powers of two >= the largest row / column degree (<= 32)
    int MW;               // 32-bit words per syndrome, ceil(m / 32)
    int max_iter;
    double ms_scaling;
    int uniform_prior;
    double prior0;
    const uint32_t *synd_packed;  // [B][mwp]
    int mwp;
    long long batch;
    unsigned long long *counter;
    const uint32_t *index_list;           // second stage: batch indices to decode (null = 0..batch-1)
    const unsigned long long *batch_dev;  // second stage: number of entries of index_list (device value)
    double *msg_global;   // [grid][nnz] message scratch when one syndrome's messages do not fit in shared memory
    uint8_t *out_dec;     // [B][n]
    uint8_t *out_conv;    // [B] or null
    int32_t *out_iters;   // [B] or null
    double *out_llr;      // [B][n] or null
    int llr_last_only;
};

using EdgeKernel = void (*)(const EdgeParams);

// defined in bp_edge_{ms,ps}.cu
EdgeKernel pick_edge_ms(bool llr, bool msg_global);
EdgeKernel pick_edge_ps(bool llr, bool msg_global);

}  // namespace bpb
