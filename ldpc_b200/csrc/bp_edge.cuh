// bp_edge.cuh -- the edge-parallel kernel family ("one CTA = one syndrome, one lane = one edge, reductions by
// warp shuffle").
//
// Replaces ldpc::bp::BpDecoder::bp_decode_parallel (reference src_cpp/bp.hpp:192-325) for the regimes the other two
// families serve badly: small batches and single decodes (the thread-group family gives a syndrome 128 threads that
// walk the rows one after another; here every edge of H has its own lane), and the ramp-down / second stage of codes
// whose messages do not fit in shared memory (messages then live in a per-CTA scratch in global memory, which the
// 126 MB L2 keeps resident).
//
// Check pass.  A row of H occupies G consecutive lanes, G = the power of two >= the largest row degree (lane k of
// the group holds the row's k-th edge in ascending column order, the reference's iterate_row order).
//   min-sum (bp.hpp:220-273): the reference's prefix / suffix running minima give |c_k| = min_{k' != k} |b_k'| where
//     a NaN or a value >= DBL_MAX never replaces the running value.  min over the clamped magnitudes
//     a' = (|b| < DBL_MAX ? |b| : DBL_MAX) is exact and order-free, so a butterfly does it: after exchanging with
//     lane^1, lane^2, lane^4 ... (each time the minimum of the OTHER half of the current block) the minimum of the
//     received values is the minimum over all other lanes of the group: log2(G) shuffles.  The sign is the parity of
//     the syndrome bit plus a ballot of (b <= 0) over the group (bp.hpp:236-260).
//   product-sum (bp.hpp:201-219): the prefix products P_k = ((1*t_0)*t_1)... and suffix products are taken in the
//     reference's order (floating-point products do not re-associate): G-1 shuffle steps in which lane k multiplies
//     by t_j for j < k (prefix) and j > k (suffix), then c_k = sigma * log((1 + P_k*S_k) / (1 - P_k*S_k)).
// Bit pass (bp.hpp:276-318).  A column occupies GV lanes in ascending row order; the running sums
//   pre_k = ((prior + c_0) + c_1) ... and suf_k = ((0 + c_{d-1}) + c_{d-2}) ... are again taken in the reference's
//   order with GV-1 shuffle steps each; b_k = pre_k + suf_k, the posterior is pre_{d-1} + c_{d-1}.  A decided-1 bit
//   XORs its checks into the candidate-syndrome accumulator exactly like the thread-group family.
// Messages are stored once per edge at its CSR position and updated in place (b2c -> c2b -> b2c).
#pragma once
#include "bp_edge_params.h"
#include "bp_update.cuh"

namespace bpb {

template <int METHOD, bool LLR, bool MSG_GLOBAL>
__global__ void __launch_bounds__(1024, 1) bp_edge_kernel(const EdgeParams p) {
    extern __shared__ __align__(16) uint8_t esm[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int m = p.m, n = p.n, nnz = p.nnz, G = p.G, GV = p.GV;
    double *msg = MSG_GLOBAL ? (p.msg_global + (size_t) blockIdx.x * (size_t) nnz) : reinterpret_cast<double *>(esm);
    uint8_t *rest = esm + (MSG_GLOBAL ? 0 : (size_t) nnz * 8);
    uint32_t *synw = reinterpret_cast<uint32_t *>(rest);
    uint32_t *acc = synw + p.MW;
    uint8_t *dec = reinterpret_cast<uint8_t *>(acc + p.MW);
    __shared__ long long ctl;
    const uint32_t *row_ptr = p.row_ptr, *col_idx = p.col_idx, *col_ptr = p.col_ptr, *csc2csr = p.csc2csr,
                   *row_idx = p.row_idx;
    const int rows_per_pass = T / G, cols_per_pass = T / GV;
    const int kr = tid & (G - 1), kc = tid & (GV - 1);
    const int lane = tid & 31;
    // lanes of my row group / column group inside the warp
    const uint32_t rmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    const long long limit = p.batch_dev ? (long long) *p.batch_dev : p.batch;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const long long claim = (long long) atomicAdd(p.counter, 1ull);
            ctl = (claim < limit) ? (p.index_list ? (long long) p.index_list[claim] : claim) : -1;
        }
        __syncthreads();
        const long long idx = ctl;
        if (idx < 0) break;
        const uint32_t *srow = p.synd_packed + idx * p.mwp;
        for (int w = tid; w < p.MW; w += T) synw[w] = __ldg(srow + w);
        for (int e = tid; e < nnz; e += T) msg[e] = p.uniform_prior ? p.prior0 : p.prior[__ldg(col_idx + e)];  // bp.hpp:147-157
        for (int j = tid; j < n; j += T) dec[j] = 0;
        __syncthreads();

        int it = 0;
        bool conv = false;
        while (it < p.max_iter) {
            ++it;
            const double alpha = ms_alpha(p.ms_scaling, it);
            for (int w = tid; w < p.MW; w += T) acc[w] = synw[w];
            // ---- check -> bit (bp.hpp:201-273): G lanes per row ------------------------------------------
            for (int r0 = 0; r0 < m; r0 += rows_per_pass) {  // whole warps run the same number of passes
                const int i = r0 + tid / G;
                uint32_t beg = 0;
                int deg = 0;
                if (i < m) {
                    beg = __ldg(row_ptr + i);
                    deg = (int) (__ldg(row_ptr + i + 1) - beg);
                }
                const bool live = kr < deg;
                const double b = live ? msg[beg + kr] : 0.0;
                const uint32_t s = (i < m) ? ((synw[i >> 5] >> (i & 31)) & 1u) : 0u;
                double c;
                if (METHOD == kMinimumSum) {
                    const bool neg = live && (b <= 0);
                    const uint32_t negs = __ballot_sync(0xffffffffu, neg) & rmask;
                    const double a = fabs(b);
                    const double mine = (live && a < DBL_MAX) ? a : DBL_MAX;  // what a strict '<' lets through
                    double blockmin = mine, others = DBL_MAX;
                    for (int off = 1; off < G; off <<= 1) {
                        const double got = __shfl_xor_sync(0xffffffffu, blockmin, off);  // min of the sibling block
                        others = (got < others) ? got : others;
                        blockmin = (got < blockmin) ? got : blockmin;
                    }
                    const uint32_t sg = s + (uint32_t) __popc(negs) + (neg ? 1u : 0u);  // bp.hpp:252-260
                    c = others * ((sg & 1u) ? -alpha : alpha);                           // bp.hpp:262
                } else {
                    const double t = live ? ps_tanh_half(b) : 1.0;
                    double pre = 1.0, suf = 1.0;
                    const int base = lane & ~(G - 1);
                    for (int j = 0; j < G - 1; ++j) {  // prefix: multiply by t_0, t_1, ... t_{k-1} in that order
                        const double tj = __shfl_sync(0xffffffffu, t, base + j);
                        if (j < kr && j < deg) pre *= tj;
                    }
                    for (int j = G - 1; j > 0; --j) {  // suffix: t_{d-1}, t_{d-2}, ... t_{k+1}
                        const double tj = __shfl_sync(0xffffffffu, t, base + j);
                        if (j > kr && j < deg) suf *= tj;
                    }
                    const double x = pre * suf;
                    c = live ? (s ? -1.0 : 1.0) * ps_atanh2(x) : 0.0;
                }
                if (live) msg[beg + kr] = c;
            }
            __syncthreads();
            // ---- posterior, decision, bit -> check (bp.hpp:276-318): GV lanes per column ------------------------
            for (int c0 = 0; c0 < n; c0 += cols_per_pass) {
                const int j = c0 + tid / GV;
                uint32_t beg = 0;
                int deg = 0;
                if (j < n) {
                    beg = __ldg(col_ptr + j);
                    deg = (int) (__ldg(col_ptr + j + 1) - beg);
                }
                const bool live = kc < deg;
                const uint32_t e = live ? __ldg(csc2csr + beg + kc) : 0u;
                const double c = live ? msg[e] : 0.0;
                const double prior = (j < n) ? (p.uniform_prior ? p.prior0 : p.prior[j]) : 0.0;
                const int base = lane & ~(GV - 1);
                double pre = prior, suf = 0.0;
                for (int q = 0; q < GV - 1; ++q) {
                    const double cq = __shfl_sync(0xffffffffu, c, base + q);
                    if (q < kc && q < deg) pre += cq;
                }
                for (int q = GV - 1; q > 0; --q) {
                    const double cq = __shfl_sync(0xffffffffu, c, base + q);
                    if (q > kc && q < deg) suf += cq;
                }
                // posterior = pre_{d-1} + c_{d-1}, computed by the group's last live lane (lane 0 for a degree-0 column)
                const int last = deg > 0 ? deg - 1 : 0;
                const double tot = (kc == last) ? (deg > 0 ? pre + c : prior) : 0.0;
                const double llr = __shfl_sync(0xffffffffu, tot, base + last);
                if (live) msg[e] = pre + suf;
                const bool x = (j < n) && (llr <= 0);
                if (j < n && kc == 0) {
                    dec[j] = x ? 1 : 0;
                    if (LLR && (!p.llr_last_only || it == p.max_iter)) p.out_llr[idx * n + j] = llr;
                }
                if (x && live) {  // bp.hpp:290-294
                    const uint32_t r = __ldg(row_idx + beg + kc);
                    atomicXor(&acc[r >> 5], 1u << (r & 31));
                }
            }
            __syncthreads();
            uint32_t bad = 0;
            for (int w = tid; w < p.MW; w += T) bad |= acc[w];
            conv = !__syncthreads_or(bad != 0);
            if (conv) break;
        }
        uint8_t *drow = p.out_dec + idx * n;
        for (int j = tid; j < n; j += T) drow[j] = dec[j];
        if (tid == 0) {
            if (p.out_iters) p.out_iters[idx] = it;
            if (p.out_conv) p.out_conv[idx] = conv ? 1 : 0;
        }
    }
}

template <int METHOD>
EdgeKernel pick_edge_kernel(bool llr, bool msg_global) {
    if (llr) return msg_global ? bp_edge_kernel<METHOD, true, true> : bp_edge_kernel<METHOD, true, false>;
    return msg_global ? bp_edge_kernel<METHOD, false, true> : bp_edge_kernel<METHOD, false, false>;
}

}  // namespace bpb
