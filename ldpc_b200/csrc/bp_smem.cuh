// bp_smem.cuh -- the on-chip kernel family ("one thread group = one syndrome, messages in shared memory").
//
// Replaces ldpc::bp::BpDecoder::bp_decode_parallel (reference src_cpp/bp.hpp:192-325) when 8*E bytes of
// messages fit in shared memory several times over.  The algorithmic HBM traffic of the streaming family
// (4*E*8 bytes per iteration per syndrome) never leaves the SM here: per syndrome only the
// packed syndrome comes in and the hard decisions (plus, optionally, the posterior LLRs) go out.
//
// One CTA per SM holds ONE copy of the graph tables and G independent thread groups of T threads; each group
// decodes one syndrome at a time and claims the next from a global counter when it is done (per-syndrome early
// exit, bp.hpp:300-308).  Groups synchronise only with themselves through named barriers (bar.sync id, T /
// barrier.red.or for the convergence vote), so a group stuck on a 50-iteration non-converger does not hold up
// its neighbours.
//
// Shared-memory placement of the messages.  Thread i owns row i in the check pass, thread j owns column j in the
// bit pass; the update is in place (b2c -> c2b -> b2c) like in the streaming family.  The position of every
// message comes from two ELL tables, row_pos[k*M + i] (k-th edge of row i, ascending column = the reference's
// iterate_row order) and col_pos[k*N + j] (k-th edge of column j, ascending row).  The host chooses the positions by
// 16-colouring the edges of the (row half-warp, slot) x (column half-warp, slot) incidence graph (Koenig), colour
// = shared-memory bank pair, so that both passes are free of bank conflicts (bp_plan.cpp, build_smem_plan).
// Hard decisions are ballot words (one bit per column).  The candidate syndrome (bp.hpp:290-300) is accumulated the
// way the reference does it, from the columns: a bit decided 1 XORs its checks into a word array preset to the
// syndrome (shared-memory atomicXor; only ~p*n bits are 1), and an OR-reduction barrier tests it for zero.
#pragma once
#include "bp_smem_params.h"
#include "bp_update.cuh"

namespace bpb {

// entry `k` of row / column x in a pair-packed 16-bit table
__device__ __forceinline__ uint32_t slot16(const uint32_t *tab, int stride, int x, int k) {
    return (tab[(k >> 1) * stride + x] >> ((k & 1) * 16)) & 0xffffu;
}

__device__ __forceinline__ void group_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ bool group_any(int id, int count, bool pred) {
    uint32_t out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(out)
        : "r"((uint32_t) pred), "r"(id), "r"(count)
        : "memory");
    return out != 0;
}

// UNI: every row has exactly DC entries and every column exactly DV (regular codes): no degree tables, no
// predication.  Otherwise DC / DV are upper bounds and each slot is predicated on the actual degree.
template <int METHOD, int DC, int DV, bool LLR, int MAXT, bool UNI>
__global__ void __launch_bounds__(MAXT, 1) bp_smem_kernel(const SmemParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tab);
        uint4 *dst = reinterpret_cast<uint4 *>(sm);
        for (uint32_t i = threadIdx.x; i < p.tab_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int T = p.T;
    const int g = threadIdx.x / T;
    const int t = threadIdx.x - g * T;
    const int bar = g + 1;  // named barrier of this group (0 is the CTA-wide one used above)
    const int m = p.m, n = p.n, M = p.M, N = p.N;
    const uint8_t *row_deg = sm + p.off_row_deg;
    const uint8_t *col_deg = sm + p.off_col_deg;
    // 16-bit tables, slots (2q, 2q+1) of one row / column packed in the 32-bit word tab[q*stride + x]
    const uint32_t *col_row = reinterpret_cast<const uint32_t *>(sm + p.off_col_row);
    const uint32_t *row_pos = reinterpret_cast<const uint32_t *>(sm + p.off_row_pos);
    const uint32_t *col_pos = reinterpret_cast<const uint32_t *>(sm + p.off_col_pos);
    const double *prior = reinterpret_cast<const double *>(sm + p.off_prior);
    uint8_t *garea = sm + p.tab_bytes + (size_t) g * p.group_bytes;
    double *msg = reinterpret_cast<double *>(garea + p.goff_msg);
    uint32_t *dec = reinterpret_cast<uint32_t *>(garea + p.goff_dec);  // hard decisions, one bit per column
    uint32_t *synw = reinterpret_cast<uint32_t *>(garea + p.goff_syn);  // packed syndrome, then the candidate
    uint32_t *acc = synw + p.MW;                                        // syndrome accumulator (XOR)
    volatile long long *ctl = reinterpret_cast<volatile long long *>(garea + p.goff_ctl);

    const long long limit = p.batch_dev ? (long long) *p.batch_dev : p.batch;
    for (;;) {
        if (t == 0) {
            const long long claim = (long long) atomicAdd(p.counter, 1ull);
            ctl[0] = (claim < limit) ? (p.index_list ? (long long) p.index_list[claim] : claim) : -1;
        }
        group_sync(bar, T);
        const long long idx = ctl[0];
        if (idx < 0) break;
        // syndrome bits and initialise_log_domain_bp (bp.hpp:147-157)
        const uint32_t *srow = p.synd_packed + idx * p.mwp;
        for (int w = t; w < p.MW; w += T) synw[w] = __ldg(srow + w);
        for (int j = t; j < n; j += T) {
            const int deg = UNI ? DV : col_deg[j];
            const double pr = p.uniform_prior ? p.prior0 : prior[j];
            for (int k = 0; k < deg; ++k) msg[slot16(col_pos, N, j, k)] = pr;
        }
        group_sync(bar, T);

        int it = 0;
        bool conv = false;
        while (it < p.max_iter) {
            ++it;
            const double alpha = ms_alpha(p.ms_scaling, it);
            // ---- check -> bit, one thread per row (bp.hpp:201-273) ----
            for (int w = t; w < p.MW; w += T) acc[w] = synw[w];  // candidate ^ syndrome, must end up all zero
            for (int i = t; i < m; i += T) {
                const int deg = UNI ? DC : row_deg[i];
                uint32_t rp[DC];
                double b[DC], c[DC];
#pragma unroll
                for (int q = 0; q < (DC + 1) / 2; ++q) {
                    const uint32_t w = (2 * q < deg) ? row_pos[q * M + i] : 0u;
                    rp[2 * q] = w & 0xffffu;
                    if (2 * q + 1 < DC) rp[2 * q + 1] = w >> 16;
                }
#pragma unroll
                for (int k = 0; k < DC; ++k) b[k] = (k < deg) ? msg[rp[k]] : 0.0;
                check_node_update<METHOD, DC>(b, deg, (synw[i >> 5] >> (i & 31)) & 1u, alpha, c);
#pragma unroll
                for (int k = 0; k < DC; ++k)
                    if (k < deg) msg[rp[k]] = c[k];
            }
            group_sync(bar, T);
            // ---- posterior, decision, bit -> check, one thread per column (bp.hpp:276-318) ----
            for (int j0 = 0; j0 < N; j0 += T) {  // N is a multiple of 32: whole warps are in or out
                const int j = j0 + t;
                bool x = false;
                if (j < n) {
                    const int deg = UNI ? DV : col_deg[j];
                    uint32_t pos[DV];
                    double c[DV];
#pragma unroll
                    for (int q = 0; q < (DV + 1) / 2; ++q) {
                        const uint32_t w = (2 * q < deg) ? col_pos[q * N + j] : 0u;
                        pos[2 * q] = w & 0xffffu;
                        if (2 * q + 1 < DV) pos[2 * q + 1] = w >> 16;
                    }
#pragma unroll
                    for (int k = 0; k < DV; ++k) c[k] = (k < deg) ? msg[pos[k]] : 0.0;
                    const double llr = bit_node_update<DV>(c, deg, p.uniform_prior ? p.prior0 : prior[j]);
#pragma unroll
                    for (int k = 0; k < DV; ++k)
                        if (k < deg) msg[pos[k]] = c[k];
                    x = (llr <= 0);
                    if (LLR && (!p.llr_last_only || it == p.max_iter)) p.out_llr[idx * n + j] = llr;
                    if (x) {
                        // bp.hpp:290-294: a decided-1 bit flips the candidate syndrome of its checks
#pragma unroll
                        for (int k = 0; k < DV; ++k)
                            if (k < deg) {
                                const uint32_t r = slot16(col_row, N, j, k);
                                atomicXor(&acc[r >> 5], 1u << (r & 31));
                            }
                    }
                }
                if (j0 + (t & ~31) < N) {
                    const uint32_t word = __ballot_sync(0xffffffffu, x);
                    if ((t & 31) == 0) dec[j >> 5] = word;
                }
            }
            group_sync(bar, T);
            // ---- candidate syndrome == syndrome ?  (bp.hpp:292-308) ----
            uint32_t bad = 0;
            for (int w = t; w < p.MW; w += T) bad |= acc[w];
            conv = !group_any(bar, T, bad != 0);
            if (conv) break;
        }
        // ---- retire ----
        uint8_t *drow = p.out_dec + idx * n;
        if ((n & 3) == 0) {
            uint32_t *o32 = reinterpret_cast<uint32_t *>(drow);
            for (int w = t; w < (n >> 2); w += T) {
                const uint32_t bits = dec[w >> 3] >> ((w & 7) * 4);
                o32[w] = (bits & 1u) | ((bits & 2u) << 7) | ((bits & 4u) << 14) | ((bits & 8u) << 21);
            }
        } else {
            for (int j = t; j < n; j += T) drow[j] = (uint8_t) ((dec[j >> 5] >> (j & 31)) & 1u);
        }
        if (t == 0) {
            if (p.out_iters) p.out_iters[idx] = it;
            if (p.out_conv) p.out_conv[idx] = conv ? 1 : 0;
        }
    }
}

template <int METHOD>
SmemKernel pick_smem_bucket(int dc, int dv, bool regular, bool llr) {
#define BPB_PICK(DC_, DV_, MAXT_, UNI_)                                 \
    return llr ? bp_smem_kernel<METHOD, DC_, DV_, true, MAXT_, UNI_>    \
               : bp_smem_kernel<METHOD, DC_, DV_, false, MAXT_, UNI_>
    constexpr int SMALL = (METHOD == kMinimumSum) ? 1024 : 512;  // must agree with smem_cta_threads()
    if (regular && dc == 6 && dv == 3) { BPB_PICK(6, 3, SMALL, true); }  // (3,6)-regular LDPC, bivariate bicycle
    if (dc <= 8 && dv <= 4) { BPB_PICK(8, 4, SMALL, false); }
    if (dc <= 8 && dv <= 16) { BPB_PICK(8, 16, 512, false); }
    if (dc <= 32 && dv <= 4) { BPB_PICK(32, 4, 512, false); }
    if (dc <= 32 && dv <= 16) { BPB_PICK(32, 16, 512, false); }
#undef BPB_PICK
    return nullptr;
}

}  // namespace bpb
