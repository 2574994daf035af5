// bp_smem_serial.cuh -- the on-chip kernel for the SERIAL schedule (thread group = one syndrome, messages in shared
// memory), replacing ldpc::bp::BpDecoder::bp_decode_serial (reference src_cpp/bp.hpp:451-545) for codes that fit.
//
// The reference sweeps the bits one after another.  Bit j reads the bit->check messages of the OTHER edges of its
// checks and rewrites only its own edges (bp.hpp:489-533), so two bits that share no check commute.  The host
// levelises the schedule (level(q) = 1 + max level of the earlier positions whose bit shares a check; bp_capi.cu:
// build_serial_levels): within a level every bit is independent, so a thread group processes a level in parallel,
// one thread per bit, with one barrier per level.  Floating-point operations and their order per bit are the
// reference's, so the result is bit-identical (tests/test_gpu_parity.py, serial cases, both families).
//
// Only b2c is stored (one double per edge, same placement tables as bp_smem.cuh); c2b lives in registers for the
// few instructions between its computation and the bit update.  Hard decisions are bytes (different levels write
// different bits of a word); the candidate syndrome is accumulated with atomicXor as in the parallel kernel, which
// equals H * decoding after the sweep (bp.hpp:537).
#pragma once
#include "bp_smem.cuh"

namespace bpb {

template <int METHOD, int DC, int DV, bool LLR, int MAXT, bool UNI>
__global__ void __launch_bounds__(MAXT, 1) bp_smem_serial_kernel(const SmemParams p) {
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
    const int bar = g + 1;
    const int n = p.n, M = p.M, N = p.N;
    const uint8_t *row_deg = sm + p.off_row_deg;
    const uint8_t *col_deg = sm + p.off_col_deg;
    const uint32_t *col_row = reinterpret_cast<const uint32_t *>(sm + p.off_col_row);
    const uint32_t *row_pos = reinterpret_cast<const uint32_t *>(sm + p.off_row_pos);
    const uint32_t *col_pos = reinterpret_cast<const uint32_t *>(sm + p.off_col_pos);
    const uint8_t *col_self = sm + p.off_col_self;   // slot of edge (j,k) inside its row
    const uint16_t *lev_ptr = reinterpret_cast<const uint16_t *>(sm + p.off_lev_ptr);
    const uint16_t *lev_bits = reinterpret_cast<const uint16_t *>(sm + p.off_lev_bits);
    const double *prior = reinterpret_cast<const double *>(sm + p.off_prior);
    uint8_t *garea = sm + p.tab_bytes + (size_t) g * p.group_bytes;
    double *msg = reinterpret_cast<double *>(garea + p.goff_msg);
    uint8_t *dec = garea + p.goff_dec;  // hard decisions, one byte per column
    uint32_t *synw = reinterpret_cast<uint32_t *>(garea + p.goff_syn);
    uint32_t *acc = synw + p.MW;
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
        const uint32_t *srow = p.synd_packed + idx * p.mwp;
        for (int w = t; w < p.MW; w += T) synw[w] = __ldg(srow + w);
        for (int j = t; j < n; j += T) {  // initialise_log_domain_bp (bp.hpp:147-157); decoding persists like the
            const int deg = UNI ? DV : col_deg[j];  // reference's member for bits a custom order never visits
            const double pr = p.uniform_prior ? p.prior0 : prior[j];
            for (int k = 0; k < deg; ++k) msg[slot16(col_pos, N, j, k)] = pr;
            dec[j] = 0;
        }
        group_sync(bar, T);

        int it = 0;
        bool conv = false;
        while (it < p.max_iter) {
            ++it;
            const double alpha = ms_alpha(p.ms_scaling, it);
            for (int w = t; w < p.MW; w += T) acc[w] = synw[w];
            group_sync(bar, T);
            for (int lv = 0; lv < p.n_levels; ++lv) {
                const int qe = lev_ptr[lv + 1];
                for (int q = lev_ptr[lv] + t; q < qe; q += T) {
                    const int j = lev_bits[q];
                    const int deg = UNI ? DV : col_deg[j];
                    double c[DV];
                    uint32_t pe[DV], ri[DV];
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        c[k] = 0.0;
                        pe[k] = 0;
                        ri[k] = 0;
                        if (k < deg) {
                            const uint32_t i = slot16(col_row, N, j, k);
                            const int self = col_self[k * N + j];
                            const int rdeg = UNI ? DC : row_deg[i];
                            const uint32_t s = (synw[i >> 5] >> (i & 31)) & 1u;
                            ri[k] = i;
                            pe[k] = slot16(col_pos, N, j, k);
                            double bv[DC];
#pragma unroll
                            for (int h = 0; h < (DC + 1) / 2; ++h) {
                                const uint32_t w = (2 * h < rdeg) ? row_pos[h * M + i] : 0u;
                                bv[2 * h] = (2 * h < rdeg && 2 * h != self) ? msg[w & 0xffffu] : 0.0;
                                if (2 * h + 1 < DC)
                                    bv[2 * h + 1] = (2 * h + 1 < rdeg && 2 * h + 1 != self) ? msg[w >> 16] : 0.0;
                            }
                            if (METHOD == kMinimumSum) {
                                uint32_t sg = s;  // bp.hpp:503-519
                                double temp = DBL_MAX;
#pragma unroll
                                for (int f = 0; f < DC; ++f) {
                                    if (f < rdeg && f != self) {
                                        const double a = fabs(bv[f]);
                                        if (a < temp) temp = a;
                                        if (bv[f] <= 0) sg += 1;
                                    }
                                }
                                c[k] = ((sg & 1u) ? -alpha : alpha) * temp;
                            } else {
                                double x = 1.0;  // bp.hpp:489-498
#pragma unroll
                                for (int f = 0; f < DC; ++f)
                                    if (f < rdeg && f != self) x *= ps_tanh_half(bv[f]);
                                c[k] = (s ? -1.0 : 1.0) * ps_atanh2(x);
                            }
                        }
                    }
                    const double L = bit_node_update<DV>(c, deg, p.uniform_prior ? p.prior0 : prior[j]);
#pragma unroll
                    for (int k = 0; k < DV; ++k)
                        if (k < deg) msg[pe[k]] = c[k];
                    const bool x = (L <= 0);
                    if (LLR && (!p.llr_last_only || it == p.max_iter)) p.out_llr[idx * n + j] = L;
                    dec[j] = x ? 1 : 0;
                }
                group_sync(bar, T);
            }
            // ---- candidate = H * decoding (bp.hpp:537, gf2sparse.hpp:177-196), from the columns ----
            for (int j = t; j < n; j += T) {
                if (dec[j]) {
                    const int deg = UNI ? DV : col_deg[j];
                    for (int k = 0; k < deg; ++k) {
                        const uint32_t r = slot16(col_row, N, j, k);
                        atomicXor(&acc[r >> 5], 1u << (r & 31));
                    }
                }
            }
            group_sync(bar, T);
            uint32_t bad = 0;
            for (int w = t; w < p.MW; w += T) bad |= acc[w];
            conv = !group_any(bar, T, bad != 0);
            if (conv) break;
        }
        uint8_t *drow = p.out_dec + idx * n;
        if ((n & 3) == 0) {
            const uint32_t *d32 = reinterpret_cast<const uint32_t *>(dec);
            uint32_t *o32 = reinterpret_cast<uint32_t *>(drow);
            for (int w = t; w < (n >> 2); w += T) o32[w] = d32[w];
        } else {
            for (int j = t; j < n; j += T) drow[j] = dec[j];
        }
        if (t == 0) {
            if (p.out_iters) p.out_iters[idx] = it;
            if (p.out_conv) p.out_conv[idx] = conv ? 1 : 0;
        }
    }
}

template <int METHOD>
SmemKernel pick_smem_serial_bucket(int dc, int dv, bool regular, bool llr) {
#define BPB_PICK(DC_, DV_, UNI_)                                              \
    return llr ? bp_smem_serial_kernel<METHOD, DC_, DV_, true, 512, UNI_>     \
               : bp_smem_serial_kernel<METHOD, DC_, DV_, false, 512, UNI_>
    if (regular && dc == 6 && dv == 3) { BPB_PICK(6, 3, true); }
    if (dc <= 8 && dv <= 4) { BPB_PICK(8, 4, false); }
    if (dc <= 8 && dv <= 16) { BPB_PICK(8, 16, false); }
    if (dc <= 32 && dv <= 4) { BPB_PICK(32, 4, false); }
    if (dc <= 32 && dv <= 16) { BPB_PICK(32, 16, false); }
#undef BPB_PICK
    return nullptr;
}

}  // namespace bpb
