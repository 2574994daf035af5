// bp_stream.cuh -- the HBM-streaming kernel family ("one lane = one syndrome").
//
// Replaces ldpc::bp::BpDecoder::bp_decode_parallel (reference src_cpp/bp.hpp:192-325) and
// bp_decode_serial (bp.hpp:451-545) for a whole batch of syndromes sharing one read-only H.
//
// Layout in HBM.  A warp owns a *tile* of 32 in-flight syndromes.  The tile's messages are one
// array msg[e][lane] of doubles, e = CSR edge id (rows ascending, columns ascending inside a row,
// i.e. the reference's iterate_row order, sparse_matrix_base.hpp:423-482), so every message access of a
// warp is one fully coalesced 256-byte transaction and a whole check row is d_c*256 contiguous bytes.
// The reference keeps two doubles per edge (bit_to_check_msg, check_to_bit_msg, bp.hpp:42-48); here
// one slot per edge suffices because each half-iteration reads a row (column) completely before it
// overwrites it: the check pass turns b2c into c2b in place, the bit pass turns c2b back into b2c.
//
// Early exit.  Lanes are persistent: when a lane's syndrome converges (bp.hpp:300-308) or reaches
// maximum_iterations it writes decoding/iterations/converge(/log_prob_ratios) for that syndrome and
// claims the next unclaimed syndrome from a global counter.  Lanes of one warp may therefore be at
// different iteration numbers; the control flow only depends on H, which all lanes share.
//
// Hard decisions are exchanged as ballot words: dec_w[j] holds bit `lane` = decoding[j] of that lane's
// syndrome; syn_w[i] likewise for the syndrome.  The candidate-syndrome test of bp.hpp:292-300 then
// costs one XOR per (row, column) pair per warp instead of per syndrome.
#pragma once
#include "bp_stream_params.h"
#include "bp_update.cuh"

namespace bpb {

// How many rows / columns / serial-schedule bits a lane keeps in flight per step (loads are issued for the whole
// batch before any of it is consumed: memory-level parallelism per warp = batch * degree 256-byte transactions).
template <int DC, bool UNI> struct RowsPerBatch { static constexpr int v = UNI ? 4 : (DC <= 8 ? 2 : 1); };
template <int DV, bool UNI> struct ColsPerBatch { static constexpr int v = UNI ? 8 : (DV <= 4 ? 4 : 1); };

// UNI: every row has exactly DC entries and every column exactly DV (regular codes): no degree predication.
// Parallel schedule: two 256-thread CTAs per SM (<= 128 registers) measured best (0.87 vs 0.77 of HBM peak with one
// CTA of 168 registers); the serial schedule keeps the registers for its larger load batches.
template <int METHOD, int SCHED, int DC, int DV, bool LLR, bool UNI>
#ifndef BPB_SERIAL_MINBLOCKS
#define BPB_SERIAL_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(256, (DC <= 8 && DV <= 4) ? ((SCHED == kParallel || UNI) ? 2 : BPB_SERIAL_MINBLOCKS) : 1)
    bp_stream_kernel(const StreamParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const long long gw = (long long) blockIdx.x * wpb + wib;
    const int m = p.m, n = p.n, nnz = p.nnz;

    const uint32_t *base = p.blob;
    if (p.smem_graph) {
        for (uint32_t i = threadIdx.x; i < p.blob_words; i += blockDim.x) smem[i] = p.blob[i];
        __syncthreads();
        base = smem;
    }
    const uint32_t *row_ptr = base;
    const uint32_t *col_idx = row_ptr + (m + 1);
    const uint32_t *col_ptr = col_idx + nnz;
    const uint32_t *csc2csr = col_ptr + (n + 1);
    const uint32_t *row_idx = csc2csr + nnz;
    const double *prior = reinterpret_cast<const double *>(base + p.prior_off);
    (void) row_idx;

    uint32_t *syn_w = p.smem_syn ? (smem + p.smem_syn_off + (size_t) wib * p.m_pad) : (p.syn_w_g + gw * p.m_pad);
    uint32_t *dec_w = p.dec_w + gw * p.n_pad;
    // GEN (DC == 0): any degree.  The reference's own two-array scheme (b2c at tile, c2b behind it), its two sweeps
    // per row / column with running values instead of register arrays (bp.hpp:201-318, 484-534 literally).
    constexpr bool GEN = (DC == 0);
    double *tile = p.msg + (size_t) gw * (size_t) nnz * 32 * (GEN ? 2 : 1) + lane;
    double *c2b = tile + (size_t) nnz * 32;  // GEN only
    (void) c2b;
    double *llr_tile = LLR ? (p.llr_tile + (size_t) gw * (size_t) n * 32 + lane) : nullptr;

    long long idx = -1;  // syndrome this lane is decoding, -1 = idle
    int it = 0;
    bool exhausted = false;  // warp-uniform: the global queue is empty

    for (;;) {
        // ---------------- claim work -----------------------------------------------------------------
        const bool need = (idx < 0) && !exhausted;
        const uint32_t needmask = __ballot_sync(0xffffffffu, need);
        if (needmask) {
            unsigned long long first_idx = 0;
            if (lane == 0) first_idx = atomicAdd(p.counter, (unsigned long long) __popc(needmask));
            first_idx = __shfl_sync(0xffffffffu, first_idx, 0);
            bool fresh = false;
            if (need) {
                const long long mine = (long long) first_idx + __popc(needmask & lanemask_lt());
                if (mine < p.batch) {
                    idx = mine;
                    it = 0;
                    fresh = true;
                }
            }
            if ((long long) first_idx + __popc(needmask) >= p.batch) exhausted = true;
            const uint32_t newmask = __ballot_sync(0xffffffffu, fresh);
            if (newmask) {
                // transpose the new lanes' packed syndromes into ballot words
                const uint32_t *srow = p.synd_packed + (fresh ? idx : 0) * p.mwp;
                for (int w = 0; w * 32 < m; ++w) {
                    const uint32_t v = fresh ? __ldg(srow + w) : 0u;
                    uint32_t mine_w = 0;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const uint32_t t = __ballot_sync(0xffffffffu, (v >> r) & 1u);
                        if (lane == r) mine_w = t;
                    }
                    const int i = w * 32 + lane;
                    if (i < m) syn_w[i] = (syn_w[i] & ~newmask) | (mine_w & newmask);
                }
                if ((SCHED == kSerial && !p.serial_no_init) || GEN) {
                    // explicit initialise_log_domain_bp (bp.hpp:147-157) for the new lanes.  A lone fresh lane writes
                    // 8 bytes of a 32-byte sector: the memory system turns that into a sector read + a sector write,
                    // 8x the payload, which is why the regular-code serial program avoids it (serial_no_init below)
                    if (fresh)
                        for (int e = 0; e < nnz; ++e)
                            st_msg(tile + (size_t) e * 32, p.uniform_prior ? p.prior0 : prior[col_idx[e]]);
                }
                __syncwarp();
            }
        }
        const bool active = idx >= 0;
        const uint32_t actmask = __ballot_sync(0xffffffffu, active);
        if (!actmask) break;
        it += 1;
        const bool first = (it == 1);
        const bool any_first = __any_sync(0xffffffffu, active && first);
        const double alpha = ms_alpha(p.ms_scaling, it);
        (void) alpha;
        (void) any_first;

        if constexpr (GEN) {
            if (SCHED == kParallel) {
                for (int i = 0; i < m; ++i) {  // bp.hpp:201-273, two sweeps per row
                    const uint32_t rb = row_ptr[i], re = row_ptr[i + 1];
                    const uint32_t s = (syn_w[i] >> lane) & 1u;
                    if (METHOD == kMinimumSum) {
                        uint32_t tsgn = s;
                        double temp = DBL_MAX;
                        for (uint32_t e = rb; e < re; ++e) {
                            const double b = active ? ld_msg(tile + (size_t) e * 32) : 0.0;
                            if (b <= 0) tsgn += 1;
                            if (active) st_msg(c2b + (size_t) e * 32, temp);
                            const double a = fabs(b);
                            if (a < temp) temp = a;
                        }
                        temp = DBL_MAX;
                        for (uint32_t e = re; e-- > rb;) {
                            const double b = active ? ld_msg(tile + (size_t) e * 32) : 0.0;
                            double c = active ? ld_msg(c2b + (size_t) e * 32) : 0.0;
                            const uint32_t sg = tsgn + ((b <= 0) ? 1u : 0u);
                            if (temp < c) c = temp;
                            c *= (sg & 1u) ? -alpha : alpha;
                            if (active) st_msg(c2b + (size_t) e * 32, c);
                            const double a = fabs(b);
                            if (a < temp) temp = a;
                        }
                    } else {
                        double temp = 1.0;
                        for (uint32_t e = rb; e < re; ++e) {
                            const double b = active ? ld_msg(tile + (size_t) e * 32) : 0.0;
                            if (active) st_msg(c2b + (size_t) e * 32, temp);
                            temp *= ps_tanh_half(b);
                        }
                        temp = 1.0;
                        const double sigma = s ? -1.0 : 1.0;
                        for (uint32_t e = re; e-- > rb;) {
                            const double b = active ? ld_msg(tile + (size_t) e * 32) : 0.0;
                            double c = active ? ld_msg(c2b + (size_t) e * 32) : 0.0;
                            c *= temp;
                            c = sigma * ps_atanh2(c);
                            if (active) st_msg(c2b + (size_t) e * 32, c);
                            temp *= ps_tanh_half(b);
                        }
                    }
                }
                uint32_t acc_w = 0;
                for (int j = 0; j < n; ++j) {  // bp.hpp:277-298 and 312-318
                    const uint32_t cb = col_ptr[j], ce = col_ptr[j + 1];
                    double t = p.uniform_prior ? p.prior0 : prior[j];
                    for (uint32_t q = cb; q < ce; ++q) {
                        const size_t e = (size_t) csc2csr[q] * 32;
                        if (active) st_msg(tile + e, t);
                        t += active ? ld_msg(c2b + e) : 0.0;
                    }
                    if (LLR) {
                        if (active) llr_tile[(size_t) j * 32] = t;
                    }
                    const uint32_t W = __ballot_sync(0xffffffffu, active && (t <= 0));
                    if (lane == (j & 31)) acc_w = W;
                    if ((j & 31) == 31 || j == n - 1) dec_w[(j & ~31) + lane] = acc_w;
                    double u = 0;
                    for (uint32_t q = ce; q-- > cb;) {
                        const size_t e = (size_t) csc2csr[q] * 32;
                        if (active) st_msg(tile + e, ld_msg(tile + e) + u);
                        u += active ? ld_msg(c2b + e) : 0.0;
                    }
                }
            } else {
                for (int oi = 0; oi < p.order_len; ++oi) {  // bp.hpp:484-534, one bit after the other
                    const uint32_t j = p.order[oi];
                    if (j == 0xffffffffu) continue;
                    const uint32_t cb = col_ptr[j], ce = col_ptr[j + 1];
                    double L = p.uniform_prior ? p.prior0 : prior[j];
                    for (uint32_t q = cb; q < ce; ++q) {
                        const uint32_t e = csc2csr[q], i = row_idx[q];
                        const uint32_t rb = row_ptr[i], re = row_ptr[i + 1];
                        const uint32_t s = (syn_w[i] >> lane) & 1u;
                        double c;
                        if (METHOD == kMinimumSum) {
                            uint32_t sg = s;
                            double temp = DBL_MAX;
                            for (uint32_t f = rb; f < re; ++f) {
                                if (f == e) continue;
                                const double b = active ? ld_msg(tile + (size_t) f * 32) : 0.0;
                                const double a = fabs(b);
                                if (a < temp) temp = a;
                                if (b <= 0) sg += 1;
                            }
                            c = ((sg & 1u) ? -alpha : alpha) * temp;
                        } else {
                            double x = 1.0;
                            for (uint32_t f = rb; f < re; ++f) {
                                if (f == e) continue;
                                x *= ps_tanh_half(active ? ld_msg(tile + (size_t) f * 32) : 0.0);
                            }
                            c = (s ? -1.0 : 1.0) * ps_atanh2(x);
                        }
                        if (active) {
                            st_msg(c2b + (size_t) e * 32, c);
                            st_msg(tile + (size_t) e * 32, L);
                        }
                        L += c;
                    }
                    if (LLR) {
                        if (active) llr_tile[(size_t) j * 32] = L;
                    }
                    const uint32_t W = __ballot_sync(0xffffffffu, active && (L <= 0));
                    if (lane == 0) dec_w[j] = W;
                    double u = 0;
                    for (uint32_t q = ce; q-- > cb;) {
                        const size_t e = (size_t) csc2csr[q] * 32;
                        if (active) st_msg(tile + e, ld_msg(tile + e) + u);
                        u += active ? ld_msg(c2b + e) : 0.0;
                    }
                }
            }
        } else
        if (SCHED == kParallel) {
            // ---------------- check -> bit (bp.hpp:201-273), in place --------------------------------
            constexpr int RB = RowsPerBatch<DC, UNI>::v;
            for (int i0 = 0; i0 < m; i0 += RB) {
                uint32_t beg[RB];
                int deg[RB];
                double b[RB][DC];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int i = i0 + r;
                    if (UNI) {
                        beg[r] = (uint32_t) (i < m ? i : 0) * DC;
                        deg[r] = (i < m) ? DC : 0;
                    } else if (i < m) {
                        beg[r] = row_ptr[i];
                        deg[r] = (int) (row_ptr[i + 1] - beg[r]);
                    } else {
                        beg[r] = 0;
                        deg[r] = 0;
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
#pragma unroll
                    for (int k = 0; k < DC; ++k) {
                        double v = p.prior0;
                        if (k < deg[r]) {
                            if (any_first && !p.uniform_prior) v = prior[col_idx[beg[r] + k]];
                            if (active && !first) v = ld_msg(tile + (size_t) (beg[r] + k) * 32);
                        }
                        b[r][k] = v;
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int i = i0 + r;
                    if (i >= m) continue;
                    const uint32_t s = (syn_w[i] >> lane) & 1u;
                    double c[DC];
                    check_node_update<METHOD, DC>(b[r], UNI ? DC : deg[r], s, alpha, c);
#pragma unroll
                    for (int k = 0; k < DC; ++k)
                        if (active && k < deg[r]) st_msg(tile + (size_t) (beg[r] + k) * 32, c[k]);
                }
            }
            // ---------------- bit pass: posterior, decision, b2c (bp.hpp:276-318), in place ----------
            constexpr int CB = ColsPerBatch<DV, UNI>::v;
            uint32_t acc_w = 0;
            for (int j0 = 0; j0 < n; j0 += CB) {
                uint32_t eid[CB][DV];
                int deg[CB];
                double c[CB][DV];
#pragma unroll
                for (int r = 0; r < CB; ++r) {
                    const int j = j0 + r;
                    uint32_t beg = 0;
                    deg[r] = 0;
                    if (j < n) {
                        beg = UNI ? (uint32_t) j * DV : col_ptr[j];
                        deg[r] = UNI ? DV : (int) (col_ptr[j + 1] - beg);
                    }
#pragma unroll
                    for (int k = 0; k < DV; ++k) eid[r][k] = (k < deg[r]) ? csc2csr[beg + k] : 0u;
                }
#pragma unroll
                for (int r = 0; r < CB; ++r) {
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        double v = 0.0;
                        if (active && k < deg[r]) v = ld_msg(tile + (size_t) eid[r][k] * 32);
                        c[r][k] = v;
                    }
                }
#pragma unroll
                for (int r = 0; r < CB; ++r) {
                    const int j = j0 + r;
                    if (j >= n) continue;
                    const double t = bit_node_update<DV>(c[r], UNI ? DV : deg[r], p.uniform_prior ? p.prior0 : prior[j]);
                    const bool x = (t <= 0);
                    if (LLR) {
                        if (active) llr_tile[(size_t) j * 32] = t;
                    }
                    const uint32_t W = __ballot_sync(0xffffffffu, active && x);
                    if (lane == (j & 31)) acc_w = W;
                    if ((j & 31) == 31 || j == n - 1) dec_w[(j & ~31) + lane] = acc_w;
#pragma unroll
                    for (int k = 0; k < DV; ++k)
                        if (active && k < deg[r]) st_msg(tile + (size_t) eid[r][k] * 32, c[r][k]);
                }
            }
        } else {
            // ---------------- serial schedule (bp.hpp:484-534) ----------------------------------------
            // The reference updates the bits one after another.  Two bits that share no check touch disjoint
            // messages (bit j reads the b2c of the OTHER edges of its checks and writes its own edges), so they
            // commute; the host sorts the schedule into levels (level(j) > level of every earlier bit that shares a
            // check with j, bp_plan.cpp: build_serial_batches) and hands the kernel batches of SB bits of one level.
            // All loads of a batch are issued before any of its stores, which multiplies the transactions in
            // flight per warp by SB; the result is the reference's, bit for bit.
            constexpr int SB = SerialBatch<DC, DV, UNI>::v;
            if (UNI) {
                // Regular codes: the host compiles the levelised schedule into a program of one 128-bit word per
                // (bit, padding included): {j, w_0, w_1, w_2,...}, w_k = row index (20 bits) | "the f-th other edge of the
                // row was already visited in this sweep" (bit 20 + f) | self position << 28, so the
                // only dependent step between the (sequential, warp-uniform) program fetch and the message loads
                // is address arithmetic.  Of a row's DC contiguous messages the DC-1 that are not the bit's own
                // are loaded: offset f + (f >= self).
                static_assert(!UNI || DV <= 3, "program word layout holds three edges");
                const uint4 *prog = reinterpret_cast<const uint4 *>(p.order);
                const bool fresh_flags = p.serial_no_init != 0;
                for (int o0 = 0; o0 < p.order_len; o0 += SB) {
                    uint4 pw[SB];
                    double bv[SB][DV][DC - 1];
#pragma unroll
                    for (int q = 0; q < SB; ++q) pw[q] = __ldg(prog + o0 + q);
#pragma unroll
                    for (int q = 0; q < SB; ++q) {
                        const uint32_t wk[3] = {pw[q].y, pw[q].z, pw[q].w};
#pragma unroll
                        for (int k = 0; k < DV; ++k) {
                            const uint32_t rb = (wk[k] & kProgRowMask) * DC, sp = wk[k] >> 28;
#pragma unroll
                            for (int f = 0; f < DC - 1; ++f) {
                                // In a syndrome's first iteration a neighbour that has not been visited yet still holds
                                // its prior (bp.hpp:147-157 set every message to it): the program says which ones have
                                // (bit 20+f), so nothing is initialised in memory and nothing stale is read.
                                const uint32_t e = rb + f + (f >= (int) sp ? 1 : 0);
                                const bool written = !(fresh_flags && first) || ((wk[k] >> (20 + f)) & 1u);
                                double v = p.prior0;  // serial_no_init is only set for a uniform prior
                                if (active && pw[q].x != 0xffffffffu && written) v = ld_msg(tile + (size_t) e * 32);
                                bv[q][k][f] = v;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < SB; ++q) {
                        if (pw[q].x == 0xffffffffu) continue;  // padding at the end of a level (warp-uniform)
                        const uint32_t j = pw[q].x;
                        const uint32_t wk[3] = {pw[q].y, pw[q].z, pw[q].w};
                        double c[DV];
#pragma unroll
                        for (int k = 0; k < DV; ++k) {
                            const uint32_t i = wk[k] & kProgRowMask;
                            const uint32_t s = (syn_w[i] >> lane) & 1u;
                            if (METHOD == kMinimumSum) {
                                uint32_t sg = s;  // bp.hpp:503-519
                                double temp = DBL_MAX;
#pragma unroll
                                for (int f = 0; f < DC - 1; ++f) {
                                    const double a = fabs(bv[q][k][f]);
                                    if (a < temp) temp = a;
                                    if (bv[q][k][f] <= 0) sg += 1;
                                }
                                c[k] = ((sg & 1u) ? -alpha : alpha) * temp;
                            } else {
                                double x = 1.0;  // bp.hpp:489-498
#pragma unroll
                                for (int f = 0; f < DC - 1; ++f) x *= ps_tanh_half(bv[q][k][f]);
                                c[k] = (s ? -1.0 : 1.0) * ps_atanh2(x);
                            }
                        }
                        const double L = bit_node_update<DV>(c, DV, p.uniform_prior ? p.prior0 : prior[j]);
                        if (LLR) {
                            if (active) llr_tile[(size_t) j * 32] = L;
                        }
                        const uint32_t W = __ballot_sync(0xffffffffu, active && (L <= 0));
                        if (lane == 0) dec_w[j] = W;
#pragma unroll
                        for (int k = 0; k < DV; ++k) {
                            const uint32_t e = (wk[k] & kProgRowMask) * DC + (wk[k] >> 28);
                            if (active) st_msg(tile + (size_t) e * 32, c[k]);
                        }
                    }
                }
            } else
            for (int o0 = 0; o0 < p.order_len; o0 += SB) {
                uint32_t jj[SB], eid[SB][DV], rbeg[SB][DV];
                int cdeg[SB], rdeg[SB][DV];
                double bv[SB][DV][DC];
#pragma unroll
                for (int q = 0; q < SB; ++q) {
                    jj[q] = p.order[o0 + q];
                    const bool valid = jj[q] != 0xffffffffu;
                    const uint32_t j = valid ? jj[q] : 0u;
                    const uint32_t cbeg = UNI ? j * DV : col_ptr[j];
                    cdeg[q] = valid ? (UNI ? DV : (int) (col_ptr[j + 1] - cbeg)) : 0;
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        eid[q][k] = 0;
                        rbeg[q][k] = 0;
                        rdeg[q][k] = 0;
                        if (k < cdeg[q]) {
                            eid[q][k] = csc2csr[cbeg + k];
                            const uint32_t i = row_idx[cbeg + k];
                            rbeg[q][k] = UNI ? i * DC : row_ptr[i];
                            rdeg[q][k] = UNI ? DC : (int) (row_ptr[i + 1] - rbeg[q][k]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < SB; ++q)
#pragma unroll
                    for (int k = 0; k < DV; ++k)
#pragma unroll
                        for (int f = 0; f < DC; ++f) {
                            double v = 0.0;
                            if (active && f < rdeg[q][k] && rbeg[q][k] + f != eid[q][k])
                                v = ld_msg(tile + (size_t) (rbeg[q][k] + f) * 32);
                            bv[q][k][f] = v;
                        }
#pragma unroll
                for (int q = 0; q < SB; ++q) {
                    if (jj[q] == 0xffffffffu) continue;  // padding at the end of a level (warp-uniform)
                    const uint32_t j = jj[q];
                    const uint32_t cbeg = UNI ? j * DV : col_ptr[j];
                    double c[DV];
#pragma unroll
                    for (int k = 0; k < DV; ++k) {
                        if (k < cdeg[q]) {
                            const uint32_t i = row_idx[cbeg + k];
                            const uint32_t s = (syn_w[i] >> lane) & 1u;
                            if (METHOD == kMinimumSum) {
                                // bp.hpp:503-519
                                uint32_t sg = s;
                                double temp = DBL_MAX;
#pragma unroll
                                for (int f = 0; f < DC; ++f) {
                                    if (f < rdeg[q][k] && rbeg[q][k] + f != eid[q][k]) {
                                        const double a = fabs(bv[q][k][f]);
                                        if (a < temp) temp = a;
                                        if (bv[q][k][f] <= 0) sg += 1;
                                    }
                                }
                                c[k] = ((sg & 1u) ? -alpha : alpha) * temp;
                            } else {
                                // bp.hpp:489-498
                                double x = 1.0;
#pragma unroll
                                for (int f = 0; f < DC; ++f)
                                    if (f < rdeg[q][k] && rbeg[q][k] + f != eid[q][k]) x *= ps_tanh_half(bv[q][k][f]);
                                c[k] = (s ? -1.0 : 1.0) * ps_atanh2(x);
                            }
                        }
                    }
                    // bp.hpp:499-500 / 520-533: b2c := running sum, posterior, decision, extrinsic b2c
                    const double L = bit_node_update<DV>(c, cdeg[q], p.uniform_prior ? p.prior0 : prior[j]);
                    const bool x = (L <= 0);
                    if (LLR) {
                        if (active) llr_tile[(size_t) j * 32] = L;
                    }
                    const uint32_t W = __ballot_sync(0xffffffffu, active && x);
                    if (lane == 0) dec_w[j] = W;
#pragma unroll
                    for (int k = 0; k < DV; ++k)
                        if (active && k < cdeg[q]) st_msg(tile + (size_t) eid[q][k] * 32, c[k]);
                }
            }
        }
        __syncwarp();

        // ---------------- candidate syndrome == syndrome ?  (bp.hpp:292-308 / 537-542) ---------------
        uint32_t mis = 0;
        for (int i = lane; i < m; i += 32) {
            uint32_t cs = 0;
            const uint32_t rb = row_ptr[i], re = row_ptr[i + 1];
            for (uint32_t e = rb; e < re; ++e) cs ^= __ldcg(dec_w + col_idx[e]);
            mis |= cs ^ syn_w[i];
        }
        mis = __reduce_or_sync(0xffffffffu, mis);
        const bool conv = active && !((mis >> lane) & 1u);
        const bool done = active && (conv || it >= p.max_iter);
        uint32_t donemask = __ballot_sync(0xffffffffu, done);
        // Ramp-down: once the queue is empty, a lane that is still iterating after iter_cap iterations is a
        // straggler that would keep its whole warp alive for up to maximum_iterations latency-bound sweeps.  Hand
        // its syndrome to the second stage (decoded from scratch by the thread-group kernel; BP is deterministic,
        // so the result is the same) instead of finishing it here.
        const bool defer = active && !done && exhausted && it >= p.iter_cap;
        const uint32_t defmask = __ballot_sync(0xffffffffu, defer);
        if (defmask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.handoff_count, (unsigned long long) __popc(defmask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (defer) p.handoff_list[base + __popc(defmask & lanemask_lt())] = (uint32_t) idx;
        }

        // ---------------- retire finished syndromes ------------------------------------------------
        while (donemask) {
            const int l = __ffs(donemask) - 1;
            donemask &= donemask - 1;
            const long long oidx = __shfl_sync(0xffffffffu, idx, l);
            const int oit = __shfl_sync(0xffffffffu, it, l);
            const int oconv = __shfl_sync(0xffffffffu, (int) conv, l);
            uint8_t *drow = p.out_dec + oidx * n;
            if ((n & 3) == 0) {
                for (int j = lane * 4; j < n; j += 128) {
                    const uint4 w4 = __ldcg(reinterpret_cast<const uint4 *>(dec_w + j));
                    const uint32_t packed = ((w4.x >> l) & 1u) | (((w4.y >> l) & 1u) << 8) |
                                            (((w4.z >> l) & 1u) << 16) | (((w4.w >> l) & 1u) << 24);
                    *reinterpret_cast<uint32_t *>(drow + j) = packed;
                }
            } else {
                for (int j = lane; j < n; j += 32) drow[j] = (uint8_t) ((__ldcg(dec_w + j) >> l) & 1u);
            }
            if (lane == 0) {
                if (p.out_iters) p.out_iters[oidx] = oit;
                if (p.out_conv) p.out_conv[oidx] = (uint8_t) oconv;
            }
            if (LLR && !(p.llr_last_only && oconv)) {
                const double *src = p.llr_tile + (size_t) gw * (size_t) n * 32 + l;
                double *dst = p.out_llr + oidx * n;
                for (int j = lane; j < n; j += 32) dst[j] = __ldcg(src + (size_t) j * 32);
            }
        }
        if (done || defer) {
            atomicAdd(p.iter_total, (unsigned long long) it);  // iterations this kernel really executed (roofline)
            idx = -1;
        }

        // ---------------- ramp-down: compact the live lanes of the warp ---------------------------------
        // Once the queue is empty, lanes that finish leave holes.  A message row of the tile is 8 DRAM sectors of 4
        // lanes each, and a sector with one live lane is fetched and written whole: ncu at n = 10^4 shows the memory
        // system saturated during the ramp-down while a third of its traffic is dead lanes.  When the live lanes
        // would fit in at most half of the sectors they occupy, move them to lanes 0..k-1: per-lane registers and the
        // syndrome ballot words by shuffle / bit permutation, the message columns by one read + one write of the tile
        // (a third of a serial-schedule iteration).  Warp-local and exact: only where a syndrome's state lives changes.
        if (exhausted && !p.no_compaction) {
            const uint32_t live = __ballot_sync(0xffffffffu, idx >= 0);
            const int k = __popc(live);
            uint32_t quads = live | (live >> 1) | (live >> 2) | (live >> 3);
            quads &= 0x11111111u;
            const int occupied = __popc(quads);
            if (k > 0 && ((k + 3) >> 2) * p.compact_num <= occupied * p.compact_den) {
                const bool was_live = idx >= 0;
                const int src = (lane < k) ? (int) __fns(live, 0, lane + 1) : lane;  // old lane of new lane `lane`
                const long long nidx = __shfl_sync(0xffffffffu, idx, src);
                const int nit = __shfl_sync(0xffffffffu, it, src);
                idx = (lane < k) ? nidx : -1;
                it = (lane < k) ? nit : 0;
                for (int i0 = 0; i0 < p.m_pad; i0 += 32) {
                    const uint32_t w = syn_w[i0 + lane];
                    uint32_t nw = 0;
                    for (int q = 0; q < k; ++q) {
                        const int sq = __shfl_sync(0xffffffffu, src, q);
                        nw |= ((w >> sq) & 1u) << q;
                    }
                    syn_w[i0 + lane] = nw;
                }
                constexpr int U = 8;
                const int words = nnz * (GEN ? 2 : 1);  // the generic kernel keeps c2b right behind b2c
                for (int e0 = 0; e0 < words; e0 += U) {
                    double v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        v[u] = (was_live && e0 + u < words) ? ld_msg(tile + (size_t) (e0 + u) * 32) : 0.0;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const double nv = __shfl_sync(0xffffffffu, v[u], src);
                        if (lane < k && e0 + u < words) st_msg(tile + (size_t) (e0 + u) * 32, nv);
                    }
                }
                __syncwarp();
            }
        }
    }
}

// One translation unit per (method, schedule) instantiates its degree buckets through this helper.
template <int METHOD, int SCHED>
StreamKernel pick_stream_bucket(int dc, int dv, bool regular, bool llr) {
#define BPB_PICK(DC_, DV_, UNI_)                                                   \
    return llr ? bp_stream_kernel<METHOD, SCHED, DC_, DV_, true, UNI_>             \
               : bp_stream_kernel<METHOD, SCHED, DC_, DV_, false, UNI_>
    if (regular && dc == 6 && dv == 3) { BPB_PICK(6, 3, true); }  // (3,6)-regular LDPC, bivariate bicycle
    if (dc <= 8 && dv <= 4) { BPB_PICK(8, 4, false); }
    if (dc <= 8 && dv <= 16) { BPB_PICK(8, 16, false); }
    if (dc <= 32 && dv <= 4) { BPB_PICK(32, 4, false); }
    if (dc <= 32 && dv <= 16) { BPB_PICK(32, 16, false); }
    BPB_PICK(0, 0, false);  // any degree: the two-array scheme of the reference
#undef BPB_PICK
}

}  // namespace bpb
