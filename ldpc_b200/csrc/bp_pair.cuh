// bp_pair.cuh -- the paired on-chip kernel family ("one thread group = TWO syndromes, messages in shared memory as
// double2").
//
// Same job as bp_smem.cuh -- ldpc::bp::BpDecoder::bp_decode_parallel (reference src_cpp/bp.hpp:192-325) for codes whose
// messages fit in shared memory -- with half the instruction stream per syndrome.  The on-chip kernel is bound by
// instruction issue (profiles/r1_ncu_smem_final.json: issue slots 87 %), and about half of what it issues is not
// arithmetic: index-table loads and unpacking, address arithmetic, shared-memory load/store instructions, loop and
// barrier overhead.  All of that is per EDGE, not per syndrome.  Here a thread group decodes two syndromes at once:
// slot p of the message array is a double2 {message of syndrome A, message of syndrome B}, so one LDS.128 / STS.128 and
// one table decode serve both, and the two independent dependency chains give the scheduler instruction-level
// parallelism inside a warp.
//
// The two syndromes of a pair are independent decodes that merely share the control flow, which depends only on H:
// each has its own iteration counter, converges (bp.hpp:300-308) and retires on its own, and its half of the pair is
// refilled from the global queue while the other half keeps iterating.  Arithmetic per syndrome is exactly that of
// bp_update.cuh, so results are bit-identical to the other families.
//
// Placement.  A 16-byte shared-memory access is served per quarter-warp and is conflict-free iff its 8 lanes hit 8
// different bank quads; the host 8-colours the edges of the (row quarter-warp, slot) x (column quarter-warp, slot)
// incidence graph (bp_plan.cpp: place_messages with L = 8), colour = bank quad, so both passes are conflict-free.
//
// Barriers.  Two per iteration (check | bit | next check): after the bit pass every warp reads the candidate-syndrome
// accumulator words itself and votes with __any_sync, so no third barrier is needed.
//
// Candidate syndrome and decisions.  Hard decisions stay in registers (each thread owns its columns), and the
// accumulator  syndrome ^ H x  is kept up to date incrementally: a column XORs its checks in only when its decision
// flips (few do after the first iterations), which is the reference's candidate_syndrome (bp.hpp:290-300) by
// linearity.  Converged <=> all accumulator words are zero.
#pragma once
#include "bp_pair_params.h"
#include "bp_smem.cuh"

namespace bpb {

// 32-bit shared-window addressing with explicit 16-byte accesses: position p of the message array is the double2 at
// msg_base + 16 * p; the 16-bit table entry is extracted with one PRMT and scaled with one IMAD.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t slot_addr(uint32_t base, uint32_t word, int hi) {
    return base + __byte_perm(word, 0u, hi ? 0x4432u : 0x4410u) * 16u;
}
__device__ __forceinline__ double2 lds128(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
// shared-memory word ^= bit (reduction: no return value; 64-bit XOR reductions would be emulated with a CAS loop)
__device__ __forceinline__ void smem_xor(uint32_t addr, uint32_t bit) {
    asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(addr), "r"(bit) : "memory");
}

template <int METHOD, int DC, int DV, bool LLR, int MAXT, bool UNI>
__global__ void __launch_bounds__(MAXT, 1) bp_pair_kernel(const PairParams p) {
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
    const int lane = t & 31;
    const int bar = g + 1;  // named barrier of this group (0 is the CTA-wide one used above)
    const int m = p.m, n = p.n, M = p.M, N = p.N, MW = p.MW;
    constexpr int DCp = (DC + 1) / 2, DVp = (DV + 1) / 2;
    const uint8_t *row_deg = sm + p.off_row_deg;
    const uint8_t *col_deg = sm + p.off_col_deg;
    // 16-bit tables, slots (2q, 2q+1) of one row / column packed in the 32-bit word tab[q*stride + x]
    const uint32_t *col_row = reinterpret_cast<const uint32_t *>(sm + p.off_col_row);
    const uint32_t *row_pos = reinterpret_cast<const uint32_t *>(sm + p.off_row_pos);
    const uint32_t *col_pos = reinterpret_cast<const uint32_t *>(sm + p.off_col_pos);
    const double *prior = reinterpret_cast<const double *>(sm + p.off_prior);
    uint8_t *garea = sm + p.tab_bytes + (size_t) g * p.group_bytes;
    // per-syndrome words are interleaved {A, B}: one 64-bit access serves both halves of the pair
    const uint32_t msg_s = smem_addr(garea + p.goff_msg);         // double2 slots: .x syndrome A, .y syndrome B
    uint2 *synw = reinterpret_cast<uint2 *>(garea + p.goff_syn);  // packed syndromes [MW]{A, B}
    uint2 *acc = reinterpret_cast<uint2 *>(garea + p.goff_acc);   // candidate ^ syndrome [MW]{A, B}
    const uint32_t acc_s = smem_addr(acc);
    volatile long long *ctl = reinterpret_cast<volatile long long *>(garea + p.goff_ctl);

    const long long limit = p.batch_dev ? (long long) *p.batch_dev : p.batch;
    long long idx0 = -1, idx1 = -1;  // batch index each half is decoding, -1 = idle
    int it0 = 0, it1 = 0;
    bool need0 = true, need1 = true;  // the half wants a syndrome from the queue (group-uniform)
    // Hard decisions of this thread's columns j = r*T + t, two bits per round r: bit 2r = syndrome A, 2r+1 = B (the
    // host guarantees at most 16 rounds).  The candidate syndrome is maintained incrementally: the accumulator starts
    // as the syndrome, and a column XORs its checks in (bp.hpp:290-294) only when its decision FLIPS, so that
    // acc == syndrome ^ H x at the end of every bit pass.
    uint32_t xm = 0;
    // The group leader always holds one syndrome claimed AHEAD in a register: the global atomic (and the index-list
    // lookup of the second-stage use) is issued when a half retires but only consumed when the next half retires, so
    // its latency never sits on the group's critical path.
    auto claim_next = [&]() -> long long {
        const long long claim = (long long) atomicAdd(p.counter, 1ull);
        return (claim < limit) ? (p.index_list ? (long long) p.index_list[claim] : claim) : -1;
    };
    long long pend = -1;
    if (t == 0) pend = claim_next();
    for (;;) {
        if (need0 || need1) {
            if (t == 0) {
                if (need0) {
                    ctl[0] = pend;
                    pend = (pend >= 0) ? claim_next() : -1;
                }
                if (need1) {
                    ctl[1] = pend;
                    pend = (pend >= 0) ? claim_next() : -1;
                }
            }
            group_sync(bar, T);
            if (need0) {
                idx0 = ctl[0];
                it0 = 0;
            }
            if (need1) {
                idx1 = ctl[1];
                it1 = 0;
            }
            const bool f0 = need0 && idx0 >= 0, f1 = need1 && idx1 >= 0;
            need0 = need1 = false;
            if (idx0 < 0 && idx1 < 0) break;
            // syndrome bits, x = 0, and initialise_log_domain_bp (bp.hpp:147-157) for the fresh halves
            if (f0) {
                const uint32_t *srow = p.synd_packed + idx0 * p.mwp;
                for (int w = t; w < MW; w += T) acc[w].x = synw[w].x = __ldg(srow + w);
                xm &= 0xaaaaaaaau;
            }
            if (f1) {
                const uint32_t *srow = p.synd_packed + idx1 * p.mwp;
                for (int w = t; w < MW; w += T) acc[w].y = synw[w].y = __ldg(srow + w);
                xm &= 0x55555555u;
            }
            if ((f0 || f1) && p.uniform_prior) {
                // every message starts as the same prior: fill the slots linearly, no table lookups
                double2 *q = reinterpret_cast<double2 *>(garea + p.goff_msg);
                for (int s = t; s < p.msg_slots; s += T) {
                    if (f0) q[s].x = p.prior0;
                    if (f1) q[s].y = p.prior0;
                }
            } else if (f0 || f1) {
                for (int j = t; j < n; j += T) {
                    const int deg = UNI ? DV : col_deg[j];
                    const double pr = p.uniform_prior ? p.prior0 : prior[j];
                    for (int k = 0; k < deg; ++k) {
                        double2 *q = reinterpret_cast<double2 *>(garea + p.goff_msg) + slot16(col_pos, N, j, k);
                        if (f0) q->x = pr;
                        if (f1) q->y = pr;
                    }
                }
            }
            group_sync(bar, T);
        }
        ++it0;
        ++it1;
        const double alpha0 = ms_alpha(p.ms_scaling, it0), alpha1 = ms_alpha(p.ms_scaling, it1);
        // ---- check -> bit, one thread per row (bp.hpp:201-273); the table words of the next row are fetched while
        // this one is computed ----
        {
            uint32_t wn[DCp];
#pragma unroll
            for (int q = 0; q < DCp; ++q) wn[q] = (t < m) ? row_pos[q * M + t] : 0u;
            for (int i = t; i < m; i += T) {
                const int deg = UNI ? DC : row_deg[i];
                uint32_t rp[DC];
                double b0[DC], b1[DC], c0[DC], c1[DC];
#pragma unroll
                for (int q = 0; q < DCp; ++q) {
                    rp[2 * q] = slot_addr(msg_s, wn[q], 0);
                    if (2 * q + 1 < DC) rp[2 * q + 1] = slot_addr(msg_s, wn[q], 1);
                }
#pragma unroll
                for (int k = 0; k < DC; ++k) {
                    double2 v = make_double2(0.0, 0.0);
                    if (k < deg) v = lds128(rp[k]);
                    b0[k] = v.x;
                    b1[k] = v.y;
                }
                const uint2 sw = synw[i >> 5];
                if (i + T < m) {
#pragma unroll
                    for (int q = 0; q < DCp; ++q) wn[q] = row_pos[q * M + i + T];
                }
                const uint32_t sh = (uint32_t) i & 31u;
                check_node_update<METHOD, DC>(b0, deg, (sw.x >> sh) & 1u, alpha0, c0);
                check_node_update<METHOD, DC>(b1, deg, (sw.y >> sh) & 1u, alpha1, c1);
#pragma unroll
                for (int k = 0; k < DC; ++k)
                    if (k < deg) sts128(rp[k], c0[k], c1[k]);
            }
        }
        group_sync(bar, T);
        // ---- posterior, decision, bit -> check, one thread per column (bp.hpp:276-318) ----
        const bool llr0_on = LLR && idx0 >= 0 && (!p.llr_last_only || it0 == p.max_iter);
        const bool llr1_on = LLR && idx1 >= 0 && (!p.llr_last_only || it1 == p.max_iter);
        {
            uint32_t wn[DVp];
#pragma unroll
            for (int q = 0; q < DVp; ++q) wn[q] = (t < n) ? col_pos[q * N + t] : 0u;
            int rounds = 0;
            for (int j = t; j < n; j += T, ++rounds) {
                const int deg = UNI ? DV : col_deg[j];
                uint32_t pos[DV];
                double c0[DV], c1[DV];
#pragma unroll
                for (int q = 0; q < DVp; ++q) {
                    pos[2 * q] = slot_addr(msg_s, wn[q], 0);
                    if (2 * q + 1 < DV) pos[2 * q + 1] = slot_addr(msg_s, wn[q], 1);
                }
#pragma unroll
                for (int k = 0; k < DV; ++k) {
                    double2 v = make_double2(0.0, 0.0);
                    if (k < deg) v = lds128(pos[k]);
                    c0[k] = v.x;
                    c1[k] = v.y;
                }
                if (j + T < n) {
#pragma unroll
                    for (int q = 0; q < DVp; ++q) wn[q] = col_pos[q * N + j + T];
                }
                const double pr = p.uniform_prior ? p.prior0 : prior[j];
                const double llr0 = bit_node_update<DV>(c0, deg, pr);
                const double llr1 = bit_node_update<DV>(c1, deg, pr);
#pragma unroll
                for (int k = 0; k < DV; ++k)
                    if (k < deg) sts128(pos[k], c0[k], c1[k]);
                if (LLR) {
                    if (llr0_on) p.out_llr[idx0 * n + j] = llr0;
                    if (llr1_on) p.out_llr[idx1 * n + j] = llr1;
                }
                // decisions of this round sit in the two low bits of xm; rotate to the next round afterwards
                const uint32_t now = ((llr0 <= 0) ? 1u : 0u) | ((llr1 <= 0) ? 2u : 0u);
                const uint32_t flip = (xm ^ now) & 3u;
                xm ^= flip;
                if (flip) {  // few columns flip after the first iterations: the row indices are fetched only here
                    uint32_t cr[DVp];
#pragma unroll
                    for (int q = 0; q < DVp; ++q) cr[q] = (2 * q < deg) ? col_row[q * N + j] : 0u;
#pragma unroll
                    for (int k = 0; k < DV; ++k)
                        if (k < deg) {
                            const uint32_t r = __byte_perm(cr[k >> 1], 0u, (k & 1) ? 0x4432u : 0x4410u);
                            const uint32_t bit = 1u << (r & 31u);
                            const uint32_t a = acc_s + (r >> 5) * 8u;
                            smem_xor(a, (flip & 1u) ? bit : 0u);
                            smem_xor(a + 4u, (flip & 2u) ? bit : 0u);
                        }
                }
                xm = __funnelshift_r(xm, xm, 2);
            }
            xm = __funnelshift_l(xm, xm, 2 * rounds);  // back to round 0 in the low bits
        }
        group_sync(bar, T);
        // ---- candidate syndrome == syndrome ?  (bp.hpp:292-308); every warp reads the words itself ----
        uint32_t v0 = 0, v1 = 0;
        for (int w = lane; w < MW; w += 32) {
            const uint2 v = acc[w];
            v0 |= v.x;
            v1 |= v.y;
        }
        const bool bad0 = __any_sync(0xffffffffu, v0 != 0);
        const bool bad1 = __any_sync(0xffffffffu, v1 != 0);
        const bool done0 = idx0 >= 0 && (!bad0 || it0 >= p.max_iter);
        const bool done1 = idx1 >= 0 && (!bad1 || it1 >= p.max_iter);
        // ---- retire: every thread writes the decisions of its own columns ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!(h ? done1 : done0)) continue;
            const long long idx = h ? idx1 : idx0;
            uint8_t *drow = p.out_dec + idx * n;
            uint32_t bits = xm >> h;
            for (int j = t; j < n; j += T, bits >>= 2) drow[j] = (uint8_t) (bits & 1u);
            if (t == 0) {
                if (p.out_iters) p.out_iters[idx] = h ? it1 : it0;
                if (p.out_conv) p.out_conv[idx] = (h ? bad1 : bad0) ? 0 : 1;
            }
        }
        need0 = done0;
        need1 = done1;
    }
}

// maxt: CTA size the kernel is compiled for (__launch_bounds__): 512 threads leave 128 registers per thread, 640 leave
// 102, 768 leave 85.  The regular (3,6) bucket exists in all three; the others in 512 only.
template <int METHOD>
PairKernel pick_pair_bucket(int dc, int dv, bool regular, bool llr, int maxt) {
#define BPB_PICK(DC_, DV_, MAXT_, UNI_)                                   \
    return llr ? bp_pair_kernel<METHOD, DC_, DV_, true, MAXT_, UNI_>      \
               : bp_pair_kernel<METHOD, DC_, DV_, false, MAXT_, UNI_>
    if (regular && dc == 6 && dv == 3) {  // (3,6)-regular LDPC, bivariate bicycle
        if (maxt == 768) { BPB_PICK(6, 3, 768, true); }
        if (maxt == 640) { BPB_PICK(6, 3, 640, true); }
        if (maxt == 512) { BPB_PICK(6, 3, 512, true); }
        return nullptr;
    }
    if (maxt != 512) return nullptr;
    if (dc <= 4 && dv <= 2) { BPB_PICK(4, 2, 512, false); }  // surface codes: half the register arrays of (8, 4)
    if (dc <= 8 && dv <= 4) { BPB_PICK(8, 4, 512, false); }
    if (dc <= 8 && dv <= 16) { BPB_PICK(8, 16, 512, false); }
    if (dc <= 32 && dv <= 4) { BPB_PICK(32, 4, 512, false); }
#undef BPB_PICK
    return nullptr;
}

}  // namespace bpb
