// bp_update.cuh -- the per-node arithmetic of the parallel schedule, shared by the kernel families.
//
// One check-node update = reference src_cpp/bp.hpp:201-219 (product-sum) / :220-273 (min-sum) for one row;
// one bit-node update   = bp.hpp:276-298 (posterior + hard decision) and :311-318 (extrinsic b2c) for one column.
// Values live in registers (fully unrolled over the template degree, predicated on the actual degree); the
// order of floating-point operations is the reference's.
#pragma once
#include "bp_common.cuh"

namespace bpb {

// b[0..deg) : bit->check messages of the row in ascending column order.  On return c[0..deg) holds the
// check->bit messages.  s = syndrome bit of the row.
template <int METHOD, int DC>
__device__ __forceinline__ void check_node_update(const double (&b)[DC], int deg, uint32_t s, double alpha,
                                                  double (&c)[DC]) {
    if (METHOD == kMinimumSum) {
        // The reference builds |c_k| = min_{k' != k} |b_k'| from a prefix and a suffix running minimum with
        // strict '<' updates starting at DBL_MAX (bp.hpp:237-268).  min is exact and order-free, NaNs never
        // win a '<', so the same value is min2 for the argmin edge and min1 for every other edge.
        // The same two running minima as the reference (bp.hpp:237-268): a forward one and a backward one, both
        // seeded with DBL_MAX and updated on strict '<' (a NaN or +inf magnitude never replaces the running value),
        // |c_k| = min(prefix_k, suffix_k).  Slots beyond the degree hold DBL_MAX, the neutral element.
        uint32_t tsgn = s;  // total_sgn, bp.hpp:236-242
        double a[DC], pre[DC];
        bool neg[DC];
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            neg[k] = (k < deg) && (b[k] <= 0);
            tsgn += neg[k] ? 1u : 0u;
            a[k] = (k < deg) ? fabs(b[k]) : DBL_MAX;
        }
        double run = DBL_MAX;
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            pre[k] = run;
            run = (a[k] < run) ? a[k] : run;
        }
        run = DBL_MAX;
#pragma unroll
        for (int k = DC - 1; k >= 0; --k) {
            if (k < deg) {
                const double mag = (run < pre[k]) ? run : pre[k];    // bp.hpp:256-258
                const uint32_t sg = tsgn + (neg[k] ? 1u : 0u);        // bp.hpp:252-260
                c[k] = mag * ((sg & 1u) ? -alpha : alpha);            // bp.hpp:262
            }
            run = (a[k] < run) ? a[k] : run;
        }
    } else {
        // bp.hpp:205-218: c_k = (prod_{k'<k} t_k') * (prod_{k'>k} t_k'), products accumulated left-to-right
        // and right-to-left exactly as the two sweeps do; tanh is evaluated twice per edge in the reference on
        // the same argument, so once here.
        double t[DC];
        double pre = 1.0;
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            if (k < deg) {
                t[k] = ps_tanh_half(b[k]);
                c[k] = pre;
                pre *= t[k];
            }
        }
        double suf = 1.0;
        const double sigma = s ? -1.0 : 1.0;
#pragma unroll
        for (int k = DC - 1; k >= 0; --k) {
            if (k < deg) {
                const double x = c[k] * suf;
                c[k] = sigma * ps_atanh2(x);
                suf *= t[k];
            }
        }
    }
}

// c[0..deg): check->bit messages of the column in ascending row order.  Returns the posterior LLR
// (bp.hpp:278-288) and overwrites c[] with the new bit->check messages prefix_k + suffix_k (bp.hpp:280,312-317).
template <int DV>
__device__ __forceinline__ double bit_node_update(double (&c)[DV], int deg, double prior) {
    double pre[DV];
    double t = prior;
#pragma unroll
    for (int k = 0; k < DV; ++k) {
        if (k < deg) {
            pre[k] = t;
            t += c[k];
        }
    }
    double u = 0;
#pragma unroll
    for (int k = DV - 1; k >= 0; --k) {
        if (k < deg) {
            const double ck = c[k];
            c[k] = pre[k] + u;
            u += ck;
        }
    }
    return t;
}

}  // namespace bpb
