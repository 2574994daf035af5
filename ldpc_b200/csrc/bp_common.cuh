// bp_common.cuh -- shared device helpers for the B200 BP kernels (sm_100a).
//
// Arithmetic contract (what "bit-exact against the reference" needs):
//   * all message arithmetic is IEEE binary64, one rounding per reference operation: this
//     translation unit is compiled with -fmad=false so a*b+c is never contracted;
//   * comparisons are the reference's: `x <= 0` (zero counts as negative, NaN compares false,
//     reference src_cpp/bp.hpp:240,253,290,513,524) and strict `a < temp` for running minima
//     (bp.hpp:245,256,266,510);
//   * tanh is evaluated with the fdlibm formula glibc uses (sysdeps/ieee754/dbl-64/s_tanh.c:
//     1 - 2/(expm1(2|x|)+2) for 1<=|x|<22, -t/(t+2) with t=expm1(-2|x|) below 1, +-1 above 22) so
//     that the saturation pattern (where tanh rounds to exactly 1 and the check message becomes
//     +-inf, bp.hpp:216) is reproduced; only expm1/log differ from glibc, by <= 1 ulp.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace bpb {

constexpr int kProductSum = 0;  // reference bp.hpp:23-26
constexpr int kMinimumSum = 1;
constexpr int kSerial = 0;      // reference bp.hpp:28-32
constexpr int kParallel = 1;

__device__ __forceinline__ double ref_tanh(double x) {
    const double ax = fabs(x);
    double z;
    if (!(ax < 22.0)) {
        if (ax != ax) return x;  // NaN
        z = 1.0;                 // |x| >= 22 or inf
    } else if (ax >= 1.0) {
        const double t = expm1(2.0 * ax);
        z = 1.0 - 2.0 / (t + 2.0);
    } else if (ax < 2.77555756156289135e-17 /* 2^-55 */) {
        return x * (1.0 + x);
    } else {
        const double t = expm1(-2.0 * ax);
        z = -t / (t + 2.0);
    }
    return (x >= 0.0) ? z : -z;  // x = -0.0 handled by the 2^-55 branch
}

// alpha of the min-sum update, reference bp.hpp:222-228 / 459-465
__device__ __forceinline__ double ms_alpha(double ms_scaling_factor, int it) {
    if (ms_scaling_factor == 0.0) return 1.0 - ldexp(1.0, -it);  // 1 - 2^-it, exact like std::pow(2.0,-it)
    return ms_scaling_factor;
}

// streaming (evict-first) message traffic: every message is touched once per half-iteration and the
// working set (GBs) never fits L2, so keep it from displacing the graph tables in L1/L2.
__device__ __forceinline__ double ld_msg(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_msg(double *p, double v) { __stcs(p, v); }

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t r;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
    return r;
}

}  // namespace bpb
