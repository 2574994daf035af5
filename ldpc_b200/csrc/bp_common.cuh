// bp_common.cuh -- shared device helpers for the B200 BP kernels (sm_100a).
//
// Arithmetic contract (what "bit-exact against the reference" needs):
//   * all message arithmetic is IEEE binary64, one rounding per reference operation: this
//     translation unit is compiled with -fmad=false so a*b+c is never contracted;
//   * comparisons are the reference's: `x <= 0` (zero counts as negative, NaN compares false,
//     reference src_cpp/bp.hpp:240,253,290,513,524) and strict `a < temp` for running minima
//     (bp.hpp:245,256,266,510);
//   * std::tanh and std::log of the product-sum update (bp.hpp:208,216,494,498) are evaluated by
//     rl_tanh / rl_log (ref_libm.h): the glibc 2.39 algorithms restated operation by operation, so the
//     product-sum messages -- including where tanh rounds to exactly 1 and the check message becomes
//     +-inf (bp.hpp:216) -- are the doubles the reference computes on an x86-64 host with FMA.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "ref_libm.h"

namespace bpb {

constexpr int kProductSum = 0;  // reference bp.hpp:23-26
constexpr int kMinimumSum = 1;
constexpr int kSerial = 0;      // reference bp.hpp:28-32
constexpr int kParallel = 1;

// Out-of-line on purpose: a product-sum row update makes 2*d_c tanh and d_c log evaluations; inlining every
// copy bloats the unrolled kernels past the instruction cache (and ptxas takes minutes per kernel).
__device__ __noinline__ static double ps_tanh_half(double b) { return rl_tanh(b / 2); }   // tanh(b2c / 2), bp.hpp:208
__device__ __noinline__ static double ps_atanh2(double x) { return rl_log((1 + x) / (1 - x)); }  // bp.hpp:216

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
