// ref_libm.h -- bit-exact restatement of the three libm functions the reference's product-sum update
// calls (std::tanh and std::log, reference src_cpp/bp.hpp:208,216,217,494,498), as glibc 2.39 (the libm the
// reference links against in this image, Ubuntu GLIBC 2.39-0ubuntu8.5) evaluates them on x86-64 CPUs with
// FMA + AVX2 (the ifunc variants __log_fma and __expm1_fma; tanh itself is not multiarch):
//
//   tanh(x)  : sysdeps/ieee754/dbl-64/s_tanh.c (fdlibm):  |x|<2^-55: x*(1+x);  |x|<1: t=expm1(-2|x|), -t/(t+2);
//              |x|<22: t=expm1(2|x|), 1-2/(t+2);  else +-1
//   expm1(x) : sysdeps/ieee754/dbl-64/s_expm1.c (fdlibm): k=round(x/ln2), r=x-k*ln2 in two pieces, a rational
//              approximation in hxs=r*r/2, then scaling by 2^k
//   log(x)   : sysdeps/ieee754/dbl-64/e_log.c (Arm optimized-routines): 128-entry table of (1/c, log c), a
//              degree-5 polynomial in r=z/c-1, and a separate degree-11 polynomial near 1
//
// glibc is a third-party dependency of the reference, absent from /root/reference; its algorithm is restated
// here.  These functions are not correctly rounded, and the FMA variants are the C sources compiled with
// -mfma, so WHICH multiply-adds are fused matters for the last bit.  The fusion pattern below follows the
// machine code of this image's libm.so.6 instruction by instruction, every operation is written with an
// explicit FMA / MUL / ADD / SUB / DIV (correctly rounded on both x86 and the GPU, never re-associated or
// contracted), and the constants were read out of that library's .rodata.  tests/test_ref_libm.py compiles
// this header for the host and checks it bit-for-bit against the live libm on ~10^7 arguments.
//
// The same header is compiled for the device (bp kernels) and for the host (the test).
#pragma once
#include <stdint.h>

#include <math.h>
#include <stdbool.h>
#include <string.h>

#if defined(__CUDACC__)
#define RL_FN __host__ __device__ __forceinline__
#else
#define RL_FN static inline
#endif

#if !defined(__CUDA_ARCH__)
static inline uint64_t rl_bits_(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}
static inline double rl_dbl_(uint64_t u) {
    double x;
    memcpy(&x, &u, 8);
    return x;
}
#endif

#if defined(__CUDA_ARCH__)
#define RL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RL_MUL(a, b) __dmul_rn((a), (b))
#define RL_ADD(a, b) __dadd_rn((a), (b))
#define RL_SUB(a, b) __dsub_rn((a), (b))
#define RL_DIV(a, b) __ddiv_rn((a), (b))
#define RL_D2I(a) __double2int_rz(a)
#define RL_BITS(x) ((uint64_t) __double_as_longlong(x))
#define RL_DBL(u) __longlong_as_double((long long) (u))
#define RL_LOG_TAB rl_log_tab_dev
#else
#define RL_FMA(a, b, c) fma((a), (b), (c))
#define RL_MUL(a, b) ((a) * (b))
#define RL_ADD(a, b) ((a) + (b))
#define RL_SUB(a, b) ((a) - (b))
#define RL_DIV(a, b) ((a) / (b))
#define RL_D2I(a) ((int) (a))
#define RL_BITS(x) rl_bits_(x)
#define RL_DBL(u) rl_dbl_(u)
#define RL_LOG_TAB rl_log_tab
#endif

// ---- log ----------------------------------------------------------------------------------------------
struct rl_log_entry {
    double invc, logc;
};
#define RL_LOG_TABLE_BODY \
    {0x1.734f0c3e0de9fp+0, -0x1.7cc7f79e69000p-2}, \
    {0x1.713786a2ce91fp+0, -0x1.76feec20d0000p-2}, \
    {0x1.6f26008fab5a0p+0, -0x1.713e31351e000p-2}, \
    {0x1.6d1a61f138c7dp+0, -0x1.6b85b38287800p-2}, \
    {0x1.6b1490bc5b4d1p+0, -0x1.65d5590807800p-2}, \
    {0x1.69147332f0cbap+0, -0x1.602d076180000p-2}, \
    {0x1.6719f18224223p+0, -0x1.5a8ca86909000p-2}, \
    {0x1.6524f99a51ed9p+0, -0x1.54f4356035000p-2}, \
    {0x1.63356aa8f24c4p+0, -0x1.4f637c36b4000p-2}, \
    {0x1.614b36b9ddc14p+0, -0x1.49da7fda85000p-2}, \
    {0x1.5f66452c65c4cp+0, -0x1.445923989a800p-2}, \
    {0x1.5d867b5912c4fp+0, -0x1.3edf439b0b800p-2}, \
    {0x1.5babccb5b90dep+0, -0x1.396ce448f7000p-2}, \
    {0x1.59d61f2d91a78p+0, -0x1.3401e17bda000p-2}, \
    {0x1.5805612465687p+0, -0x1.2e9e2ef468000p-2}, \
    {0x1.56397cee76bd3p+0, -0x1.2941b3830e000p-2}, \
    {0x1.54725e2a77f93p+0, -0x1.23ec58cda8800p-2}, \
    {0x1.52aff42064583p+0, -0x1.1e9e129279000p-2}, \
    {0x1.50f22dbb2bddfp+0, -0x1.1956d2b48f800p-2}, \
    {0x1.4f38f4734ded7p+0, -0x1.141679ab9f800p-2}, \
    {0x1.4d843cfde2840p+0, -0x1.0edd094ef9800p-2}, \
    {0x1.4bd3ec078a3c8p+0, -0x1.09aa518db1000p-2}, \
    {0x1.4a27fc3e0258ap+0, -0x1.047e65263b800p-2}, \
    {0x1.4880524d48434p+0, -0x1.feb224586f000p-3}, \
    {0x1.46dce1b192d0bp+0, -0x1.f474a7517b000p-3}, \
    {0x1.453d9d3391854p+0, -0x1.ea4443d103000p-3}, \
    {0x1.43a2744b4845ap+0, -0x1.e020d44e9b000p-3}, \
    {0x1.420b54115f8fbp+0, -0x1.d60a22977f000p-3}, \
    {0x1.40782da3ef4b1p+0, -0x1.cc00104959000p-3}, \
    {0x1.3ee8f5d57fe8fp+0, -0x1.c202956891000p-3}, \
    {0x1.3d5d9a00b4ce9p+0, -0x1.b81178d811000p-3}, \
    {0x1.3bd60c010c12bp+0, -0x1.ae2c9ccd3d000p-3}, \
    {0x1.3a5242b75dab8p+0, -0x1.a45402e129000p-3}, \
    {0x1.38d22cd9fd002p+0, -0x1.9a877681df000p-3}, \
    {0x1.3755bc5847a1cp+0, -0x1.90c6d69483000p-3}, \
    {0x1.35dce49ad36e2p+0, -0x1.87120a645c000p-3}, \
    {0x1.34679984dd440p+0, -0x1.7d68fb4143000p-3}, \
    {0x1.32f5cceffcb24p+0, -0x1.73cb83c627000p-3}, \
    {0x1.3187775a10d49p+0, -0x1.6a39a9b376000p-3}, \
    {0x1.301c8373e3990p+0, -0x1.60b3154b7a000p-3}, \
    {0x1.2eb4ebb95f841p+0, -0x1.5737d76243000p-3}, \
    {0x1.2d50a0219a9d1p+0, -0x1.4dc7b8fc23000p-3}, \
    {0x1.2bef9a8b7fd2ap+0, -0x1.4462c51d20000p-3}, \
    {0x1.2a91c7a0c1babp+0, -0x1.3b08abc830000p-3}, \
    {0x1.293726014b530p+0, -0x1.31b996b490000p-3}, \
    {0x1.27dfa5757a1f5p+0, -0x1.2875490a44000p-3}, \
    {0x1.268b39b1d3bbfp+0, -0x1.1f3b9f879a000p-3}, \
    {0x1.2539d838ff5bdp+0, -0x1.160c8252ca000p-3}, \
    {0x1.23eb7aac9083bp+0, -0x1.0ce7f57f72000p-3}, \
    {0x1.22a012ba940b6p+0, -0x1.03cdc49fea000p-3}, \
    {0x1.2157996cc4132p+0, -0x1.f57bdbc4b8000p-4}, \
    {0x1.201201dd2fc9bp+0, -0x1.e370896404000p-4}, \
    {0x1.1ecf4494d480bp+0, -0x1.d17983ef94000p-4}, \
    {0x1.1d8f5528f6569p+0, -0x1.bf9674ed8a000p-4}, \
    {0x1.1c52311577e7cp+0, -0x1.adc79202f6000p-4}, \
    {0x1.1b17c74cb26e9p+0, -0x1.9c0c3e7288000p-4}, \
    {0x1.19e010c2c1ab6p+0, -0x1.8a646b372c000p-4}, \
    {0x1.18ab07bb670bdp+0, -0x1.78d01b3ac0000p-4}, \
    {0x1.1778a25efbcb6p+0, -0x1.674f145380000p-4}, \
    {0x1.1648d354c31dap+0, -0x1.55e0e6d878000p-4}, \
    {0x1.151b990275fddp+0, -0x1.4485cdea1e000p-4}, \
    {0x1.13f0ea432d24cp+0, -0x1.333d94d6aa000p-4}, \
    {0x1.12c8b7210f9dap+0, -0x1.22079f8c56000p-4}, \
    {0x1.11a3028ecb531p+0, -0x1.10e4698622000p-4}, \
    {0x1.107fbda8434afp+0, -0x1.ffa6c6ad20000p-5}, \
    {0x1.0f5ee0f4e6bb3p+0, -0x1.dda8d4a774000p-5}, \
    {0x1.0e4065d2a9fcep+0, -0x1.bbcece4850000p-5}, \
    {0x1.0d244632ca521p+0, -0x1.9a1894012c000p-5}, \
    {0x1.0c0a77ce2981ap+0, -0x1.788583302c000p-5}, \
    {0x1.0af2f83c636d1p+0, -0x1.5715e67d68000p-5}, \
    {0x1.09ddb98a01339p+0, -0x1.35c8a49658000p-5}, \
    {0x1.08cabaf52e7dfp+0, -0x1.149e364154000p-5}, \
    {0x1.07b9f2f4e28fbp+0, -0x1.e72c082eb8000p-6}, \
    {0x1.06ab58c358f19p+0, -0x1.a55f152528000p-6}, \
    {0x1.059eea5ecf92cp+0, -0x1.63d62cf818000p-6}, \
    {0x1.04949cdd12c90p+0, -0x1.228fb8caa0000p-6}, \
    {0x1.038c6c6f0ada9p+0, -0x1.c317b20f90000p-7}, \
    {0x1.02865137932a9p+0, -0x1.419355daa0000p-7}, \
    {0x1.0182427ea7348p+0, -0x1.81203c2ec0000p-8}, \
    {0x1.008040614b195p+0, -0x1.0040979240000p-9}, \
    {0x1.fe01ff726fa1ap-1, 0x1.feff384900000p-9}, \
    {0x1.fa11cc261ea74p-1, 0x1.7dc41353d0000p-7}, \
    {0x1.f6310b081992ep-1, 0x1.3cea3c4c28000p-6}, \
    {0x1.f25f63ceeadcdp-1, 0x1.b9fc114890000p-6}, \
    {0x1.ee9c8039113e7p-1, 0x1.1b0d8ce110000p-5}, \
    {0x1.eae8078cbb1abp-1, 0x1.58a5bd001c000p-5}, \
    {0x1.e741aa29d0c9bp-1, 0x1.95c8340d88000p-5}, \
    {0x1.e3a91830a99b5p-1, 0x1.d276aef578000p-5}, \
    {0x1.e01e009609a56p-1, 0x1.07598e598c000p-4}, \
    {0x1.dca01e577bb98p-1, 0x1.253f5e30d2000p-4}, \
    {0x1.d92f20b7c9103p-1, 0x1.42edd8b380000p-4}, \
    {0x1.d5cac66fb5ccep-1, 0x1.606598757c000p-4}, \
    {0x1.d272caa5ede9dp-1, 0x1.7da76356a0000p-4}, \
    {0x1.cf26e3e6b2ccdp-1, 0x1.9ab434e1c6000p-4}, \
    {0x1.cbe6da2a77902p-1, 0x1.b78c7bb0d6000p-4}, \
    {0x1.c8b266d37086dp-1, 0x1.d431332e72000p-4}, \
    {0x1.c5894bd5d5804p-1, 0x1.f0a3171de6000p-4}, \
    {0x1.c26b533bb9f8cp-1, 0x1.067152b914000p-3}, \
    {0x1.bf583eeece73fp-1, 0x1.147858292b000p-3}, \
    {0x1.bc4fd75db96c1p-1, 0x1.2266ecdca3000p-3}, \
    {0x1.b951e0c864a28p-1, 0x1.303d7a6c55000p-3}, \
    {0x1.b65e2c5ef3e2cp-1, 0x1.3dfc33c331000p-3}, \
    {0x1.b374867c9888bp-1, 0x1.4ba366b7a8000p-3}, \
    {0x1.b094b211d304ap-1, 0x1.5933928d1f000p-3}, \
    {0x1.adbe885f2ef7ep-1, 0x1.66acd2418f000p-3}, \
    {0x1.aaf1d31603da2p-1, 0x1.740f8ec669000p-3}, \
    {0x1.a82e63fd358a7p-1, 0x1.815c0f51af000p-3}, \
    {0x1.a5740ef09738bp-1, 0x1.8e92954f68000p-3}, \
    {0x1.a2c2a90ab4b27p-1, 0x1.9bb3602f84000p-3}, \
    {0x1.a01a01393f2d1p-1, 0x1.a8bed1c2c0000p-3}, \
    {0x1.9d79f24db3c1bp-1, 0x1.b5b515c01d000p-3}, \
    {0x1.9ae2505c7b190p-1, 0x1.c2967ccbcc000p-3}, \
    {0x1.9852ef297ce2fp-1, 0x1.cf635d5486000p-3}, \
    {0x1.95cbaeea44b75p-1, 0x1.dc1bd3446c000p-3}, \
    {0x1.934c69de74838p-1, 0x1.e8c01b8cfe000p-3}, \
    {0x1.90d4f2f6752e6p-1, 0x1.f5509c0179000p-3}, \
    {0x1.8e6528effd79dp-1, 0x1.00e6c121fb800p-2}, \
    {0x1.8bfce9fcc007cp-1, 0x1.071b80e93d000p-2}, \
    {0x1.899c0dabec30ep-1, 0x1.0d46b9e867000p-2}, \
    {0x1.87427aa2317fbp-1, 0x1.13687334bd000p-2}, \
    {0x1.84f00acb39a08p-1, 0x1.1980d67234800p-2}, \
    {0x1.82a49e8653e55p-1, 0x1.1f8ffe0cc8000p-2}, \
    {0x1.8060195f40260p-1, 0x1.2595fd7636800p-2}, \
    {0x1.7e22563e0a329p-1, 0x1.2b9300914a800p-2}, \
    {0x1.7beb377dcb5adp-1, 0x1.3187210436000p-2}, \
    {0x1.79baa679725c2p-1, 0x1.377266dec1800p-2}, \
    {0x1.77907f2170657p-1, 0x1.3d54ffbaf3000p-2}, \
    {0x1.756cadbd6130cp-1, 0x1.432eee32fe000p-2}, \

static const struct rl_log_entry rl_log_tab[128] = {RL_LOG_TABLE_BODY};
#if defined(__CUDACC__)
__device__ static const struct rl_log_entry rl_log_tab_dev[128] = {RL_LOG_TABLE_BODY};
#endif

RL_FN double rl_log(double x) {
    const double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    const double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3,
                 A3 = 0x1.999b324f10111p-3, A4 = -0x1.55575e506c89fp-3;
    uint64_t ix = RL_BITS(x);
    if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {
        // 1 - 2^-4 <= x < 1 + 0x1.09p-4: degree-11 polynomial in r = x - 1 with a split leading term
        const double B0 = -0x1.0000000000000p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3,
                     B3 = 0x1.999999995dd0cp-3, B4 = -0x1.55555556745a7p-3, B5 = 0x1.24924a344de30p-3,
                     B6 = -0x1.fffffa4423d65p-4, B7 = 0x1.c7184282ad6cap-4, B8 = -0x1.999eb43b068ffp-4,
                     B9 = 0x1.78182f7afd085p-4, B10 = -0x1.5521375d145cdp-4;
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = RL_SUB(x, 1.0);
        double p1 = RL_FMA(r, B2, B1);
        double p2 = RL_FMA(r, B5, B4);
        const double r2 = RL_MUL(r, r);
        double p3 = RL_FMA(r, B8, B7);
        p1 = RL_FMA(r2, B3, p1);
        p2 = RL_FMA(r2, B6, p2);
        const double r3 = RL_MUL(r, r2);
        p3 = RL_FMA(r2, B9, p3);
        p3 = RL_FMA(r3, B10, p3);
        p3 = RL_FMA(p3, r3, p2);
        p3 = RL_FMA(p3, r3, p1);
        const double t = RL_FMA(r, 0x1p27, r);
        const double rhi = RL_FMA(-0x1p27, r, t);
        const double rhi2 = RL_MUL(rhi, rhi);
        const double rlo = RL_SUB(r, rhi);
        const double hi = RL_FMA(rhi2, B0, r);
        const double rmh = RL_SUB(r, hi);
        const double rsum = RL_ADD(r, rhi);
        double lo = RL_FMA(rhi2, B0, rmh);
        lo = RL_FMA(RL_MUL(B0, rlo), rsum, lo);
        const double y = RL_FMA(p3, r3, lo);
        return RL_ADD(hi, y);
    }
    const uint32_t top = (uint32_t) (ix >> 48);
    if (top - 0x0010u > 0x7fdfu) {
        // x < 2^-1022, or inf, or NaN
        if ((ix << 1) == 0) return RL_DBL(0xfff0000000000000ull);            // log(+-0) = -inf
        if (ix == 0x7ff0000000000000ull) return x;                            // log(inf) = inf
        if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) {
            // x < 0 -> NaN (invalid); NaN -> NaN
            return RL_DIV(RL_SUB(x, x), RL_SUB(x, x));
        }
        ix = RL_BITS(RL_MUL(x, 0x1p52));  // subnormal: normalise
        ix -= 52ull << 52;
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int) ((tmp >> 45) & 127);
    const int k = (int) ((int64_t) tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = RL_LOG_TAB[i].invc, logc = RL_LOG_TAB[i].logc;
    const double z = RL_DBL(iz);
    const double kd = (double) k;
    const double w = RL_FMA(kd, Ln2hi, logc);
    const double r = RL_FMA(z, invc, -1.0);
    const double q12 = RL_FMA(r, A2, A1);
    const double hi = RL_ADD(r, w);
    const double r2 = RL_MUL(r, r);
    double lo = RL_SUB(w, hi);
    lo = RL_ADD(lo, r);
    lo = RL_FMA(kd, Ln2lo, lo);
    const double r3 = RL_MUL(r, r2);
    double q = RL_FMA(r, A4, A3);
    lo = RL_FMA(r2, A0, lo);
    q = RL_FMA(q, r2, q12);
    const double y = RL_FMA(r3, q, lo);
    return RL_ADD(y, hi);
}

// ---- expm1 --------------------------------------------------------------------------------------------
RL_FN double rl_add_exponent(double y, int k) {  // high word += k << 20 (fdlibm SET_HIGH_WORD idiom)
    uint64_t u = RL_BITS(y);
    const uint32_t hi = (uint32_t) (u >> 32) + ((uint32_t) k << 20);
    return RL_DBL(((uint64_t) hi << 32) | (u & 0xffffffffull));
}

RL_FN double rl_expm1(double x0) {
    const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33, invln2 = 0x1.71547652b82fep+0;
    const double Q1 = -0x1.11111111110f4p-5, Q2 = 0x1.a01a019fe5585p-10, Q3 = -0x1.4ce199eaadbb7p-14,
                 Q4 = 0x1.0cfca86e65239p-18, Q5 = -0x1.afdb76e09c32dp-23;
    const uint64_t bits = RL_BITS(x0);
    const uint32_t hw = (uint32_t) (bits >> 32);
    const uint32_t hx = hw & 0x7fffffffu;
    const bool neg = (hw & 0x80000000u) != 0;
    if (hx > 0x40436879u) {          // |x| >= 56 ln2
        if (hx > 0x40862e41u) {      // |x| >= 709.78
            if (hx > 0x7fefffffu) {  // inf / NaN
                if (((hw & 0xfffffu) | (uint32_t) bits) != 0) return RL_ADD(x0, x0);  // NaN
                return neg ? -1.0 : x0;                                                // expm1(+-inf)
            }
            if (x0 > 0x1.62e42fefa39efp+9) return RL_MUL(1e300, 1e300);  // overflow -> inf
        }
        if (neg) return RL_SUB(1e-300, 1.0);  // x < -56 ln2: -1 (inexact)
    } else if (hx <= 0x3c8fffffu) {  // |x| < 2^-54
        return x0;
    }
    // Argument reduction x0 = k ln2 + x, without branches (lanes of a warp differ in k).  fdlibm has three cases:
    // |x| <= 0.5 ln2: k = 0, x = x0, c = 0;  0.5 ln2 < |x| < 1.5 ln2: k = +-1 with hi = x0 -+ ln2_hi, lo = +-ln2_lo;
    // else k = (int)(x0/ln2 +- 0.5), hi = x0 - k ln2_hi (one rounding: FMA in the __expm1_fma variant), lo = k ln2_lo.
    // The third formula reproduces the other two operation for operation: for k = +-1 the FMA is the same single
    // rounding of x0 -+ ln2_hi and k ln2_lo = +-ln2_lo exactly (the hx thresholds sit strictly inside
    // (0.5, 1.5) ln2, so the truncation gives +-1 there); for k = 0 it gives hi = x0, lo = +0, x = x0, c = +0.
    const bool small = hx <= 0x3fd62e42u;  // |x| <= 0.5 ln2 (by the high word, as fdlibm tests it)
    const double kf = RL_ADD(neg ? -0.5 : 0.5, RL_MUL(x0, invln2));
    const int k = small ? 0 : RL_D2I(kf);
    const double tk = (double) k;
    const double rhi = RL_FMA(-tk, ln2_hi, x0);
    const double rlo = RL_MUL(tk, ln2_lo);
    const double x = RL_SUB(rhi, rlo);
    const double c = RL_SUB(RL_SUB(rhi, x), rlo);
    // x is now in the primary range
    const double hfx = RL_MUL(x, 0.5);
    const double hxs = RL_MUL(x, hfx);
    const double R2 = RL_FMA(hxs, Q3, Q2);
    const double R3 = RL_FMA(hxs, Q5, Q4);
    const double h2 = RL_MUL(hxs, hxs);
    double R1 = RL_FMA(hxs, Q1, 1.0);
    const double h4 = RL_MUL(h2, h2);
    R1 = RL_FMA(h2, R2, R1);
    R1 = RL_FMA(h4, R3, R1);
    const double t = RL_FMA(-R1, hfx, 3.0);
    double e = RL_SUB(R1, t);
    const double den = RL_FMA(-x, t, 6.0);
    e = RL_DIV(e, den);
    e = RL_MUL(e, hxs);
    if (k == 0) return RL_SUB(x, RL_FMA(e, x, -hxs));  // x - (x*e - hxs)
    e = RL_SUB(e, c);
    e = RL_FMA(e, x, -c);
    e = RL_SUB(e, hxs);
    if (k == -1) return RL_FMA(0.5, RL_SUB(x, e), -0.5);
    if (k == 1) {
        if (x < -0.25) return RL_MUL(RL_SUB(e, RL_ADD(x, 0.5)), -2.0);
        return RL_FMA(RL_SUB(x, e), 2.0, 1.0);
    }
    if ((uint32_t) (k + 1) > 0x39u) {  // k <= -2 or k > 56: exp(x) - 1 with exp(x) = 2^k * (1 - (e - x))
        double y = RL_SUB(1.0, RL_SUB(e, x));
        y = rl_add_exponent(y, k);
        return RL_SUB(y, 1.0);
    }
    if (k > 19) {
        const double twomk = RL_DBL((uint64_t) (uint32_t) ((0x3ff - k) << 20) << 32);  // 2^-k
        double y = RL_SUB(x, RL_ADD(e, twomk));
        y = RL_ADD(y, 1.0);
        return rl_add_exponent(y, k);
    }
    const double onem = RL_DBL((uint64_t) (uint32_t) (0x3ff00000 - (0x200000 >> k)) << 32);  // 1 - 2^-k
    const double y = RL_SUB(onem, RL_SUB(e, x));
    return rl_add_exponent(y, k);
}

// ---- tanh ---------------------------------------------------------------------------------------------
RL_FN double rl_tanh(double x) {
    const uint64_t bits = RL_BITS(x);
    const uint32_t hw = (uint32_t) (bits >> 32);
    const uint32_t ix = hw & 0x7fffffffu;
    const bool neg = (hw & 0x80000000u) != 0;
    double z;
    if (ix >= 0x7ff00000u) {  // inf or NaN: 1/x +- 1
        return neg ? RL_SUB(RL_DIV(1.0, x), 1.0) : RL_ADD(RL_DIV(1.0, x), 1.0);
    }
    if (ix < 0x40360000u) {   // |x| < 22
        if ((ix | (uint32_t) bits) == 0) return x;                         // +-0
        if (ix < 0x3c800000u) return RL_MUL(x, RL_ADD(1.0, x));           // |x| < 2^-55
        const double ax = RL_DBL(bits & 0x7fffffffffffffffull);
        // |x| >= 1: t = expm1(2|x|), z = 1 - 2/(t+2);  |x| < 1: t = expm1(-2|x|), z = -t/(t+2).  One expm1 and one
        // division serve both cases (the lanes of a warp straddle |x| = 1 all the time).
        const bool big = ix >= 0x3ff00000u;
        const double t = rl_expm1(RL_MUL(big ? 2.0 : -2.0, ax));
        const double q = RL_DIV(big ? 2.0 : -t, RL_ADD(t, 2.0));
        z = big ? RL_SUB(1.0, q) : q;
    } else {
        z = 1.0;              // 1 - tiny rounds to 1
    }
    return neg ? -z : z;
}
