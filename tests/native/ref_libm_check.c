/* Host check of ldpc_b200/csrc/ref_libm.h against the live libm: bit-for-bit on random and edge arguments.
 * Build: gcc -O2 -ffp-contract=off -mfma -o ref_libm_check ref_libm_check.c -lm ; prints mismatch counts. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../ldpc_b200/csrc/ref_libm.h"

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
}
static double urand(double lo, double hi) { return lo + (hi - lo) * ((rnd() >> 11) * 0x1p-53); }
static int same(double a, double b) { return rl_bits_(a) == rl_bits_(b) || (a != a && b != b); }

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 2000000;
    long bad_log = 0, bad_expm1 = 0, bad_tanh = 0, bad_chain = 0;
    double edge[] = {0.0, -0.0, 1.0, -1.0, 0.5, 2.0, INFINITY, -INFINITY, NAN, 1e-320, 4.9e-324, 1e-308, 2.2250738585072014e-308,
                     0.9375, 1.0647, 1.06469, 0.93749999, 1e300, 1.7976931348623157e308, 22.0, 21.999999, 19.06, 0x1p-55, 0x1p-54,
                     0.34657359027997264, 1.0397207708399179, 38.816242111356935, 709.782712893384, 710.0, -40.0, -38.8, 56.0, 39.5, 13.2, 13.9};
    for (unsigned i = 0; i < sizeof(edge) / sizeof(edge[0]); i++) {
        for (int sg = 0; sg < 2; sg++) {
            double x = sg ? -edge[i] : edge[i];
            if (!same(rl_log(x), log(x))) { bad_log++; printf("log edge %a: %a vs %a\n", x, rl_log(x), log(x)); }
            if (!same(rl_expm1(x), expm1(x))) { bad_expm1++; printf("expm1 edge %a: %a vs %a\n", x, rl_expm1(x), expm1(x)); }
            if (!same(rl_tanh(x), tanh(x))) { bad_tanh++; printf("tanh edge %a: %a vs %a\n", x, rl_tanh(x), tanh(x)); }
        }
    }
    for (long i = 0; i < n; i++) {
        double xs[8];
        xs[0] = urand(0.0, 4.0);                       /* log near 1 and small */
        xs[1] = exp(urand(-700.0, 700.0));             /* log over the whole range */
        xs[2] = rl_dbl_(rnd() & 0x7fffffffffffffffull); /* any positive bit pattern (incl. subnormal, nan) */
        xs[3] = urand(0.9, 1.1);
        xs[4] = urand(-45.0, 45.0);                    /* expm1 / tanh arguments */
        xs[5] = urand(-2.2, 2.2);
        xs[6] = urand(-1e-3, 1e-3);
        xs[7] = rl_dbl_(rnd());                        /* any bit pattern */
        for (int j = 0; j < 8; j++) {
            double x = xs[j];
            if (!same(rl_log(x), log(x))) { if (bad_log++ < 5) printf("log %a: %a vs %a\n", x, rl_log(x), log(x)); }
            if (!same(rl_expm1(x), expm1(x))) { if (bad_expm1++ < 5) printf("expm1 %a: %a vs %a\n", x, rl_expm1(x), expm1(x)); }
            if (!same(rl_tanh(x), tanh(x))) { if (bad_tanh++ < 5) printf("tanh %a: %a vs %a\n", x, rl_tanh(x), tanh(x)); }
        }
        /* the product-sum chain itself: log((1+c)/(1-c)) with c a product of tanh's (bp.hpp:208-217) */
        double b1 = urand(-40, 40), b2 = urand(-6, 6), b3 = urand(-60, 60);
        double c1 = tanh(b1 / 2) * tanh(b2 / 2) * tanh(b3 / 2);
        double c2 = rl_tanh(b1 / 2) * rl_tanh(b2 / 2) * rl_tanh(b3 / 2);
        if (!same(log((1 + c1) / (1 - c1)), rl_log((1 + c2) / (1 - c2)))) bad_chain++;
    }
    printf("checked %ld x 8 arguments: log %ld expm1 %ld tanh %ld chain %ld mismatches\n", n, bad_log, bad_expm1, bad_tanh, bad_chain);
    return (bad_log || bad_expm1 || bad_tanh || bad_chain) ? 1 : 0;
}
