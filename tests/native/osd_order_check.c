/* Host check of ldpc_b200/csrc/osd_order.h against the live libc qsort with the reference's record layout and
 * comparator (reference src_cpp/sort.hpp:27-62): random keys with many ties, +-0, +-inf and NaNs, n = 1 .. 4000.
 * Build: gcc -O2 -o osd_order_check osd_order_check.c ; prints the number of orders that differ. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../ldpc_b200/csrc/osd_order.h"

struct rec { double value; int index; };
static int cmp(const void *a, const void *b) {
    const struct rec *x = a, *y = b;
    if (x->value > y->value) return 1;
    if (x->value < y->value) return -1;
    return 0;
}
static unsigned long long s = 88172645463325252ull;
static unsigned rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (unsigned) (s >> 20); }

int main(int argc, char **argv) {
    int trials = argc > 1 ? atoi(argv[1]) : 3000, bad = 0, with_nan = 0;
    for (int t = 0; t < trials; t++) {
        int n = 1 + (int) (rnd() % (t % 7 == 0 ? 4000 : 300));
        double *v = malloc(sizeof(double) * n);
        struct rec *r = malloc(sizeof(struct rec) * n);
        uint16_t *a = malloc(2 * n), *b = malloc(2 * n);
        int mode = t % 4;
        for (int i = 0; i < n; i++) {
            double x = (double) (rnd() % 9) - 4.0;               /* many ties */
            if (mode == 1) x = (double) rnd() / 65536.0 - 8000.0; /* few ties */
            if (x == 0 && (rnd() & 1)) x = -0.0;
            if (mode >= 2 && rnd() % 11 == 0) x = (rnd() & 1) ? INFINITY : -INFINITY;
            if (mode == 3 && rnd() % 13 == 0) { x = NAN; if (rnd() & 1) x = -x; }
            v[i] = x; r[i].value = x; r[i].index = i; a[i] = (uint16_t) i;
        }
        qsort(r, n, sizeof(r[0]), cmp);
        uint16_t *src = a, *dst = b;
        for (int d = osd_order_depth(n) - 1; d >= 0; --d) {
            for (int k = 0; k < (1 << d); k++) {
                int lo, len;
                osd_order_node(n, d, k, &lo, &len);
                osd_order_merge(src, dst, v, lo, len);
            }
            uint16_t *tmp = src; src = dst; dst = tmp;
        }
        int diff = 0, has_nan = 0;
        for (int i = 0; i < n; i++) { diff |= (src[i] != (uint16_t) r[i].index); has_nan |= (v[i] != v[i]); }
        bad += diff; with_nan += has_nan;
        free(v); free(r); free(a); free(b);
    }
    printf("%d orders differ (%d trials, %d with NaN keys)\n", bad, trials, with_nan);
    return bad != 0;
}
