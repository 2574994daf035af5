// CPU check of ldpc_b200/csrc/stl_sort.h against the real std::sort of this toolchain's libstdc++ with the
// reference's comparator (src_cpp/bp.hpp:471-481): random keys, tie-heavy keys (few distinct values, all equal),
// NaNs, already sorted / reversed / organ-pipe inputs, and a median-of-3 killer that drives introsort into its
// heapsort fallback.  Usage: stl_sort_check <rounds>; prints "<k> permutations differ".
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../../ldpc_b200/csrc/stl_sort.h"

static long check(const std::vector<double> &key, const std::vector<int> &start) {
    std::vector<int> want = start, got = start;
    std::sort(want.begin(), want.end(), [&](int a, int b) { return key[a] > key[b]; });
    std::vector<int> stack(stlsort::kStackInts);
    stlsort::Sorter<int *, const double *> s{got.data(), key.data()};
    s.sort((int) got.size(), stack.data());
    return want == got ? 0 : 1;
}

int main(int argc, char **argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
    std::mt19937_64 rng(12345);
    long bad = 0, heap_cases = 0;
    for (int r = 0; r < rounds; r++) {
        const int n = 1 + (int) (rng() % (r % 7 == 0 ? 3000 : 400));
        std::vector<double> key((size_t) n);
        const int mode = r % 8;
        const int distinct = 1 + (int) (rng() % 6);
        for (int i = 0; i < n; i++) {
            switch (mode) {
                case 0: key[i] = (double) (rng() % 1000000) / 997.0 - 300.0; break;             // mostly distinct
                case 1: key[i] = (double) (rng() % (unsigned) distinct) * 0.625; break;          // heavy ties
                case 2: key[i] = 2.9444389791664403; break;                                      // all equal
                case 3: key[i] = (rng() % 9 == 0) ? NAN : (double) (rng() % 5); break;           // NaNs + ties
                case 4: key[i] = (double) i; break;                                              // ascending
                case 5: key[i] = (double) -i; break;                                             // descending
                case 6: key[i] = (double) (i < n / 2 ? i : n - i); break;                        // organ pipe
                default: key[i] = (rng() % 3 == 0) ? INFINITY : ((rng() % 3 == 0) ? -INFINITY : (double) (rng() % 4));
            }
        }
        std::vector<int> start((size_t) n);
        std::iota(start.begin(), start.end(), 0);
        if (r % 3 == 1) std::shuffle(start.begin(), start.end(), rng);
        if (r % 3 == 2) std::reverse(start.begin(), start.end());
        bad += check(key, start);
        // the schedule is re-sorted in place every iteration: sort again from the previous result with new keys
        std::vector<int> again = start;
        std::sort(again.begin(), again.end(), [&](int a, int b) { return key[a] > key[b]; });
        for (int i = 0; i < n; i++) key[i] = (double) (rng() % (unsigned) (distinct + 1));
        bad += check(key, again);
    }
    // median-of-3 killer (Musser): forces quadratic partitioning, so the depth limit trips and heapsort runs.
    // The comparator is "greater", so negate the classic ascending killer.
    for (int n : {64, 200, 1000, 4096}) {
        std::vector<double> key((size_t) n);
        const int k = n / 2;
        for (int i = 1; i <= k; i++) {
            if (i % 2 == 1) { key[(size_t) i - 1] = -(double) i; key[(size_t) i] = -(double) (k + i); }
            key[(size_t) k + i - 1] = -(double) (2 * i);
        }
        std::vector<int> start((size_t) n);
        std::iota(start.begin(), start.end(), 0);
        bad += check(key, start);
        heap_cases++;
    }
    printf("%ld permutations differ (%d random rounds x 2, %ld killer inputs)\n", bad, rounds, heap_cases);
    return bad ? 1 : 0;
}
