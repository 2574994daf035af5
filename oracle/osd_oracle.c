/*
 * osd_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C restatement of the reference's OSD-0 post-processor
 * (ldpc::osd::OsdDecoder::decode with osd_order == 0, src_cpp/osd.hpp:110-117):
 *   1. order the columns by ascending posterior LLR with libc qsort on
 *      {double value; int index;} records (src_cpp/sort.hpp:27-62) -- ties are
 *      whatever libc's qsort does with that comparator, so the same call is made;
 *   2. eliminate columns in that order until the syndrome lies in the span of
 *      the pivots found so far, then solve on the pivot columns, all other bits 0
 *      (RowReduce::fast_solve + lu_solve, src_cpp/gf2sparse_linalg.hpp:237-401).
 * The reference performs a sparse LU with row swaps chosen by row weight; the
 * solution does not depend on those choices: the pivot columns are the greedy
 * independent set of the ordering and the solution supported on independent
 * columns is unique.  So this restatement uses a dense byte matrix and
 * Gauss-Jordan.  Pinned against oracle/_ref (tests/test_oracle_vs_ref.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct bpo_str { /* sort.hpp:27-30 */
    double value;
    int index;
};

static int bpo_cmp(const void *a, const void *b) { /* sort.hpp:36-46 */
    const struct bpo_str *a1 = (const struct bpo_str *) a;
    const struct bpo_str *a2 = (const struct bpo_str *) b;
    if (a1->value > a2->value) return 1;
    if (a1->value < a2->value) return -1;
    return 0;
}

/* sort.hpp:48-62 */
void bpo_soft_decision_col_sort(const double *llr, int32_t *cols, int n) {
    struct bpo_str *objects = (struct bpo_str *) malloc(sizeof(struct bpo_str) * (size_t) (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        objects[i].value = llr[i];
        objects[i].index = i;
    }
    qsort(objects, (size_t) n, sizeof(objects[0]), bpo_cmp);
    for (int i = 0; i < n; i++) cols[i] = objects[i].index;
    free(objects);
}

/* OSD-0 for a batch: syndromes [B][m], llr [B][n] -> out [B][n]. */
int bpo_osd0_batch(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, const uint8_t *syndromes,
                   const double *llr, int64_t batch, uint8_t *out) {
    uint8_t *H = (uint8_t *) calloc((size_t) m * (size_t) n + 1, 1);
    uint8_t *A = (uint8_t *) malloc((size_t) m * (size_t) n + 1);
    uint8_t *y = (uint8_t *) malloc((size_t) m + 1);
    int32_t *order = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 1));
    int32_t *pivcol = (int32_t *) malloc(sizeof(int32_t) * (size_t) (m + 1));
    if (!H || !A || !y || !order || !pivcol) return -1;
    for (int64_t k = 0; k < nnz; k++) H[(size_t) rows[k] * n + cols[k]] = 1;
    int max_rank = m < n ? m : n;
    for (int64_t b = 0; b < batch; b++) {
        bpo_soft_decision_col_sort(llr + b * (int64_t) n, order, n);
        memcpy(A, H, (size_t) m * (size_t) n);
        for (int i = 0; i < m; i++) y[i] = syndromes[b * (int64_t) m + i] ? 1 : 0;
        int rank = 0;
        for (int jj = 0; jj < n && rank < max_rank; jj++) {
            int c = order[jj];
            int r = -1;
            for (int i = rank; i < m; i++)
                if (A[(size_t) i * n + c]) {
                    r = i;
                    break;
                }
            if (r < 0) continue;
            if (r != rank) {
                for (int t = 0; t < n; t++) {
                    uint8_t tmp = A[(size_t) r * n + t];
                    A[(size_t) r * n + t] = A[(size_t) rank * n + t];
                    A[(size_t) rank * n + t] = tmp;
                }
                uint8_t ty = y[r];
                y[r] = y[rank];
                y[rank] = ty;
            }
            for (int i = 0; i < m; i++) {
                if (i != rank && A[(size_t) i * n + c]) {
                    for (int t = 0; t < n; t++) A[(size_t) i * n + t] ^= A[(size_t) rank * n + t];
                    y[i] ^= y[rank];
                }
            }
            pivcol[rank] = c;
            rank++;
            int in_image = 1; /* gf2sparse_linalg.hpp:373-383 */
            for (int i = rank; i < m; i++)
                if (y[i]) {
                    in_image = 0;
                    break;
                }
            if (in_image) break;
        }
        uint8_t *x = out + b * (int64_t) n;
        memset(x, 0, (size_t) n);
        for (int r = 0; r < rank; r++) x[pivcol[r]] = y[r];
    }
    free(H);
    free(A);
    free(y);
    free(order);
    free(pivcol);
    return 0;
}
