/*
 * bp_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, single-threaded, flat-array restatement of the belief-propagation
 * syndrome decoder of quantumgizmos/ldpc (reference paths are relative to
 * /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object, and
 * only as the checker.  The CUDA product (ldpc_b200/csrc) never links or calls it.
 *
 * Parity pinning: this restatement is checked bit-for-bit (decoding, converge,
 * iterations and every posterior-LLR bit pattern) against
 *   (1) the reference's own C++ compiled in place into oracle/_ref/libref_bp.so
 *       (oracle/ref_wrap.cpp, tests/test_oracle_cpu.py::test_port_bit_identical_to_reference), and
 *   (2) the reference's known-answer tests (cpp_test/TestBPDecoder.cpp:122-344,
 *       python_test/test_bp_decoder.py:175-235) restated in tests/kat.py (tests/test_oracle_cpu.py),
 *   (3) the committed fixtures tests/golden/*.npz generated from (1).
 *
 * What is restated (all arithmetic IEEE-754 binary64, evaluation order preserved):
 *   - initialise_log_domain_bp            src_cpp/bp.hpp:147-157
 *   - decode dispatch / received vector   src_cpp/bp.hpp:159-190
 *   - bp_decode_parallel  (PS + MS)       src_cpp/bp.hpp:192-325
 *   - bp_decode_serial    (PS + MS)       src_cpp/bp.hpp:451-545, incl. the SERIAL_RELATIVE re-sort (:469-482) with
 *     libstdc++'s std::sort restated
 *   - GF2Sparse::mulvec                   src_cpp/gf2sparse.hpp:177-214
 *   - traversal order: rows by ascending column, columns by ascending row, because
 *     insert_entry keeps both lists sorted (src_cpp/sparse_matrix_base.hpp:423-482).
 *
 * The reference stores both messages inside linked-list nodes; here they are two
 * flat arrays indexed by the CSR edge id (row-major, ascending column), with the
 * column traversal expressed as a permutation csc2csr[] into that numbering.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BPO_PRODUCT_SUM 0 /* bp.hpp:23-26 */
#define BPO_MINIMUM_SUM 1
#define BPO_SERIAL 0 /* bp.hpp:28-32 */
#define BPO_PARALLEL 1
#define BPO_SERIAL_RELATIVE 2

typedef struct {
    int m, n, nnz;
    int *row_ptr; /* m+1 */
    int *col_idx; /* nnz, ascending inside each row */
    int *col_ptr; /* n+1 */
    int *row_idx; /* nnz, ascending inside each column (CSC order) */
    int *csc2csr; /* nnz: CSC position -> CSR edge id */
} bpo_graph;

static void bpo_graph_free(bpo_graph *g) {
    free(g->row_ptr);
    free(g->col_idx);
    free(g->col_ptr);
    free(g->row_idx);
    free(g->csc2csr);
}

/* Build sorted CSR + CSC from COO in any order.  Mirrors what repeated
 * insert_entry calls produce (sparse_matrix_base.hpp:423-482): each row list sorted
 * by column, each column list sorted by row, duplicates collapse to one entry. */
static int bpo_graph_build(bpo_graph *g, int m, int n, int64_t nnz_in, const int32_t *rows, const int32_t *cols) {
    memset(g, 0, sizeof(*g));
    g->m = m;
    g->n = n;
    /* dense bitmap free approach: counting sort by (row, col) */
    int64_t *key = (int64_t *) malloc(sizeof(int64_t) * (size_t) (nnz_in > 0 ? nnz_in : 1));
    if (!key) return -1;
    for (int64_t k = 0; k < nnz_in; k++) {
        if (rows[k] < 0 || rows[k] >= m || cols[k] < 0 || cols[k] >= n) {
            free(key);
            return -2;
        }
        key[k] = (int64_t) rows[k] * n + cols[k];
    }
    /* simple qsort on int64 keys */
    int cmp64(const void *a, const void *b);
    qsort(key, (size_t) nnz_in, sizeof(int64_t), cmp64);
    int64_t u = 0;
    for (int64_t k = 0; k < nnz_in; k++)
        if (k == 0 || key[k] != key[k - 1]) key[u++] = key[k];
    int nnz = (int) u;
    g->nnz = nnz;
    g->row_ptr = (int *) calloc((size_t) m + 1, sizeof(int));
    g->col_idx = (int *) malloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    g->col_ptr = (int *) calloc((size_t) n + 1, sizeof(int));
    g->row_idx = (int *) malloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    g->csc2csr = (int *) malloc(sizeof(int) * (size_t) (nnz > 0 ? nnz : 1));
    for (int e = 0; e < nnz; e++) {
        int r = (int) (key[e] / n), c = (int) (key[e] % n);
        g->row_ptr[r + 1]++;
        g->col_ptr[c + 1]++;
        g->col_idx[e] = c;
    }
    for (int i = 0; i < m; i++) g->row_ptr[i + 1] += g->row_ptr[i];
    for (int j = 0; j < n; j++) g->col_ptr[j + 1] += g->col_ptr[j];
    int *fill = (int *) calloc((size_t) n + 1, sizeof(int));
    for (int e = 0; e < nnz; e++) { /* CSR order is ascending row, so CSC gets ascending rows */
        int r = (int) (key[e] / n), c = (int) (key[e] % n);
        int p = g->col_ptr[c] + fill[c]++;
        g->row_idx[p] = r;
        g->csc2csr[p] = e;
    }
    free(fill);
    free(key);
    return 0;
}

int cmp64(const void *a, const void *b) {
    int64_t x = *(const int64_t *) a, y = *(const int64_t *) b;
    return (x > y) - (x < y);
}

/* alpha for min-sum, bp.hpp:222-228 and :459-465 */
static double bpo_alpha(double ms_scaling_factor, int it) {
    if (ms_scaling_factor == 0.0) return 1.0 - pow(2.0, -1.0 * it);
    return ms_scaling_factor;
}

typedef struct {
    const bpo_graph *g;
    const double *channel; /* n */
    int max_iter;
    int method, schedule;
    double ms_scaling_factor;
    const int32_t *serial_order; /* n entries */
    int serial_order_len;
    int *order_work; /* SERIAL_RELATIVE: the evolving serial_schedule_order member */
    /* work */
    double *b2c, *c2b; /* nnz each, CSR edge numbering */
    double *prior;     /* n */
    double *llr;       /* n */
    uint8_t *decoding; /* n */
    uint8_t *cand;     /* m */
    int iterations;
    int converge;
} bpo_state;

/* bp.hpp:147-157 */
static void bpo_init(bpo_state *s) {
    const bpo_graph *g = s->g;
    for (int j = 0; j < g->n; j++) {
        s->prior[j] = log((1 - s->channel[j]) / s->channel[j]);
        for (int p = g->col_ptr[j]; p < g->col_ptr[j + 1]; p++) s->b2c[g->csc2csr[p]] = s->prior[j];
    }
}

/* bp.hpp:192-325 */
static void bpo_parallel(bpo_state *s, const uint8_t *syndrome) {
    const bpo_graph *g = s->g;
    s->converge = 0;
    bpo_init(s);
    for (int it = 1; it <= s->max_iter; it++) {
        if (s->method == BPO_PRODUCT_SUM) {
            /* bp.hpp:201-219 */
            for (int i = 0; i < g->m; i++) {
                s->cand[i] = 0;
                double temp = 1.0;
                for (int e = g->row_ptr[i]; e < g->row_ptr[i + 1]; e++) {
                    s->c2b[e] = temp;
                    temp *= tanh(s->b2c[e] / 2);
                }
                temp = 1;
                for (int e = g->row_ptr[i + 1] - 1; e >= g->row_ptr[i]; e--) {
                    s->c2b[e] *= temp;
                    int message_sign = syndrome[i] != 0u ? -1 : 1;
                    s->c2b[e] = message_sign * log((1 + s->c2b[e]) / (1 - s->c2b[e]));
                    temp *= tanh(s->b2c[e] / 2);
                }
            }
        } else {
            /* bp.hpp:220-273 */
            double alpha = bpo_alpha(s->ms_scaling_factor, it);
            for (int i = 0; i < g->m; i++) {
                s->cand[i] = 0;
                int total_sgn = syndrome[i];
                double temp = DBL_MAX;
                for (int e = g->row_ptr[i]; e < g->row_ptr[i + 1]; e++) {
                    if (s->b2c[e] <= 0) total_sgn += 1;
                    s->c2b[e] = temp;
                    double a = fabs(s->b2c[e]);
                    if (a < temp) temp = a;
                }
                temp = DBL_MAX;
                for (int e = g->row_ptr[i + 1] - 1; e >= g->row_ptr[i]; e--) {
                    int sgn = total_sgn;
                    if (s->b2c[e] <= 0) sgn += 1;
                    if (temp < s->c2b[e]) s->c2b[e] = temp;
                    int message_sign = (sgn % 2 == 0) ? 1 : -1;
                    s->c2b[e] *= message_sign * alpha;
                    double a = fabs(s->b2c[e]);
                    if (a < temp) temp = a;
                }
            }
        }
        /* bp.hpp:276-298: posterior, hard decision, candidate syndrome */
        for (int j = 0; j < g->n; j++) {
            double temp = s->prior[j];
            for (int p = g->col_ptr[j]; p < g->col_ptr[j + 1]; p++) {
                int e = g->csc2csr[p];
                s->b2c[e] = temp;
                temp += s->c2b[e];
            }
            s->llr[j] = temp;
            if (temp <= 0) {
                s->decoding[j] = 1;
                for (int p = g->col_ptr[j]; p < g->col_ptr[j + 1]; p++) s->cand[g->row_idx[p]] ^= 1;
            } else {
                s->decoding[j] = 0;
            }
        }
        /* bp.hpp:300-308 */
        if (memcmp(s->cand, syndrome, (size_t) g->m) == 0) s->converge = 1;
        s->iterations = it;
        if (s->converge) return;
        /* bp.hpp:311-318 */
        for (int j = 0; j < g->n; j++) {
            double temp = 0;
            for (int p = g->col_ptr[j + 1] - 1; p >= g->col_ptr[j]; p--) {
                int e = g->csc2csr[p];
                s->b2c[e] += temp;
                temp += s->c2b[e];
            }
        }
    }
}

/*
 * std::sort(order.begin(), order.end(), comp) with comp(a, b) = key[a] > key[b]  (bp.hpp:469-482), as libstdc++
 * (GCC 13, bits/stl_algo.h, bits/stl_heap.h -- a third-party dependency absent from /root/reference) implements it:
 * introsort (median of three to the front, unguarded Hoare partition, recursion on the right part, heapsort after
 * 2*floor(log2 n) levels) followed by the final insertion sort with threshold 16.  The permutation of tied keys is a
 * property of this algorithm, so it is restated here function by function (recursive, like the original).
 */
static const double *bso_key;
static int bso_comp(int a, int b) { return bso_key[a] > bso_key[b]; }
static void bso_swap(int *x, int *y) { int t = *x; *x = *y; *y = t; }
static void bso_move_median_to_first(int *result, int *a, int *b, int *c) {
    if (bso_comp(*a, *b)) {
        if (bso_comp(*b, *c)) bso_swap(result, b);
        else if (bso_comp(*a, *c)) bso_swap(result, c);
        else bso_swap(result, a);
    } else if (bso_comp(*a, *c)) bso_swap(result, a);
    else if (bso_comp(*b, *c)) bso_swap(result, c);
    else bso_swap(result, b);
}
static int *bso_unguarded_partition(int *first, int *last, int *pivot) {
    for (;;) {
        while (bso_comp(*first, *pivot)) ++first;
        --last;
        while (bso_comp(*pivot, *last)) --last;
        if (!(first < last)) return first;
        bso_swap(first, last);
        ++first;
    }
}
static void bso_push_heap(int *first, long hole, long top, int value) {
    long parent = (hole - 1) / 2;
    while (hole > top && bso_comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void bso_adjust_heap(int *first, long hole, long len, int value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (bso_comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    bso_push_heap(first, hole, top, value);
}
static void bso_heap_sort(int *first, int *last) {
    long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            bso_adjust_heap(first, parent, len, first[parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        int value = *last;
        *last = *first;
        bso_adjust_heap(first, 0, last - first, value);
    }
}
static void bso_introsort_loop(int *first, int *last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) {
            bso_heap_sort(first, last);
            return;
        }
        --depth_limit;
        int *mid = first + (last - first) / 2;
        bso_move_median_to_first(first, first + 1, mid, last - 1);
        int *cut = bso_unguarded_partition(first + 1, last, first);
        bso_introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}
static void bso_unguarded_linear_insert(int *last) {
    int val = *last;
    int *next = last - 1;
    while (bso_comp(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}
static void bso_insertion_sort(int *first, int *last) {
    if (first == last) return;
    for (int *i = first + 1; i != last; ++i) {
        if (bso_comp(*i, *first)) {
            int val = *i;
            memmove(first + 1, first, sizeof(int) * (size_t) (i - first));
            *first = val;
        } else {
            bso_unguarded_linear_insert(i);
        }
    }
}
static void bso_sort_desc(int *order, int n, const double *key) {
    if (n <= 0) return;
    bso_key = key;
    long lg = 0;
    while (((long) n >> (lg + 1)) != 0) lg++;
    bso_introsort_loop(order, order + n, 2 * lg);
    if (n > 16) {
        bso_insertion_sort(order, order + 16);
        for (int *i = order + 16; i != order + n; ++i) bso_unguarded_linear_insert(i);
    } else {
        bso_insertion_sort(order, order + n);
    }
}

/* bp.hpp:451-545: SERIAL, and SERIAL_RELATIVE (schedule 2, :469-482: the order is re-sorted by descending LLR -- by
 * the priors in the first iteration -- before every sweep; `order_work` is the decoder's serial_schedule_order member,
 * which the sort permutes in place).  The random serial schedule is out of scope. */
static void bpo_serial(bpo_state *s, const uint8_t *syndrome) {
    const bpo_graph *g = s->g;
    s->converge = 0;
    bpo_init(s);
    for (int it = 1; it <= s->max_iter; it++) {
        double alpha = bpo_alpha(s->ms_scaling_factor, it);
        if (s->schedule == BPO_SERIAL_RELATIVE) {
            /* bp.hpp:469-482; `prior` holds log((1-p)/p), the key of the first iteration */
            bso_sort_desc(s->order_work, s->serial_order_len, it != 1 ? s->llr : s->prior);
        }
        for (int oi = 0; oi < s->serial_order_len; oi++) {
            int j = s->schedule == BPO_SERIAL_RELATIVE ? s->order_work[oi] : s->serial_order[oi];
            s->llr[j] = log((1 - s->channel[j]) / s->channel[j]);
            if (s->method == BPO_PRODUCT_SUM) {
                /* bp.hpp:488-501 */
                for (int p = g->col_ptr[j]; p < g->col_ptr[j + 1]; p++) {
                    int e = g->csc2csr[p];
                    int i = g->row_idx[p];
                    double c = 1.0;
                    for (int f = g->row_ptr[i]; f < g->row_ptr[i + 1]; f++)
                        if (f != e) c *= tanh(s->b2c[f] / 2);
                    c = pow(-1, syndrome[i]) * log((1 + c) / (1 - c));
                    s->c2b[e] = c;
                    s->b2c[e] = s->llr[j];
                    s->llr[j] += c;
                }
            } else {
                /* bp.hpp:502-523 */
                for (int p = g->col_ptr[j]; p < g->col_ptr[j + 1]; p++) {
                    int e = g->csc2csr[p];
                    int i = g->row_idx[p];
                    int sgn = syndrome[i];
                    double temp = DBL_MAX;
                    for (int f = g->row_ptr[i]; f < g->row_ptr[i + 1]; f++) {
                        if (f != e) {
                            double a = fabs(s->b2c[f]);
                            if (a < temp) temp = a;
                            if (s->b2c[f] <= 0) sgn += 1;
                        }
                    }
                    double message_sign = (sgn % 2 == 0) ? 1.0 : -1.0;
                    s->c2b[e] = alpha * message_sign * temp;
                    s->b2c[e] = s->llr[j];
                    s->llr[j] += s->c2b[e];
                }
            }
            /* bp.hpp:524-533 */
            s->decoding[j] = (s->llr[j] <= 0) ? 1 : 0;
            double temp = 0;
            for (int p = g->col_ptr[j + 1] - 1; p >= g->col_ptr[j]; p--) {
                int e = g->csc2csr[p];
                s->b2c[e] += temp;
                temp += s->c2b[e];
            }
        }
        /* bp.hpp:537-542 with gf2sparse.hpp:177-196 */
        for (int i = 0; i < g->m; i++) {
            uint8_t x = 0;
            for (int e = g->row_ptr[i]; e < g->row_ptr[i + 1]; e++) x ^= s->decoding[g->col_idx[e]];
            s->cand[i] = x;
        }
        s->iterations = it;
        if (memcmp(s->cand, syndrome, (size_t) g->m) == 0) {
            s->converge = 1;
            return;
        }
    }
}

/*
 * Decode a batch, one syndrome after another, exactly as B independent calls of
 * ldpc::bp::BpDecoder::decode(syndrome) would (bp.hpp:159-190, SYNDROME input).
 * `decoding` and `llr` persist between syndromes like the reference's members do
 * (matters only for serial orders that do not cover every bit).
 *
 * Returns 0, or <0 on bad arguments.
 */
int bpo_decode_batch(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, const double *channel,
                     int max_iter, int method, int schedule, double ms_scaling_factor, const int32_t *serial_order,
                     int serial_order_len, const uint8_t *syndromes, int64_t batch, uint8_t *out_decoding,
                     uint8_t *out_converged, int32_t *out_iters, double *out_llr) {
    bpo_graph g;
    int rc = bpo_graph_build(&g, m, n, nnz, rows, cols);
    if (rc) return rc;
    bpo_state s;
    memset(&s, 0, sizeof(s));
    s.g = &g;
    s.channel = channel;
    s.max_iter = max_iter;
    s.method = method;
    s.schedule = schedule;
    s.ms_scaling_factor = ms_scaling_factor;
    int32_t *ident = NULL;
    if (!serial_order) {
        ident = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n > 0 ? n : 1));
        for (int j = 0; j < n; j++) ident[j] = j;
        s.serial_order = ident;
        s.serial_order_len = n;
    } else {
        s.serial_order = serial_order;
        s.serial_order_len = serial_order_len;
        for (int k = 0; k < serial_order_len; k++)
            if (serial_order[k] < 0 || serial_order[k] >= n) {
                bpo_graph_free(&g);
                return -3;
            }
    }
    size_t ne = (size_t) (g.nnz > 0 ? g.nnz : 1);
    s.b2c = (double *) calloc(ne, sizeof(double));
    s.c2b = (double *) calloc(ne, sizeof(double));
    s.prior = (double *) calloc((size_t) n + 1, sizeof(double));
    s.llr = (double *) calloc((size_t) n + 1, sizeof(double));
    s.decoding = (uint8_t *) calloc((size_t) n + 1, 1);
    s.cand = (uint8_t *) calloc((size_t) m + 1, 1);
    s.order_work = (int *) calloc((size_t) s.serial_order_len + 1, sizeof(int));
    for (int64_t b = 0; b < batch; b++) {
        const uint8_t *syn = syndromes + b * (int64_t) m;
        s.iterations = 0;
        /* SERIAL_RELATIVE: every syndrome is decoded as by a freshly constructed decoder, i.e. starting from the
         * configured order (the reference object would carry the previous decode's final order over) */
        for (int k = 0; k < s.serial_order_len; k++) s.order_work[k] = s.serial_order[k];
        if (schedule == BPO_PARALLEL)
            bpo_parallel(&s, syn);
        else
            bpo_serial(&s, syn);
        memcpy(out_decoding + b * (int64_t) n, s.decoding, (size_t) n);
        if (out_converged) out_converged[b] = (uint8_t) s.converge;
        if (out_iters) out_iters[b] = s.iterations;
        if (out_llr) memcpy(out_llr + b * (int64_t) n, s.llr, sizeof(double) * (size_t) n);
    }
    free(s.b2c);
    free(s.c2b);
    free(s.prior);
    free(s.llr);
    free(s.decoding);
    free(s.cand);
    free(s.order_work);
    free(ident);
    bpo_graph_free(&g);
    return 0;
}

/*
 * soft_info_decode_serial (bp.hpp:547-665): serial min-sum with a SOFT syndrome.  soft_syndrome_i = 2 s_i / sigma^2,
 * hard syndrome bit = (soft_syndrome_i <= 0).  While a check's soft magnitude is below `cutoff` and below the minimum
 * incoming magnitude, the check acts as a "virtual" variable: it propagates its own magnitude, and is either refreshed
 * from the messages (when the parity of the incoming signs agrees with its hard bit) or flipped (:606-628).  The
 * convergence test compares H x with the (possibly flipped) hard syndrome (:649-660); once converged the remaining
 * iterations are skipped and `iterations` keeps the value of the converging sweep (:571-573,661).
 * Decodes a batch, one soft syndrome after another, as B independent calls would.
 */
int bpo_soft_info_decode_batch(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols,
                               const double *channel, int max_iter, double ms_scaling_factor,
                               const int32_t *serial_order, int serial_order_len, double cutoff, double sigma,
                               const double *soft_syndromes, int64_t batch, uint8_t *out_decoding,
                               uint8_t *out_converged, int32_t *out_iters, double *out_llr, double *out_soft) {
    bpo_graph g;
    int rc = bpo_graph_build(&g, m, n, nnz, rows, cols);
    if (rc) return rc;
    size_t ne = (size_t) (g.nnz > 0 ? g.nnz : 1);
    double *b2c = (double *) calloc(ne, sizeof(double)), *c2b = (double *) calloc(ne, sizeof(double));
    double *llr = (double *) calloc((size_t) n + 1, sizeof(double));
    double *soft = (double *) calloc((size_t) m + 1, sizeof(double));
    uint8_t *syn = (uint8_t *) calloc((size_t) m + 1, 1), *dec = (uint8_t *) calloc((size_t) n + 1, 1);
    for (int64_t b = 0; b < batch; b++) {
        const double *in = soft_syndromes + b * (int64_t) m;
        for (int i = 0; i < m; i++) { /* :551-558 */
            soft[i] = 2 * in[i] / (sigma * sigma);
            syn[i] = (soft[i] <= 0) ? 1 : 0;
        }
        for (int j = 0; j < n; j++) { /* initialise_log_domain_bp, bp.hpp:147-157 */
            double pr = log((1 - channel[j]) / channel[j]);
            for (int p = g.col_ptr[j]; p < g.col_ptr[j + 1]; p++) b2c[g.csc2csr[p]] = pr;
        }
        int converged = 0, iterations = 0;
        for (int it = 1; it <= max_iter; it++) {
            if (converged) continue; /* :571-573 */
            for (int oi = 0; oi < (serial_order ? serial_order_len : n); oi++) {
                int j = serial_order ? serial_order[oi] : oi;
                llr[j] = log((1 - channel[j]) / channel[j]);
                for (int p = g.col_ptr[j]; p < g.col_ptr[j + 1]; p++) {
                    int e = g.csc2csr[p], i = g.row_idx[p];
                    int sgn = 0;
                    double temp = DBL_MAX;
                    for (int f = g.row_ptr[i]; f < g.row_ptr[i + 1]; f++) {
                        if (f == e) continue;
                        if (fabs(b2c[f]) < temp) temp = fabs(b2c[f]);
                        if (b2c[f] <= 0) sgn ^= 1;
                    }
                    double min_msg = temp, propagated = min_msg;
                    double mag = fabs(soft[i]);
                    if (mag < cutoff) { /* :604-628 */
                        if (mag < fabs(min_msg)) {
                            propagated = mag;
                            int check_sgn = sgn;
                            if (b2c[e] <= 0) check_sgn ^= 1;
                            if (check_sgn == syn[i]) {
                                if (fabs(b2c[e]) < min_msg)
                                    soft[i] = pow(-1, syn[i]) * fabs(b2c[e]);
                                else
                                    soft[i] = pow(-1, syn[i]) * min_msg;
                            } else {
                                syn[i] ^= 1;
                                soft[i] *= -1;
                            }
                        }
                    }
                    sgn ^= syn[i];
                    c2b[e] = ms_scaling_factor * pow(-1, sgn) * propagated; /* :631 */
                    b2c[e] = llr[j];
                    llr[j] += c2b[e];
                }
                dec[j] = (llr[j] <= 0) ? 1 : 0;
                double t = 0;
                for (int p = g.col_ptr[j + 1] - 1; p >= g.col_ptr[j]; p--) {
                    int e = g.csc2csr[p];
                    b2c[e] += t;
                    t += c2b[e];
                }
            }
            converged = 1; /* :646-660 */
            for (int i = 0; i < m; i++) {
                uint8_t x = 0;
                for (int f = g.row_ptr[i]; f < g.row_ptr[i + 1]; f++) x ^= dec[g.col_idx[f]];
                if (x != syn[i]) {
                    converged = 0;
                    break;
                }
            }
            iterations = it;
        }
        memcpy(out_decoding + b * (int64_t) n, dec, (size_t) n);
        if (out_converged) out_converged[b] = (uint8_t) converged;
        if (out_iters) out_iters[b] = iterations;
        if (out_llr) memcpy(out_llr + b * (int64_t) n, llr, sizeof(double) * (size_t) n);
        if (out_soft) memcpy(out_soft + b * (int64_t) m, soft, sizeof(double) * (size_t) m);
    }
    free(b2c);
    free(c2b);
    free(llr);
    free(soft);
    free(syn);
    free(dec);
    bpo_graph_free(&g);
    return 0;
}

/* gf2sparse.hpp:177-196: out = H v (mod 2), used for received-vector input (bp.hpp:164). */
int bpo_mulvec(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, const uint8_t *vecs,
               int64_t batch, uint8_t *out) {
    (void) n;
    memset(out, 0, (size_t) (batch * m));
    for (int64_t b = 0; b < batch; b++)
        for (int64_t k = 0; k < nnz; k++) out[b * m + rows[k]] ^= (vecs[b * (int64_t) n + cols[k]] != 0);
    return 0;
}
