// osd_order.h -- the column order of OSD-0: glibc's qsort merge tree on (llr, index) records, one node at a time.
//
// The reference orders the columns with libc qsort on {double value; int index} records and the comparator of
// src_cpp/sort.hpp:36-46 (1 if a > b, -1 if a < b, else 0; src_cpp/sort.hpp:48-62).  glibc (2.39 on this image; the
// same since the merge sort was reinstated) sorts records of this size with a top-down merge sort:
//     msort(b, n): n1 = n / 2, n2 = n - n1; msort(b, n1); msort(b + n1, n2); merge, taking from the LEFT run while
//     cmp(left, right) <= 0.
// A third-party dependency absent from /root/reference, restated here.  With a consistent comparator that is a stable
// sort; with NaN keys (which compare "equal" to everything) the result is whatever this merge tree produces, so the
// device code walks the same tree: level by level from the leaves, the nodes of one level independent of each other.
// Plain C, usable from CUDA device code and from the host check (tests/native/osd_order_check.c compares it with
// the live libc qsort).
#ifndef BPB_OSD_ORDER_H
#define BPB_OSD_ORDER_H
#include <stdint.h>

#ifdef __CUDACC__
#define BPB_HD __host__ __device__ __forceinline__
#else
#define BPB_HD static inline
#endif

/* levels of the tree: smallest d with 2^d >= n */
BPB_HD int osd_order_depth(int n) {
    int d = 0;
    while ((1 << d) < n) d++;
    return d;
}

/* range [lo, lo+len) of node k (0 .. 2^d - 1, left to right) at depth d of the split n/2 | n - n/2 */
BPB_HD void osd_order_node(int n, int d, int k, int *lo_out, int *len_out) {
    int lo = 0, len = n;
    for (int bit = d - 1; bit >= 0; --bit) {
        const int half = len >> 1;
        if ((k >> bit) & 1) {
            lo += half;
            len -= half;
        } else {
            len = half;
        }
    }
    *lo_out = lo;
    *len_out = len;
}

/* merge the two sorted runs src[lo, lo+len/2) and src[lo+len/2, lo+len) of column indices into dst[lo, lo+len),
 * comparing the columns' values; the left element goes first unless value[left] > value[right] */
BPB_HD void osd_order_merge(const uint16_t *src, uint16_t *dst, const double *value, int lo, int len) {
    if (len <= 0) return;
    if (len == 1) {
        dst[lo] = src[lo];
        return;
    }
    int i = lo, e1 = lo + (len >> 1), j = e1, e2 = lo + len, o = lo;
    uint16_t a = src[i], c = src[j];
    double va = value[a], vc = value[c];
    for (;;) {
        if (va > vc) {
            dst[o++] = c;
            if (++j == e2) break;
            c = src[j];
            vc = value[c];
        } else {
            dst[o++] = a;
            if (++i == e1) break;
            a = src[i];
            va = value[a];
        }
    }
    while (i < e1) dst[o++] = src[i++];
    while (j < e2) dst[o++] = src[j++];
}
#endif
