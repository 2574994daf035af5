// stl_sort.h -- libstdc++'s std::sort, restated, for the SERIAL_RELATIVE schedule.
//
// The reference re-sorts serial_schedule_order in every iteration with
//     std::sort(order.begin(), order.end(), [](int a, int b) { return llr[a] > llr[b]; })
// (reference src_cpp/bp.hpp:469-482).  std::sort is not stable, and min-sum on a regular code produces many equal
// LLRs (in the first iteration ALL keys are equal), so the permutation that comes out for tied keys is a property of
// the algorithm and of the order that went in.  Matching the reference's schedule -- hence its messages, decisions and
// iteration counts -- therefore needs the same algorithm.  libstdc++ (GCC 13, bits/stl_algo.h and bits/stl_heap.h; a
// third-party dependency of the reference that is absent from /root/reference) implements std::sort as introsort:
//   * __introsort_loop: while the range is longer than 16: median of (first+1, middle, last-1) moved to first,
//     unguarded Hoare partition around *first, recurse on the right part, loop on the left part; after 2*floor(log2 n)
//     levels fall back to heapsort (__partial_sort = __make_heap + __sort_heap);
//   * __final_insertion_sort: guarded insertion sort of the first 16 elements, unguarded linear inserts for the rest.
// This header restates exactly that, iteratively (explicit stack instead of the recursion on the right part), for
// int keys-by-index with the comparator comp(a, b) = key[a] > key[b].  It compiles for the host (the CPU test
// tests/native/stl_sort_check.cpp compares it with the real std::sort on tie-heavy inputs) and for the device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SS_FN __host__ __device__ __forceinline__
#else
#define SS_FN inline
#endif

namespace stlsort {

// Ord: random-access array of indices (int / uint16_t ...), Key: random-access array of doubles
template <class Ord, class Key>
struct Sorter {
    Ord a;
    Key key;
    SS_FN bool comp_idx(int x, int y) const { return key[x] > key[y]; }   // on index values
    SS_FN bool comp(int i, int j) const { return comp_idx((int) a[i], (int) a[j]); }  // on positions
    SS_FN void swap(int i, int j) {
        const auto t = a[i];
        a[i] = a[j];
        a[j] = t;
    }

    // bits/stl_algo.h: __move_median_to_first(result, a, b, c)
    SS_FN void move_median_to_first(int result, int x, int y, int z) {
        if (comp(x, y)) {
            if (comp(y, z)) swap(result, y);
            else if (comp(x, z)) swap(result, z);
            else swap(result, x);
        } else if (comp(x, z)) swap(result, x);
        else if (comp(y, z)) swap(result, z);
        else swap(result, y);
    }
    // __unguarded_partition(first, last, pivot)
    SS_FN int unguarded_partition(int first, int last, int pivot) {
        for (;;) {
            while (comp(first, pivot)) ++first;
            --last;
            while (comp(pivot, last)) --last;
            if (!(first < last)) return first;
            swap(first, last);
            ++first;
        }
    }
    // bits/stl_heap.h: __push_heap / __adjust_heap / __make_heap / __pop_heap / __sort_heap on [first, first+len)
    SS_FN void push_heap(int first, int hole, int top, int value) {
        int parent = (hole - 1) / 2;
        while (hole > top && comp_idx((int) a[first + parent], value)) {
            a[first + hole] = a[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        a[first + hole] = value;
    }
    SS_FN void adjust_heap(int first, int hole, int len, int value) {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (comp(first + child, first + (child - 1))) child--;
            a[first + hole] = a[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[first + hole] = a[first + (child - 1)];
            hole = child - 1;
        }
        push_heap(first, hole, top, value);
    }
    SS_FN void heap_sort(int first, int last) {  // __partial_sort(first, last, last)
        const int len = last - first;
        if (len >= 2) {  // __make_heap
            int parent = (len - 2) / 2;
            for (;;) {
                const int value = (int) a[first + parent];
                adjust_heap(first, parent, len, value);
                if (parent == 0) break;
                parent--;
            }
        }
        int end = last;
        while (end - first > 1) {  // __sort_heap: __pop_heap(first, end - 1, end - 1)
            --end;
            const int value = (int) a[end];
            a[end] = a[first];
            adjust_heap(first, 0, end - first, value);
        }
    }
    // __unguarded_linear_insert(last)
    SS_FN void unguarded_linear_insert(int last) {
        const int val = (int) a[last];
        int next = last - 1;
        while (comp_idx(val, (int) a[next])) {
            a[last] = a[next];
            last = next;
            --next;
        }
        a[last] = val;
    }
    // __insertion_sort(first, last)
    SS_FN void insertion_sort(int first, int last) {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (comp(i, first)) {
                const int val = (int) a[i];
                for (int q = i; q > first; --q) a[q] = a[q - 1];  // move_backward(first, i, i + 1)
                a[first] = val;
            } else {
                unguarded_linear_insert(i);
            }
        }
    }

    // std::sort(a, a + n, comp).  `stack` holds the pending right-hand ranges (begin, end, depth_limit): at most
    // 2*floor(log2 n) + 1 entries of 3 ints.
    SS_FN void sort(int n, int *stack) {
        if (n <= 0) return;
        int lg = 0;
        while ((n >> (lg + 1)) != 0) lg++;  // std::__lg(n)
        int sp = 0;
        int first = 0, last = n, depth = 2 * lg;
        for (;;) {
            // __introsort_loop(first, last, depth)
            while (last - first > 16) {
                if (depth == 0) {
                    heap_sort(first, last);
                    break;
                }
                --depth;
                const int mid = first + (last - first) / 2;
                move_median_to_first(first, first + 1, mid, last - 1);
                const int cut = unguarded_partition(first + 1, last, first);
                // the recursive call takes [cut, last) FIRST, then the loop continues with [first, cut): the two ranges
                // are disjoint, so the order in which they are processed does not change the result; the left one is
                // done now and the right one is stacked
                stack[sp++] = cut;
                stack[sp++] = last;
                stack[sp++] = depth;
                last = cut;
            }
            if (sp == 0) break;
            depth = stack[--sp];
            last = stack[--sp];
            first = stack[--sp];
        }
        // __final_insertion_sort(0, n)
        if (n > 16) {
            insertion_sort(0, 16);
            for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
        } else {
            insertion_sort(0, n);
        }
    }
};

constexpr int kStackInts = 3 * (2 * 31 + 2);

}  // namespace stlsort
