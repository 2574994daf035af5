// osd_device.cu -- OSD-0 post-processing for BP non-convergers ON THE DEVICE.
//
// Replaces ldpc::osd::OsdDecoder::decode with osd_order == 0 (reference src_cpp/osd.hpp:110-117) for a list of
// syndromes that belief propagation did not solve, without leaving the GPU:
//   (1) soft_decision_col_sort (src_cpp/sort.hpp:48-62): the reference orders the columns with libc qsort on
//       {double value; int index} records and the comparator of sort.hpp:36-46.  glibc's qsort is a top-down merge
//       sort (split n/2 | n - n/2, take from the left run unless left > right), so the order of tied LLRs -- and what
//       happens to NaNs, which compare "equal" to everything -- is a property of that merge tree.  The kernel runs the
//       same merge tree (one tree level at a time, the nodes of a level in parallel over the lanes, each merge
//       sequential), so the column order is the reference's for every input, ties and NaNs included
//       (tests/test_host_api.py::test_osd_column_order_is_the_libc_qsort_merge_tree pins the tree against the
//       live libc);
//   (2) RowReduce::fast_solve (src_cpp/gf2sparse_linalg.hpp:298-401): eliminate columns in that order until the
//       syndrome lies in the span of the pivot columns, solve on the pivots, all other bits zero.  The solution does
//       not depend on the elimination details (the pivot columns are the greedy independent set of the ordering, a
//       solution supported on independent columns is unique), so this is a bit-packed Gauss-Jordan elimination of
//       the column-permuted matrix [H P | s] held in shared memory.
//
// One warp = one syndrome; lanes own rows (row r belongs to lane r % 32), a row is W 32-bit words with an odd word
// stride (conflict-free when every lane touches the same word index of its own row).  Per column of the ordering:
// every lane tests the column's bit in its rows, a warp-wide minimum picks the pivot row, every lane XORs the pivot
// row (a shared-memory broadcast) into its rows that have the bit.  The loop stops as soon as no unused row has its
// syndrome bit set (gf2sparse_linalg.hpp:373-383).
//
// Codes whose matrix does not fit (m > 1024 or m * ceil((n+1)/32) words beyond the shared memory of an SM, e.g.
// n = 10^4) keep the host elimination (osd_host.cpp).
#include <algorithm>

#include "bp_decoder.h"
#include "osd_order.h"

namespace bpb {

struct OsdParams {
    const uint32_t *row_ptr, *col_idx;  // CSR of H (ascending columns)
    int m, n, mwp;
    int depth;         // ceil(log2 n): levels of the merge tree
    int ws;            // words per matrix row (odd)
    int rows_per_lane; // ceil(m / 32)
    uint32_t warp_bytes, off_a, off_b, off_inv, off_piv;  // per-warp shared-memory layout
    const uint32_t *synd_packed;        // [B][mwp]
    const double *llr;                  // [B][n] posterior LLRs written by the BP kernel
    const uint32_t *fail_idx;           // batch indices of the syndromes to solve
    const unsigned long long *count;    // how many (device value)
    unsigned long long *counter;        // work queue
    uint8_t *out_dec;                   // [B][n], rows of the listed syndromes are overwritten
};

__global__ void __launch_bounds__(512) osd0_kernel(const OsdParams p) {
    extern __shared__ __align__(16) uint8_t osd_sm[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    uint8_t *base = osd_sm + (size_t) wib * p.warp_bytes;
    double *llr_s = reinterpret_cast<double *>(base);   // during the sort
    uint32_t *M = reinterpret_cast<uint32_t *>(base);   // afterwards: the bit matrix (same storage)
    uint16_t *buf_a = reinterpret_cast<uint16_t *>(base + p.off_a);
    uint16_t *buf_b = reinterpret_cast<uint16_t *>(base + p.off_b);
    uint16_t *inv = reinterpret_cast<uint16_t *>(base + p.off_inv);
    uint16_t *piv = reinterpret_cast<uint16_t *>(base + p.off_piv);
    const int m = p.m, n = p.n, ws = p.ws, R = p.rows_per_lane;
    const unsigned long long total = *p.count;
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t NONE = 0xffffffffu;

    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(p.counter, 1ull);
        q = __shfl_sync(FULL, q, 0);
        if (q >= total) break;
        const size_t b = p.fail_idx[q];

        // ---- (1) column order: the merge tree of glibc's qsort on (llr, index) ------------------------------
        const double *lg = p.llr + b * (size_t) n;
        for (int j = lane; j < n; j += 32) {
            llr_s[j] = lg[j];
            buf_a[j] = (uint16_t) j;
        }
        __syncwarp();
        uint16_t *src = buf_a, *dst = buf_b;
        for (int d = p.depth - 1; d >= 0; --d) {
            const int nodes = 1 << d;
            for (int k = lane; k < nodes; k += 32) {
                int lo, len;
                osd_order_node(n, d, k, &lo, &len);
                osd_order_merge(src, dst, llr_s, lo, len);
            }
            __syncwarp();
            uint16_t *t = src;
            src = dst;
            dst = t;
        }
        const uint16_t *perm = src;  // perm[pos] = column at position pos of the ordering

        // ---- (2) [H P | s] as a bit matrix ------------------------------------------------------------------
        for (int pos = lane; pos < n; pos += 32) inv[perm[pos]] = (uint16_t) pos;
        for (int x = lane; x < m * ws; x += 32) M[x] = 0u;  // overwrites llr_s: the sort is done
        __syncwarp();
        const uint32_t *srow = p.synd_packed + b * (size_t) p.mwp;
        const int sw = n >> 5;
        const uint32_t sbit = 1u << (n & 31);
        for (int r = lane; r < m; r += 32) {
            uint32_t *row = M + (size_t) r * ws;
            for (uint32_t e = p.row_ptr[r]; e < p.row_ptr[r + 1]; ++e) {
                const uint32_t pos = inv[p.col_idx[e]];
                row[pos >> 5] |= 1u << (pos & 31);
            }
            if ((__ldg(srow + (r >> 5)) >> (r & 31)) & 1u) row[sw] |= sbit;
        }
        __syncwarp();

        // ---- (3) Gauss-Jordan in column order until the syndrome is in the span of the pivots ----------------
        uint32_t used = 0;  // bit k: my row lane + 32 k is a pivot row
        int rank = 0;
        const int max_rank = m < n ? m : n;
        bool mypend = false;
        for (int k = 0; k < R; ++k) {
            const int r = lane + 32 * k;
            if (r < m) mypend |= (M[(size_t) r * ws + sw] & sbit) != 0;
        }
        bool pending = __any_sync(FULL, mypend);
        for (int pos = 0; pos < n && rank < max_rank && pending; ++pos) {
            const int w = pos >> 5;
            const uint32_t bit = 1u << (pos & 31);
            uint32_t setmask = 0, mine = NONE;
            for (int k = 0; k < R; ++k) {
                const int r = lane + 32 * k;
                if (r < m && (M[(size_t) r * ws + w] & bit)) {
                    setmask |= 1u << k;
                    if (!((used >> k) & 1u) && mine == NONE) mine = (uint32_t) r;
                }
            }
            const uint32_t pr = __reduce_min_sync(FULL, mine);
            if (pr == NONE) continue;  // dependent on the pivots so far
            if ((int) (pr & 31u) == lane) {
                used |= 1u << (pr >> 5);
                piv[pr] = (uint16_t) pos;
            }
            rank++;
            const uint32_t *prow = M + (size_t) pr * ws;
            while (setmask) {
                const int k = __ffs(setmask) - 1;
                setmask &= setmask - 1;
                const uint32_t r = (uint32_t) (lane + 32 * k);
                if (r == pr) continue;
                uint32_t *row = M + (size_t) r * ws;
                for (int x = w; x < ws; ++x) row[x] ^= prow[x];
            }
            __syncwarp();
            mypend = false;
            for (int k = 0; k < R; ++k) {
                const int r = lane + 32 * k;
                if (r < m && !((used >> k) & 1u)) mypend |= (M[(size_t) r * ws + sw] & sbit) != 0;
            }
            pending = __any_sync(FULL, mypend);
        }

        // ---- (4) x = reduced syndrome on the pivot columns, zero elsewhere -----------------------------------
        uint8_t *drow = p.out_dec + b * (size_t) n;
        if ((n & 3) == 0) {
            uint32_t *o32 = reinterpret_cast<uint32_t *>(drow);
            for (int x = lane; x < (n >> 2); x += 32) o32[x] = 0u;
        } else {
            for (int j = lane; j < n; j += 32) drow[j] = 0;
        }
        __syncwarp();
        for (int k = 0; k < R; ++k) {
            const int r = lane + 32 * k;
            if (r < m && ((used >> k) & 1u) && (M[(size_t) r * ws + sw] & sbit)) drow[perm[piv[r]]] = 1;
        }
        __syncwarp();
    }
}

// Shared-memory plan of the kernel for this code; warps_per_cta == 0 when the device path cannot take it.
OsdDevicePlan plan_osd_device(const HostGraph &g, int max_smem_optin) {
    OsdDevicePlan pl;
    const int m = g.m, n = g.n;
    if (m > 1024 || n > 65535 || n < 1) return pl;
    pl.depth = osd_order_depth(n);
    pl.ws = ((n + 1 + 31) / 32) | 1;
    pl.rows_per_lane = (m + 31) / 32;
    auto up = [](size_t x, size_t q) { return (x + q - 1) / q * q; };
    size_t off = up(std::max((size_t) n * 8, (size_t) m * pl.ws * 4), 16);
    pl.off_a = (uint32_t) off;
    off += up((size_t) n * 2, 16);
    pl.off_b = (uint32_t) off;
    off += up((size_t) n * 2, 16);
    pl.off_inv = (uint32_t) off;
    off += up((size_t) n * 2, 16);
    pl.off_piv = (uint32_t) off;
    off += up((size_t) m * 2, 16);
    pl.warp_bytes = (uint32_t) off;
    const size_t budget = (size_t) max_smem_optin;
    if (off > budget) return pl;
    // several CTAs per SM rather than one large one: warps finish at different times
    int warps = (int) std::min<size_t>(16, budget / off);
    while (warps > 4 && (size_t) warps * off > budget / 3) warps--;
    pl.warps_per_cta = std::max(1, warps);
    return pl;
}

int launch_osd0_kernel(const OsdDevicePlan &pl, const HostGraph &g, int sm_count, const uint32_t *d_row_ptr,
                       const uint32_t *d_col_idx, const uint32_t *d_packed, int mwp, const double *d_llr,
                       const uint32_t *d_fail_idx, const unsigned long long *d_count, unsigned long long *d_counter,
                       uint8_t *d_dec, int64_t max_items, cudaStream_t st) {
    OsdParams p{};
    p.row_ptr = d_row_ptr;
    p.col_idx = d_col_idx;
    p.m = g.m;
    p.n = g.n;
    p.mwp = mwp;
    p.depth = pl.depth;
    p.ws = pl.ws;
    p.rows_per_lane = pl.rows_per_lane;
    p.warp_bytes = pl.warp_bytes;
    p.off_a = pl.off_a;
    p.off_b = pl.off_b;
    p.off_inv = pl.off_inv;
    p.off_piv = pl.off_piv;
    p.synd_packed = d_packed;
    p.llr = d_llr;
    p.fail_idx = d_fail_idx;
    p.count = d_count;
    p.counter = d_counter;
    p.out_dec = d_dec;
    const size_t smem = (size_t) pl.warps_per_cta * pl.warp_bytes;
    cudaError_t e = cudaFuncSetAttribute(osd0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, osd0_kernel, pl.warps_per_cta * 32, smem);
    if (e != cudaSuccess) return (int) e;
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t) occ * sm_count;
    const int64_t need = (max_items + pl.warps_per_cta - 1) / pl.warps_per_cta;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    osd0_kernel<<<(int) grid, pl.warps_per_cta * 32, smem, st>>>(p);
    return (int) cudaGetLastError();
}

}  // namespace bpb
