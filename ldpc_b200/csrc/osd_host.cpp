// osd_host.cpp -- OSD-0 post-processing for BP non-convergers, on the host.
//
// BASELINE.json's north_star keeps "OSD-0's fast-syndrome Gaussian elimination" on the host as the fallback
// for syndromes belief propagation did not solve.  It replaces ldpc::osd::OsdDecoder::decode with
// osd_order == 0 (reference src_cpp/osd.hpp:110-117):
//   (1) soft_decision_col_sort (src_cpp/sort.hpp:48-62): order the columns by ascending posterior LLR with
//       libc qsort on {double, int} records.  The order of tied LLRs is whatever libc's qsort produces with
//       that comparator, so this file makes the same qsort call on the same record layout;
//   (2) RowReduce::fast_solve (src_cpp/gf2sparse_linalg.hpp:298-401): eliminate columns in that order until
//       the syndrome lies in the span of the pivot columns, then solve on the pivots, all other bits zero.
// The reference does (2) with a linked-list sparse LU that it rebuilds for every call.  The answer does not
// depend on the elimination details: the pivot columns are the greedy independent set of the ordering, and a
// solution supported on independent columns is unique.  So this is a from-scratch bit-packed column
// elimination: each incoming column is reduced against the pivots found so far (64 rows per XOR), the
// syndrome is reduced alongside, and the combination is unwound at the end.
//
// Syndromes OUTSIDE the image of H (possible only with redundant checks and measurement noise that H does not model;
// s = H e is always inside): the sweep ends with a non-zero residual.  The reference then returns the solution of the
// equations of ITS pivot rows, which it picks by smallest row weight with ties broken by the order of a linked list
// that row swaps leave unsorted (gf2sparse_linalg.hpp:331-353, sparse_matrix_base.hpp:284-300) -- an artefact of its
// data structure, not of OSD.  Here (and in the device kernel, osd_device.cu) the result for such a syndrome is the
// unique solution on the same pivot COLUMNS of the equations of the lowest-index independent rows; it differs from
// the reference's in general.  osd0_host counts these syndromes (returned through `inconsistent`), the Python layer
// exposes the count, and tests/test_host_api.py::test_osd0_random_syndromes_bb144 pins both behaviours.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "bp_decoder.h"

namespace bpb {
namespace {

struct SortRec {  // same layout as the reference's record (sort.hpp:10-13)
    double value;
    int index;
};

int cmp_rec(const void *a, const void *b) {  // sort.hpp:36-46
    const SortRec *x = (const SortRec *) a, *y = (const SortRec *) b;
    if (x->value > y->value) return 1;
    if (x->value < y->value) return -1;
    return 0;
}

struct Workspace {
    int m, n, mw;                // mw = 64-bit words per m-bit vector
    std::vector<SortRec> recs;
    std::vector<uint64_t> piv;   // [rank][mw] reduced pivot vectors
    std::vector<uint64_t> used;  // [rank][mw] bitset over earlier pivots used to reduce this column
    std::vector<int> piv_row, piv_col;
    std::vector<uint64_t> v, vu, y, yu;
    Workspace(int m_, int n_) : m(m_), n(n_), mw((m_ + 63) / 64) {
        recs.resize((size_t) n);
        piv.resize((size_t) m * mw);
        used.resize((size_t) m * mw);
        piv_row.resize((size_t) m);
        piv_col.resize((size_t) m);
        v.resize((size_t) mw);
        vu.resize((size_t) mw);
        y.resize((size_t) mw);
        yu.resize((size_t) mw);
    }
};

inline bool any_set(const std::vector<uint64_t> &a) {
    for (uint64_t w: a)
        if (w) return true;
    return false;
}

// returns true when the syndrome turned out to lie outside the image of H
bool osd0_one(const HostGraph &g, Workspace &ws, const uint8_t *syn, const double *llr, uint8_t *out) {
    const int m = ws.m, n = ws.n, mw = ws.mw;
    for (int j = 0; j < n; j++) {
        ws.recs[(size_t) j].value = llr[j];
        ws.recs[(size_t) j].index = j;
    }
    qsort(ws.recs.data(), (size_t) n, sizeof(SortRec), cmp_rec);
    std::fill(ws.y.begin(), ws.y.end(), 0ull);
    std::fill(ws.yu.begin(), ws.yu.end(), 0ull);
    for (int i = 0; i < m; i++)
        if (syn[i]) ws.y[(size_t) (i >> 6)] |= 1ull << (i & 63);
    std::memset(out, 0, (size_t) n);
    int rank = 0;
    const int max_rank = std::min(m, n);
    for (int jj = 0; jj < n && rank < max_rank; jj++) {
        const int c = ws.recs[(size_t) jj].index;
        std::fill(ws.v.begin(), ws.v.end(), 0ull);
        std::fill(ws.vu.begin(), ws.vu.end(), 0ull);
        for (uint32_t p = g.col_ptr[(size_t) c]; p < g.col_ptr[(size_t) c + 1]; p++) {
            const uint32_t r = g.row_idx[p];
            ws.v[r >> 6] |= 1ull << (r & 63);
        }
        for (int k = 0; k < rank; k++) {
            const int r = ws.piv_row[(size_t) k];
            if ((ws.v[(size_t) (r >> 6)] >> (r & 63)) & 1ull) {
                const uint64_t *pk = &ws.piv[(size_t) k * mw];
                for (int w = 0; w < mw; w++) ws.v[(size_t) w] ^= pk[w];
                ws.vu[(size_t) (k >> 6)] |= 1ull << (k & 63);
            }
        }
        int r = -1;
        for (int w = 0; w < mw; w++)
            if (ws.v[(size_t) w]) {
                r = w * 64 + __builtin_ctzll(ws.v[(size_t) w]);
                break;
            }
        if (r < 0) continue;  // dependent on earlier columns: not a pivot
        std::memcpy(&ws.piv[(size_t) rank * mw], ws.v.data(), sizeof(uint64_t) * (size_t) mw);
        std::memcpy(&ws.used[(size_t) rank * mw], ws.vu.data(), sizeof(uint64_t) * (size_t) mw);
        ws.piv_row[(size_t) rank] = r;
        ws.piv_col[(size_t) rank] = c;
        if ((ws.y[(size_t) (r >> 6)] >> (r & 63)) & 1ull) {
            for (int w = 0; w < mw; w++) ws.y[(size_t) w] ^= ws.v[(size_t) w];
            ws.yu[(size_t) (rank >> 6)] |= 1ull << (rank & 63);
        }
        rank++;
        if (!any_set(ws.y)) break;  // syndrome is in the image (gf2sparse_linalg.hpp:373-383)
    }
    // unwind: y = XOR_{k in yu} piv_k and piv_k = col_k XOR XOR_{i in used_k} piv_i
    for (int k = rank - 1; k >= 0; k--) {
        if ((ws.yu[(size_t) (k >> 6)] >> (k & 63)) & 1ull) {
            out[ws.piv_col[(size_t) k]] = 1;
            const uint64_t *uk = &ws.used[(size_t) k * mw];
            for (int w = 0; w <= (k >> 6); w++) ws.yu[(size_t) w] ^= uk[w];
        }
    }
    return any_set(ws.y);
}

}  // namespace

int osd0_host(const HostGraph &g, const uint8_t *syndromes, const double *llr, const uint8_t *converged,
              int64_t batch, uint8_t *decoding, int threads, int64_t *inconsistent) {
    std::atomic<int64_t> outside{0};
    std::vector<int64_t> todo;
    for (int64_t b = 0; b < batch; b++)
        if (!converged || !converged[b]) todo.push_back(b);
    if (todo.empty()) return BPB_OK;
    if (threads <= 0) threads = (int) std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if ((size_t) threads > todo.size()) threads = (int) todo.size();
    auto run = [&](int t) {
        Workspace ws(g.m, g.n);
        for (size_t q = (size_t) t; q < todo.size(); q += (size_t) threads) {
            const int64_t b = todo[q];
            if (osd0_one(g, ws, syndromes + b * g.m, llr + b * g.n, decoding + b * g.n)) outside++;
        }
    };
    if (threads == 1) {
        run(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(run, t);
        for (auto &th: pool) th.join();
    }
    if (inconsistent) *inconsistent = outside.load();
    return BPB_OK;
}

}  // namespace bpb
