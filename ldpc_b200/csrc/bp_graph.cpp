// bp_graph.cpp -- flatten H into the reference's traversal order.
//
// The reference builds a doubly linked sparse matrix by calling insert_entry for every nonzero
// (src_python/ldpc/bp_decoder/_bp_decoder.pyx:36-47), which keeps each row sorted by column and each
// column sorted by row (src_cpp/sparse_matrix_base.hpp:423-482).  The kernels only need that order, so H
// becomes a sorted CSR plus a sorted CSC expressed as a permutation into the CSR edge numbering.
#include <algorithm>

#include "bp_decoder.h"

namespace bpb {

int build_host_graph(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, HostGraph &g,
                     std::string &err) {
    if ((int64_t) m * (int64_t) n >= ((int64_t) 1 << 62)) {
        err = "matrix too large";
        return BPB_ERR_ARG;
    }
    std::vector<int64_t> key((size_t) nnz);
    for (int64_t k = 0; k < nnz; k++) {
        if (rows[k] < 0 || rows[k] >= m || cols[k] < 0 || cols[k] >= n) {
            err = "matrix coordinate out of range";  // sparse_matrix_base.hpp:428
            return BPB_ERR_ARG;
        }
        key[(size_t) k] = (int64_t) rows[k] * n + cols[k];
    }
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    if (key.size() >= ((size_t) 1 << 31)) {
        err = "too many nonzeros";
        return BPB_ERR_ARG;
    }
    g.m = m;
    g.n = n;
    g.nnz = (int) key.size();
    g.row_ptr.assign((size_t) m + 1, 0u);
    g.col_ptr.assign((size_t) n + 1, 0u);
    g.col_idx.resize(key.size());
    g.row_idx.resize(key.size());
    g.csc2csr.resize(key.size());
    for (size_t e = 0; e < key.size(); e++) {
        const int r = (int) (key[e] / n), c = (int) (key[e] % n);
        g.row_ptr[(size_t) r + 1]++;
        g.col_ptr[(size_t) c + 1]++;
        g.col_idx[e] = (uint32_t) c;
    }
    for (int i = 0; i < m; i++) {
        g.max_row_degree = std::max(g.max_row_degree, (int) g.row_ptr[(size_t) i + 1]);
        g.row_ptr[(size_t) i + 1] += g.row_ptr[(size_t) i];
    }
    for (int j = 0; j < n; j++) {
        g.max_col_degree = std::max(g.max_col_degree, (int) g.col_ptr[(size_t) j + 1]);
        g.col_ptr[(size_t) j + 1] += g.col_ptr[(size_t) j];
    }
    g.regular = g.nnz > 0;
    for (int i = 0; i < m && g.regular; i++)
        if ((int) (g.row_ptr[(size_t) i + 1] - g.row_ptr[(size_t) i]) != g.max_row_degree) g.regular = false;
    for (int j = 0; j < n && g.regular; j++)
        if ((int) (g.col_ptr[(size_t) j + 1] - g.col_ptr[(size_t) j]) != g.max_col_degree) g.regular = false;
    std::vector<uint32_t> fill((size_t) n, 0u);
    for (size_t e = 0; e < key.size(); e++) {  // ascending row order => each column receives ascending rows
        const int r = (int) (key[e] / n), c = (int) (key[e] % n);
        const uint32_t pos = g.col_ptr[(size_t) c] + fill[(size_t) c]++;
        g.row_idx[pos] = (uint32_t) r;
        g.csc2csr[pos] = (uint32_t) e;
    }
    return BPB_OK;
}

}  // namespace bpb
