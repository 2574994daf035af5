// bp_relative.cu -- the SERIAL_RELATIVE schedule: bp_decode_serial with the schedule re-sorted
// by posterior LLR before every sweep (reference src_cpp/bp.hpp:451-545, the sort at :469-482).
//
// The order of a sweep depends on the syndrome being decoded, so it cannot be levelised on the host the way the plain
// serial schedule is (bp_plan.cpp: build_serial_batches), and tied LLRs are ordered by whatever libstdc++'s introsort
// does with them, which has to be reproduced step by step (stl_sort.h).  One warp decodes one syndrome:
//   * lane 0 runs the std::sort restatement on the schedule held in shared memory (keys = the posterior LLRs of the
//     previous sweep, the priors in the first iteration);
//   * the sweep visits the bits in that order; for one bit, lane k computes the check-to-bit message of the bit's k-th
//     check from the current bit-to-check messages of the check's OTHER bits (bp.hpp:489-498 product-sum, :503-519
//     min-sum, in ascending column order), the running sums of bp.hpp:499-500,520-533 are taken in the reference's order
//     through warp shuffles, and lane k writes the new bit-to-check message of its edge;
//   * candidate syndrome and convergence test after every sweep (bp.hpp:537-542).
// Messages (one slot per edge, updated in place) live in shared memory when they fit next to the keys and the
// schedule, else in an L2-resident scratch in global memory.  Every syndrome starts from the configured
// serial_schedule_order; the order the last syndrome of a call ended with is returned (the reference object keeps it).
#include <algorithm>

#include "bp_decoder.h"
#include "bp_update.cuh"
#include "stl_sort.h"

namespace bpb {

struct RelParams {
    const uint32_t *row_ptr, *col_idx, *col_ptr, *csc2csr, *row_idx;
    const double *prior;
    int m, n, nnz, MW;
    int max_iter, method;
    double ms_scaling;
    const uint32_t *order0;  // configured schedule (order_len entries)
    int order_len;
    uint32_t off_order, off_dec, off_syn, off_stack, off_msg;  // byte offsets in dynamic shared memory (keys at 0)
    const uint32_t *synd_packed;
    int mwp;
    long long batch;
    unsigned long long *counter;
    double *msg_global;   // [grid][nnz] when the messages do not fit in shared memory
    uint8_t *out_dec;
    uint8_t *out_conv;
    int32_t *out_iters;
    double *out_llr;
    int llr_last_only;
    int32_t *out_order;   // [order_len]: schedule after the decode of syndrome batch-1 (or null)
};

template <int METHOD, bool MSG_GLOBAL>
__global__ void __launch_bounds__(32) bp_relative_kernel(const RelParams p) {
    extern __shared__ __align__(16) uint8_t rsm[];
    const int lane = threadIdx.x;
    const int m = p.m, n = p.n, nnz = p.nnz;
    double *key = reinterpret_cast<double *>(rsm);  // posterior LLR of every bit = the sort keys
    int *order = reinterpret_cast<int *>(rsm + p.off_order);
    uint8_t *dec = rsm + p.off_dec;
    uint32_t *synw = reinterpret_cast<uint32_t *>(rsm + p.off_syn);
    int *stack = reinterpret_cast<int *>(rsm + p.off_stack);
    double *msg = MSG_GLOBAL ? (p.msg_global + (size_t) blockIdx.x * (size_t) nnz)
                             : reinterpret_cast<double *>(rsm + p.off_msg);
    __shared__ long long ctl;
    for (;;) {
        if (lane == 0) {
            const long long claim = (long long) atomicAdd(p.counter, 1ull);
            ctl = claim < p.batch ? claim : -1;
        }
        __syncwarp();
        const long long idx = ctl;
        __syncwarp();
        if (idx < 0) break;
        const uint32_t *srow = p.synd_packed + idx * p.mwp;
        for (int w = lane; w < p.MW; w += 32) synw[w] = __ldg(srow + w);
        for (int j = lane; j < n; j += 32) {
            key[j] = p.prior[j];  // first iteration: sorted by the priors (bp.hpp:474-479)
            dec[j] = 0;
        }
        for (int e = lane; e < nnz; e += 32) msg[e] = p.prior[p.col_idx[e]];  // bp.hpp:147-157
        for (int k = lane; k < p.order_len; k += 32) order[k] = (int) p.order0[k];
        __syncwarp();
        int it = 0;
        bool conv = false;
        while (it < p.max_iter) {
            ++it;
            const double alpha = ms_alpha(p.ms_scaling, it);
            if (lane == 0) {
                stlsort::Sorter<int *, const double *> s{order, key};
                s.sort(p.order_len, stack);
            }
            __syncwarp();
            for (int oi = 0; oi < p.order_len; ++oi) {
                const int j = order[oi];
                const uint32_t cb = p.col_ptr[j];
                const int deg = (int) (p.col_ptr[j + 1] - cb);
                double c = 0.0;
                uint32_t e = 0;
                if (lane < deg) {
                    e = p.csc2csr[cb + lane];
                    const uint32_t i = p.row_idx[cb + lane];
                    const uint32_t rb = p.row_ptr[i], re = p.row_ptr[i + 1];
                    const uint32_t s = (synw[i >> 5] >> (i & 31)) & 1u;
                    if (METHOD == kMinimumSum) {  // bp.hpp:503-519
                        uint32_t sg = s;
                        double temp = DBL_MAX;
                        for (uint32_t f = rb; f < re; ++f) {
                            if (f == e) continue;
                            const double b = msg[f];
                            const double a = fabs(b);
                            if (a < temp) temp = a;
                            if (b <= 0) sg += 1;
                        }
                        c = ((sg & 1u) ? -alpha : alpha) * temp;
                    } else {  // bp.hpp:489-498
                        double x = 1.0;
                        for (uint32_t f = rb; f < re; ++f) {
                            if (f == e) continue;
                            x *= ps_tanh_half(msg[f]);
                        }
                        c = (s ? -1.0 : 1.0) * ps_atanh2(x);
                    }
                }
                __syncwarp();  // every lane has read its neighbours before any message of this bit is rewritten
                // bp.hpp:499-500 / 520-521: b_k := running sum, running sum += c_k (ascending rows);
                // bp.hpp:529-533: b_k += sum of the c of the later edges (descending rows)
                double t = p.prior[j], mypre = 0.0, mysuf = 0.0;
                for (int k = 0; k < deg; ++k) {
                    const double ck = __shfl_sync(0xffffffffu, c, k);
                    if (lane == k) mypre = t;
                    t += ck;
                }
                double u = 0.0;
                for (int k = deg - 1; k >= 0; --k) {
                    const double ck = __shfl_sync(0xffffffffu, c, k);
                    if (lane == k) mysuf = u;
                    u += ck;
                }
                if (lane < deg) msg[e] = mypre + mysuf;
                if (lane == 0) {
                    key[j] = t;  // log_prob_ratios[j]
                    dec[j] = (t <= 0) ? 1 : 0;
                }
                __syncwarp();
            }
            // candidate syndrome == syndrome ?  (bp.hpp:537-542)
            bool bad = false;
            for (int i = lane; i < m; i += 32) {
                uint32_t x = (synw[i >> 5] >> (i & 31)) & 1u;
                for (uint32_t f = p.row_ptr[i]; f < p.row_ptr[i + 1]; ++f) x ^= dec[p.col_idx[f]];
                bad |= (x != 0);
            }
            conv = !__any_sync(0xffffffffu, bad);
            if (conv) break;
        }
        uint8_t *drow = p.out_dec + idx * n;
        for (int j = lane; j < n; j += 32) drow[j] = dec[j];
        if (p.out_llr && !(p.llr_last_only && conv)) {
            double *lrow = p.out_llr + idx * n;
            for (int j = lane; j < n; j += 32) lrow[j] = key[j];
        }
        if (lane == 0) {
            if (p.out_iters) p.out_iters[idx] = it;
            if (p.out_conv) p.out_conv[idx] = conv ? 1 : 0;
        }
        if (p.out_order && idx == p.batch - 1)
            for (int k = lane; k < p.order_len; k += 32) p.out_order[k] = order[k];
        __syncwarp();
    }
}

// ---- soft-information serial min-sum (reference src_cpp/bp.hpp:547-665, SoftInfoBpDecoder) -------------------------
// The syndrome arrives as real numbers: soft_i = 2 s_i / sigma^2, hard bit = (soft_i <= 0).  A check whose soft
// magnitude is below `cutoff` and below the smallest incoming magnitude acts as a virtual variable node: it propagates
// its own magnitude and is refreshed from the messages or flipped (:604-628), so every check carries mutable state
// through the sweep.  Same mapping as the SERIAL_RELATIVE kernel: one warp per syndrome, bits in the configured
// order, lane k takes the bit's k-th check (the checks of one bit are distinct, so the lanes touch distinct state).
struct SoftParams {
    const uint32_t *row_ptr, *col_idx, *col_ptr, *csc2csr, *row_idx;
    const double *prior;
    int m, n, nnz;
    int max_iter;
    double ms_scaling, cutoff, sigma;
    const uint32_t *order0;
    int order_len;
    uint32_t off_soft, off_dec, off_syn, off_msg;  // byte offsets in dynamic shared memory (LLRs at 0)
    const double *soft_in;  // [B][m]
    long long batch;
    unsigned long long *counter;
    double *msg_global;
    uint8_t *out_dec;
    uint8_t *out_conv;
    int32_t *out_iters;
    double *out_llr;   // [B][n] or null
    double *out_soft;  // [B][m] or null: the soft syndrome after decoding
};

template <bool MSG_GLOBAL>
__global__ void __launch_bounds__(32) bp_softinfo_kernel(const SoftParams p) {
    extern __shared__ __align__(16) uint8_t ssm[];
    const int lane = threadIdx.x;
    const int m = p.m, n = p.n, nnz = p.nnz;
    double *llr = reinterpret_cast<double *>(ssm);
    double *soft = reinterpret_cast<double *>(ssm + p.off_soft);
    uint8_t *dec = ssm + p.off_dec;
    uint8_t *syn = ssm + p.off_syn;
    double *msg = MSG_GLOBAL ? (p.msg_global + (size_t) blockIdx.x * (size_t) nnz)
                             : reinterpret_cast<double *>(ssm + p.off_msg);
    __shared__ long long ctl;
    for (;;) {
        if (lane == 0) {
            const long long claim = (long long) atomicAdd(p.counter, 1ull);
            ctl = claim < p.batch ? claim : -1;
        }
        __syncwarp();
        const long long idx = ctl;
        __syncwarp();
        if (idx < 0) break;
        const double *in = p.soft_in + idx * m;
        for (int i = lane; i < m; i += 32) {  // bp.hpp:551-558
            const double sv = 2 * in[i] / (p.sigma * p.sigma);
            soft[i] = sv;
            syn[i] = (sv <= 0) ? 1 : 0;
        }
        for (int j = lane; j < n; j += 32) {
            llr[j] = p.prior[j];
            dec[j] = 0;
        }
        for (int e = lane; e < nnz; e += 32) msg[e] = p.prior[p.col_idx[e]];  // bp.hpp:147-157
        __syncwarp();
        int it = 0;
        bool conv = false;
        while (it < p.max_iter && !conv) {  // a converged decode skips its remaining iterations (:571-573)
            ++it;
            for (int oi = 0; oi < p.order_len; ++oi) {
                const int j = (int) p.order0[oi];
                const uint32_t cb = p.col_ptr[j];
                const int deg = (int) (p.col_ptr[j + 1] - cb);
                double c = 0.0;
                uint32_t e = 0;
                if (lane < deg) {
                    e = p.csc2csr[cb + lane];
                    const uint32_t i = p.row_idx[cb + lane];
                    const uint32_t rb = p.row_ptr[i], re = p.row_ptr[i + 1];
                    uint32_t sgn = 0;
                    double temp = DBL_MAX;
                    for (uint32_t f = rb; f < re; ++f) {  // :590-600
                        if (f == e) continue;
                        const double b = msg[f];
                        if (fabs(b) < temp) temp = fabs(b);
                        if (b <= 0) sgn ^= 1u;
                    }
                    const double min_msg = temp;
                    double propagated = min_msg;
                    const double own = msg[e];
                    uint32_t s = syn[i];
                    const double mag = fabs(soft[i]);
                    if (mag < p.cutoff && mag < fabs(min_msg)) {  // :604-628
                        propagated = mag;
                        const uint32_t check_sgn = sgn ^ ((own <= 0) ? 1u : 0u);
                        if (check_sgn == s) {
                            const double v = (fabs(own) < min_msg) ? fabs(own) : min_msg;
                            soft[i] = (s ? -1.0 : 1.0) * v;
                        } else {
                            s ^= 1u;
                            syn[i] = (uint8_t) s;
                            soft[i] = soft[i] * -1;
                        }
                    }
                    sgn ^= s;
                    c = (p.ms_scaling * (sgn ? -1.0 : 1.0)) * propagated;  // :631
                }
                __syncwarp();
                double t = p.prior[j], mypre = 0.0, mysuf = 0.0;
                for (int k = 0; k < deg; ++k) {
                    const double ck = __shfl_sync(0xffffffffu, c, k);
                    if (lane == k) mypre = t;
                    t += ck;
                }
                double u = 0.0;
                for (int k = deg - 1; k >= 0; --k) {
                    const double ck = __shfl_sync(0xffffffffu, c, k);
                    if (lane == k) mysuf = u;
                    u += ck;
                }
                if (lane < deg) msg[e] = mypre + mysuf;
                if (lane == 0) {
                    llr[j] = t;
                    dec[j] = (t <= 0) ? 1 : 0;
                }
                __syncwarp();
            }
            bool bad = false;  // :646-660: H x against the (possibly flipped) hard syndrome
            for (int i = lane; i < m; i += 32) {
                uint32_t x = syn[i];
                for (uint32_t f = p.row_ptr[i]; f < p.row_ptr[i + 1]; ++f) x ^= dec[p.col_idx[f]];
                bad |= (x != 0);
            }
            conv = !__any_sync(0xffffffffu, bad);
        }
        uint8_t *drow = p.out_dec + idx * n;
        for (int j = lane; j < n; j += 32) drow[j] = dec[j];
        if (p.out_llr)
            for (int j = lane; j < n; j += 32) p.out_llr[idx * n + j] = llr[j];
        if (p.out_soft)
            for (int i = lane; i < m; i += 32) p.out_soft[idx * m + i] = soft[i];
        if (lane == 0) {
            if (p.out_iters) p.out_iters[idx] = it;
            if (p.out_conv) p.out_conv[idx] = conv ? 1 : 0;
        }
        __syncwarp();
    }
}

int launch_softinfo_kernel(const HostGraph &g, int sm_count, int max_smem_optin, const uint32_t *d_blob,
                           uint32_t prior_off, int max_iter, double ms_scaling, double cutoff, double sigma,
                           const uint32_t *d_order0, int order_len, const double *d_soft, int64_t batch,
                           unsigned long long *d_counter, DeviceBuffer *scratch, uint8_t *d_dec, uint8_t *d_conv,
                           int32_t *d_iters, double *d_llr, double *d_soft_out, cudaStream_t st) {
    if (g.max_col_degree > 32) return -1;
    SoftParams p{};
    p.row_ptr = d_blob;
    p.col_idx = p.row_ptr + (g.m + 1);
    p.col_ptr = p.col_idx + g.nnz;
    p.csc2csr = p.col_ptr + (g.n + 1);
    p.row_idx = p.csc2csr + g.nnz;
    p.prior = reinterpret_cast<const double *>(d_blob + prior_off);
    p.m = g.m;
    p.n = g.n;
    p.nnz = g.nnz;
    p.max_iter = max_iter;
    p.ms_scaling = ms_scaling;
    p.cutoff = cutoff;
    p.sigma = sigma;
    p.order0 = d_order0;
    p.order_len = order_len;
    auto up = [](size_t x, size_t q) { return (x + q - 1) / q * q; };
    size_t off = (size_t) g.n * 8;
    p.off_soft = (uint32_t) off;
    off += (size_t) g.m * 8;
    p.off_dec = (uint32_t) off;
    off += up((size_t) g.n, 8);
    p.off_syn = (uint32_t) off;
    off += up((size_t) g.m, 16);
    p.off_msg = (uint32_t) off;
    const size_t fixed = off;
    if (fixed > (size_t) max_smem_optin) return -1;
    const bool msg_global = fixed + (size_t) g.nnz * 8 > (size_t) max_smem_optin / 4;
    const size_t smem = fixed + (msg_global ? 0 : (size_t) g.nnz * 8);
    using K = void (*)(const SoftParams);
    K k = msg_global ? (K) bp_softinfo_kernel<true> : (K) bp_softinfo_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 32, smem);
    if (e != cudaSuccess) return (int) e;
    if (occ < 1) return -1;
    int64_t grid = std::min<int64_t>((int64_t) occ * sm_count, batch);
    if (msg_global) {
        grid = std::min<int64_t>(grid, std::max<int64_t>(1, ((int64_t) 100 << 20) / ((int64_t) g.nnz * 8)));
        const size_t need = (size_t) grid * (size_t) g.nnz * 8;
        if (scratch->bytes < need) {
            if (scratch->ptr) cudaFree(scratch->ptr);
            scratch->ptr = nullptr;
            scratch->bytes = 0;
            e = cudaMalloc(&scratch->ptr, need);
            if (e != cudaSuccess) return (int) e;
            scratch->bytes = need;
        }
        p.msg_global = (double *) scratch->ptr;
    }
    if (grid < 1) grid = 1;
    p.soft_in = d_soft;
    p.batch = batch;
    p.counter = d_counter;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.out_soft = d_soft_out;
    k<<<(int) grid, 32, smem, st>>>(p);
    return (int) cudaGetLastError();
}

// Host side.  Returns a cudaError_t value (0 = ok), -1 when the code does not fit this kernel.
int launch_relative_kernel(const HostGraph &g, int sm_count, int max_smem_optin, const uint32_t *d_blob,
                           uint32_t prior_off, int method, int max_iter, double ms_scaling, const uint32_t *d_order0,
                           int order_len, const uint32_t *d_packed, int mwp, int64_t batch,
                           unsigned long long *d_counter, DeviceBuffer *scratch, uint8_t *d_dec, uint8_t *d_conv,
                           int32_t *d_iters, double *d_llr, int llr_last_only, int32_t *d_order_out, cudaStream_t st,
                           int *grid_out) {
    if (g.max_col_degree > 32) return -1;  // a lane per check of the bit
    RelParams p{};
    p.row_ptr = d_blob;
    p.col_idx = p.row_ptr + (g.m + 1);
    p.col_ptr = p.col_idx + g.nnz;
    p.csc2csr = p.col_ptr + (g.n + 1);
    p.row_idx = p.csc2csr + g.nnz;
    p.prior = reinterpret_cast<const double *>(d_blob + prior_off);
    p.m = g.m;
    p.n = g.n;
    p.nnz = g.nnz;
    p.MW = (g.m + 31) / 32;
    p.max_iter = max_iter;
    p.method = method;
    p.ms_scaling = ms_scaling;
    p.order0 = d_order0;
    p.order_len = order_len;
    auto up = [](size_t x, size_t q) { return (x + q - 1) / q * q; };
    size_t off = (size_t) g.n * 8;
    p.off_order = (uint32_t) off;
    off += up((size_t) std::max(order_len, 1) * 4, 8);
    p.off_dec = (uint32_t) off;
    off += up((size_t) g.n, 8);
    p.off_syn = (uint32_t) off;
    off += up((size_t) p.MW * 4, 8);
    p.off_stack = (uint32_t) off;
    off += up((size_t) stlsort::kStackInts * 4, 16);
    p.off_msg = (uint32_t) off;
    const size_t fixed = off;
    if (fixed > (size_t) max_smem_optin) return -1;
    // messages on chip only if at least four warps still fit on an SM
    const bool msg_global = fixed + (size_t) g.nnz * 8 > (size_t) max_smem_optin / 4;
    const size_t smem = fixed + (msg_global ? 0 : (size_t) g.nnz * 8);
    using K = void (*)(const RelParams);
    K k = method == BPB_MINIMUM_SUM ? (msg_global ? (K) bp_relative_kernel<kMinimumSum, true>
                                                  : (K) bp_relative_kernel<kMinimumSum, false>)
                                    : (msg_global ? (K) bp_relative_kernel<kProductSum, true>
                                                  : (K) bp_relative_kernel<kProductSum, false>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 32, smem);
    if (e != cudaSuccess) return (int) e;
    if (occ < 1) return -1;
    int64_t grid = std::min<int64_t>((int64_t) occ * sm_count, batch);
    if (msg_global) {
        // keep the message scratch of the resident warps inside the L2 (about 100 MB of its 126 MB)
        grid = std::min<int64_t>(grid, std::max<int64_t>(1, ((int64_t) 100 << 20) / ((int64_t) g.nnz * 8)));
        const size_t need = (size_t) grid * (size_t) g.nnz * 8;
        if (scratch->bytes < need) {
            if (scratch->ptr) cudaFree(scratch->ptr);
            scratch->ptr = nullptr;
            scratch->bytes = 0;
            e = cudaMalloc(&scratch->ptr, need);
            if (e != cudaSuccess) return (int) e;
            scratch->bytes = need;
        }
        p.msg_global = (double *) scratch->ptr;
    }
    if (grid < 1) grid = 1;
    p.synd_packed = d_packed;
    p.mwp = mwp;
    p.batch = batch;
    p.counter = d_counter;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.llr_last_only = llr_last_only;
    p.out_order = d_order_out;
    k<<<(int) grid, 32, smem, st>>>(p);
    if (grid_out) *grid_out = (int) grid;
    return (int) cudaGetLastError();
}

}  // namespace bpb
