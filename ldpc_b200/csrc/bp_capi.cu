// bp_capi.cu -- C-ABI implementation (include/bp_b200.h) and kernel dispatch.
//
// Host-side responsibilities that the reference spreads over the Cython shim and the BpDecoder
// constructor (src_python/ldpc/bp_decoder/_bp_decoder.pyx:9-49,88-160; src_cpp/bp.hpp:77-157):
// flatten H into sorted CSR/CSC, compute the channel priors log((1-p)/p) with the host libm (so they
// are the very doubles the reference computes, bp.hpp:150-151), stage everything on the device, and
// drive the batched kernels.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "bp_decoder.h"
#include "bp_smem_params.h"
#include "bp_stream_params.h"

namespace {

std::string g_create_err;
std::mutex g_create_mu;

#define BPB_CUDA(h, call)                                                                               \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                              \
            return BPB_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

int ensure(bpb_decoder *h, bpb::DeviceBuffer &b, size_t bytes, bool zero = false, cudaStream_t zs = nullptr) {
    if (bytes == 0) bytes = 16;
    if (b.bytes >= bytes) return BPB_OK;
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
    cudaError_t e = cudaMalloc(&b.ptr, bytes);
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        return BPB_ERR_NOMEM;
    }
    b.bytes = bytes;
    if (zero) BPB_CUDA(h, cudaMemsetAsync(b.ptr, 0, bytes, zs));
    return BPB_OK;
}

std::vector<bpb::DeviceBuffer *> all_buffers(bpb_decoder *h) {
    return {&h->blob,     &h->order_d,   &h->counter,   &h->msg,        &h->dec_w,      &h->syn_w,    &h->llr_tile,
            &h->packed,   &h->smem_tab,  &h->handoff,   &h->osd_llr,    &h->osd_fail_llr, &h->osd_fail_idx, &h->osd_count,  &h->st_in[0],  &h->st_in[1],   &h->st_dec[0],  &h->st_dec[1],
            &h->st_conv[0], &h->st_conv[1], &h->st_iters[0], &h->st_iters[1], &h->st_llr[0], &h->st_llr[1]};
}

void release(bpb::DeviceBuffer &b) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
}

inline int round_up(int x, int q) { return (x + q - 1) / q * q; }

// ---- small helper kernels ---------------------------------------------------------------------------

// [B][m] uint8 (0/1) -> [B][mwp] packed words; one warp per syndrome row, lane = bit position.
__global__ void pack_syndromes_kernel(const uint8_t *__restrict__ in, long long batch, int m, int mwp,
                                      uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < batch; b += nwarps) {
        const uint8_t *row = in + b * m;
        for (int w = 0; w < mwp; ++w) {
            const int i = w * 32 + lane;
            const uint32_t bit = (i < m) ? (row[i] != 0) : 0u;
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) out[b * mwp + w] = word;
        }
    }
}

// received vectors [B][n] -> packed syndromes s = H v (reference GF2Sparse::mulvec, gf2sparse.hpp:177-196,
// used by BpDecoder::decode for received-vector input, bp.hpp:162-165).  One warp per vector.
__global__ void pack_received_kernel(const uint8_t *__restrict__ in, long long batch, int m, int n, int mwp,
                                     const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col_idx,
                                     uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < batch; b += nwarps) {
        const uint8_t *v = in + b * n;
        for (int w = 0; w < mwp; ++w) {
            const int i = w * 32 + lane;
            uint32_t bit = 0;
            if (i < m)
                for (uint32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) bit ^= (v[col_idx[e]] != 0);
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) out[b * mwp + w] = word;
        }
    }
}

// decoding ^= received vector (bp.hpp:174-176)
__global__ void xor_received_kernel(uint8_t *__restrict__ dec, const uint8_t *__restrict__ v, long long count) {
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dec[i] ^= (v[i] != 0);
}

using bpb::build_serial_batches;
using bpb::build_smem_plan;
using bpb::compute_priors;

// BP+OSD: gather the posterior LLR rows of the syndromes BP did not solve (one warp per syndrome) so that only
// those cross PCIe (the reference hands bpd.log_prob_ratios to OSD only when !bpd.converge, _bposd_decoder.pyx:128-134)
__global__ void compact_failures_kernel(const uint8_t *__restrict__ conv, const double *__restrict__ llr,
                                        long long batch, int n, unsigned long long *count,
                                        uint32_t *__restrict__ fail_idx, double *__restrict__ fail_llr) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < batch; b += nwarps) {
        if (conv[b]) continue;
        unsigned long long slot = 0;
        if (lane == 0) slot = atomicAdd(count, 1ull);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (lane == 0) fail_idx[slot] = (uint32_t) b;
        for (int j = lane; j < n; j += 32) fail_llr[slot * n + j] = llr[b * n + j];
    }
}

// ---- graph blob ---------------------------------------------------------------------------------------

int upload_graph(bpb_decoder *h) {
    const bpb::HostGraph &g = h->g;
    compute_priors(h);
    size_t words = (size_t) (g.m + 1) + (size_t) g.nnz + (size_t) (g.n + 1) + (size_t) g.nnz + (size_t) g.nnz;
    if (words & 1) words++;
    h->prior_off = (uint32_t) words;
    words += 2 * (size_t) g.n;
    h->blob_words = (uint32_t) words;
    std::vector<uint32_t> blob(words, 0u);
    uint32_t *p = blob.data();
    std::memcpy(p, g.row_ptr.data(), sizeof(uint32_t) * (size_t) (g.m + 1));
    p += g.m + 1;
    std::memcpy(p, g.col_idx.data(), sizeof(uint32_t) * (size_t) g.nnz);
    p += g.nnz;
    std::memcpy(p, g.col_ptr.data(), sizeof(uint32_t) * (size_t) (g.n + 1));
    p += g.n + 1;
    std::memcpy(p, g.csc2csr.data(), sizeof(uint32_t) * (size_t) g.nnz);
    p += g.nnz;
    std::memcpy(p, g.row_idx.data(), sizeof(uint32_t) * (size_t) g.nnz);
    std::memcpy(blob.data() + h->prior_off, h->prior.data(), sizeof(double) * (size_t) g.n);
    int rc = ensure(h, h->blob, words * sizeof(uint32_t));
    if (rc) return rc;
    BPB_CUDA(h, cudaMemcpyAsync(h->blob.ptr, blob.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                h->stream));
    if (h->serial_order.empty()) {
        h->serial_order.resize((size_t) g.n);
        for (int j = 0; j < g.n; j++) h->serial_order[(size_t) j] = (uint32_t) j;
    }
    const int sb = bpb::serial_batch(g.max_row_degree, g.max_col_degree, g.regular);
    h->serial_batches = build_serial_batches(g, h->serial_order, sb);
    std::vector<uint32_t> upload = h->serial_batches;
    h->serial_entries = (int) h->serial_batches.size();
    if (g.regular && g.max_row_degree == 6 && g.max_col_degree == 3 && g.m < (1 << 28)) {
        // regular-code serial program (see bp_stream.cuh): {j, row | self << 28 for each of the three edges}
        upload.assign(h->serial_batches.size() * 4, 0xffffffffu);
        for (size_t q = 0; q < h->serial_batches.size(); q++) {
            const uint32_t j = h->serial_batches[q];
            if (j == 0xffffffffu) continue;
            upload[4 * q] = j;
            for (uint32_t e = g.col_ptr[j]; e < g.col_ptr[j + 1]; e++) {
                const uint32_t i = g.row_idx[e], self = g.csc2csr[e] - g.row_ptr[i];
                upload[4 * q + 1 + (e - g.col_ptr[j])] = i | (self << 28);
            }
        }
    }
    rc = ensure(h, h->order_d, upload.size() * sizeof(uint32_t));
    if (rc) return rc;
    BPB_CUDA(h, cudaMemcpy(h->order_d.ptr, upload.data(), upload.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    build_smem_plan(h);
    if (h->smem_plan.ok) {
        rc = ensure(h, h->smem_tab, h->smem_plan.blob.size());
        if (rc) return rc;
        BPB_CUDA(h, cudaMemcpyAsync(h->smem_tab.ptr, h->smem_plan.blob.data(), h->smem_plan.blob.size(),
                                    cudaMemcpyHostToDevice, h->stream));
    }
    BPB_CUDA(h, cudaStreamSynchronize(h->stream));  // the host vector `blob` dies here
    h->graph_dirty = false;
    return BPB_OK;
}

// ---- stream-family dispatch -------------------------------------------------------------------------

using bpb::StreamKernel;

StreamKernel pick_stream(int method, int schedule, int dc, int dv, bool reg, bool llr) {
    if (method == BPB_MINIMUM_SUM && schedule == BPB_PARALLEL) return bpb::pick_stream_ms_parallel(dc, dv, reg, llr);
    if (method == BPB_PRODUCT_SUM && schedule == BPB_PARALLEL) return bpb::pick_stream_ps_parallel(dc, dv, reg, llr);
    if (method == BPB_MINIMUM_SUM && schedule == BPB_SERIAL) return bpb::pick_stream_ms_serial(dc, dv, reg, llr);
    if (method == BPB_PRODUCT_SUM && schedule == BPB_SERIAL) return bpb::pick_stream_ps_serial(dc, dv, reg, llr);
    return nullptr;
}

int launch_smem(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
                int32_t *d_iters, double *d_llr, cudaStream_t st, const uint32_t *index_list,
                const unsigned long long *batch_dev);

int launch_stream(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
                  int32_t *d_iters, double *d_llr, cudaStream_t st) {
    const bpb::HostGraph &g = h->g;
    const bool llr = d_llr != nullptr;
    StreamKernel k = pick_stream(h->method, h->schedule, g.max_row_degree, g.max_col_degree, g.regular, llr);
    if (!k) {
        h->err = "no streaming kernel for this configuration";
        return BPB_ERR_UNSUPPORTED;
    }
    const bool generic = g.max_row_degree > 32 || g.max_col_degree > 16;  // two message arrays per tile
    const int block = 256, wpb = block / 32;
    const int m_pad = round_up(g.m, 32), n_pad = round_up(g.n, 32);
    const size_t blob_bytes = (size_t) h->blob_words * 4;
    bpb::StreamParams p{};
    p.smem_graph = blob_bytes <= 96 * 1024;
    const size_t syn_bytes = (size_t) wpb * m_pad * 4;
    const size_t base_bytes = p.smem_graph ? blob_bytes : 0;
    p.smem_syn = (base_bytes + syn_bytes) <= 112 * 1024;
    p.smem_syn_off = (uint32_t) ((base_bytes + 15) / 16 * 4);
    const size_t smem_bytes = (size_t) p.smem_syn_off * 4 + (p.smem_syn ? syn_bytes : 0);
    BPB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes));
    int occ = 0;
    BPB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, block, smem_bytes));
    if (occ < 1) {
        h->err = "stream kernel does not fit on an SM";
        return BPB_ERR_UNSUPPORTED;
    }
    int64_t grid64 = (int64_t) occ * h->sm_count;
    const int64_t need = (batch + block - 1) / block;
    if (grid64 > need) grid64 = need;
    if (grid64 < 1) grid64 = 1;
    const int grid = (int) grid64;
    const size_t warps = (size_t) grid * wpb;
    int rc;
    if ((rc = ensure(h, h->msg, warps * (size_t) g.nnz * 32 * sizeof(double) * (generic ? 2 : 1)))) return rc;
    if ((rc = ensure(h, h->dec_w, warps * (size_t) n_pad * 4, true, st))) return rc;
    if ((rc = ensure(h, h->syn_w, p.smem_syn ? 16 : warps * (size_t) m_pad * 4, true, st))) return rc;
    if (llr && (rc = ensure(h, h->llr_tile, warps * (size_t) g.n * 32 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->counter, 64))) return rc;
    BPB_CUDA(h, cudaMemsetAsync(h->counter.ptr, 0, 64, st));
    // second stage for the ramp-down (parallel schedule, when the thread-group kernels can take the code)
    const bool second_stage = h->smem_plan.ok && h->max_iter > 16;
    if (second_stage && (rc = ensure(h, h->handoff, (size_t) warps * 32 * sizeof(uint32_t)))) return rc;
    p.iter_cap = second_stage ? 12 : h->max_iter + 1;
    p.handoff_count = (unsigned long long *) h->counter.ptr + 1;
    p.handoff_list = (uint32_t *) h->handoff.ptr;
    p.iter_total = (unsigned long long *) h->counter.ptr + 3;

    p.blob = (const uint32_t *) h->blob.ptr;
    p.blob_words = h->blob_words;
    p.prior_off = h->prior_off;
    p.m = g.m;
    p.n = g.n;
    p.nnz = g.nnz;
    p.mwp = mwp;
    p.m_pad = m_pad;
    p.n_pad = n_pad;
    p.max_iter = h->max_iter;
    p.ms_scaling = h->ms_scaling;
    p.uniform_prior = h->uniform_prior ? 1 : 0;
    p.prior0 = h->prior.empty() ? 0.0 : h->prior[0];
    p.synd_packed = d_packed;
    p.batch = batch;
    p.counter = (unsigned long long *) h->counter.ptr;
    p.msg = (double *) h->msg.ptr;
    p.dec_w = (uint32_t *) h->dec_w.ptr;
    p.syn_w_g = (uint32_t *) h->syn_w.ptr;
    p.llr_tile = (double *) h->llr_tile.ptr;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.order = (const uint32_t *) h->order_d.ptr;
    p.order_len = h->serial_entries;
    BPB_CUDA(h, cudaEventRecord(h->kev0, st));
    k<<<grid, block, smem_bytes, st>>>(p);
    BPB_CUDA(h, cudaGetLastError());
    BPB_CUDA(h, cudaEventRecord(h->kev1, st));
    h->kernel_timed = true;
    h->launches += 1;
    h->last_family = BPB_KERNEL_STREAM;
    h->last_grid = grid;
    h->last_block = block;
    if (second_stage) {
        rc = launch_smem(h, d_packed, mwp, batch, d_dec, d_conv, d_iters, d_llr, st, (const uint32_t *) h->handoff.ptr,
                         (const unsigned long long *) h->counter.ptr + 1);
        if (rc) return rc;
        h->last_family = BPB_KERNEL_STREAM;
        h->last_grid = grid;
        h->last_block = block;
    }
    return BPB_OK;
}

// ---- on-chip family: launch -------------------------------------------------------------------------------

inline uint32_t align_up(uint32_t x, uint32_t q) { return (x + q - 1) / q * q; }

bpb::SmemKernel pick_smem(int method, int schedule, int dc, int dv, bool regular, bool llr) {
    if (schedule == BPB_SERIAL)
        return method == BPB_MINIMUM_SUM ? bpb::pick_smem_serial_ms(dc, dv, regular, llr)
                                         : bpb::pick_smem_serial_ps(dc, dv, regular, llr);
    if (method == BPB_MINIMUM_SUM) return bpb::pick_smem_ms(dc, dv, regular, llr);
    return bpb::pick_smem_ps(dc, dv, regular, llr);
}

int launch_smem(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
                int32_t *d_iters, double *d_llr, cudaStream_t st, const uint32_t *index_list,
                const unsigned long long *batch_dev) {
    const bpb::HostGraph &g = h->g;
    const bpb::SmemPlan &pl = h->smem_plan;
    const bool llr = d_llr != nullptr;
    bpb::SmemKernel k = pick_smem(h->method, h->schedule, g.max_row_degree, g.max_col_degree, g.regular, llr);
    if (!k || !pl.ok || pl.serial != (h->schedule == BPB_SERIAL)) {
        h->err = "on-chip kernel family not available for this code: " + pl.why;
        return BPB_ERR_UNSUPPORTED;
    }
    const bool serial = h->schedule == BPB_SERIAL;
    const int maxt = serial ? 512 : bpb::smem_cta_threads(h->method, g.max_row_degree, g.max_col_degree);
    const size_t tab = pl.blob.size();
    int G = (int) (((size_t) h->max_smem_optin - tab) / pl.group_bytes);
    G = std::min(G, 15);
    // threads per group: enough to cover the rows, a multiple of 32, within the CTA budget
    int T = maxt / G / 32 * 32;
    if (T < 32) {
        T = 32;
        G = maxt / 32;
    }
    // parallel: a thread per row / two columns; serial: a thread per bit of a level
    const int want = serial ? std::max(32, (int) align_up((uint32_t) std::max(pl.mean_level, 1), 32))
                            : std::max(32, (int) align_up((uint32_t) std::max(g.m, (g.n + 1) / 2), 32));
    T = std::min(T, std::min(want, serial ? 128 : 256));
    if (const char *ov = std::getenv("BPB_SMEM_GROUP_THREADS")) {  // tuning override: threads per group
        const int t_ov = std::atoi(ov);
        if (t_ov >= 32 && t_ov % 32 == 0 && t_ov <= maxt) {
            T = t_ov;
            G = std::min(G, maxt / T);
        }
    }
    const int block = G * T;
    const size_t smem_bytes = tab + (size_t) G * pl.group_bytes;
    BPB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes));
    int64_t grid64 = std::min<int64_t>(h->sm_count, (batch + G - 1) / G);
    if (grid64 < 1) grid64 = 1;
    int rc;
    if ((rc = ensure(h, h->counter, 64))) return rc;
    // counter words: [0] streaming queue, [1] hand-off count, [2] thread-group queue
    if (!index_list) BPB_CUDA(h, cudaMemsetAsync(h->counter.ptr, 0, 64, st));
    if (index_list) grid64 = h->sm_count;  // the count lives on the device
    bpb::SmemParams p{};
    p.tab = (const uint32_t *) h->smem_tab.ptr;
    p.tab_bytes = (uint32_t) tab;
    p.off_row_deg = pl.off_row_deg;
    p.off_col_deg = pl.off_col_deg;
    p.off_col_row = pl.off_col_row;
    p.off_row_pos = pl.off_row_pos;
    p.off_col_pos = pl.off_col_pos;
    p.off_prior = pl.off_prior;
    p.off_col_self = pl.off_col_self;
    p.off_lev_ptr = pl.off_lev_ptr;
    p.off_lev_bits = pl.off_lev_bits;
    p.n_levels = pl.n_levels;
    p.group_bytes = pl.group_bytes;
    p.goff_msg = pl.goff_msg;
    p.goff_dec = pl.goff_dec;
    p.goff_syn = pl.goff_syn;
    p.goff_ctl = pl.goff_ctl;
    p.m = g.m;
    p.n = g.n;
    p.M = pl.M;
    p.N = pl.N;
    p.MW = (g.m + 31) / 32;
    p.groups = G;
    p.T = T;
    p.max_iter = h->max_iter;
    p.ms_scaling = h->ms_scaling;
    p.uniform_prior = h->uniform_prior ? 1 : 0;
    p.prior0 = h->prior.empty() ? 0.0 : h->prior[0];
    p.synd_packed = d_packed;
    p.mwp = mwp;
    p.batch = batch;
    p.counter = (unsigned long long *) h->counter.ptr + 2;
    p.index_list = index_list;
    p.batch_dev = batch_dev;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    if (!index_list) BPB_CUDA(h, cudaEventRecord(h->kev0, st));
    k<<<(int) grid64, block, smem_bytes, st>>>(p);
    BPB_CUDA(h, cudaGetLastError());
    if (!index_list) BPB_CUDA(h, cudaEventRecord(h->kev1, st));
    h->kernel_timed = true;
    h->launches += 1;
    h->last_family = BPB_KERNEL_SMEM;
    h->last_grid = (int) grid64;
    h->last_block = block;
    return BPB_OK;
}

int check_ready(bpb_decoder *h) {
    if (!h) return BPB_ERR_ARG;
    if (h->channel.empty()) {
        h->err = "channel probabilities not set";
        return BPB_ERR_ARG;
    }
    if (h->max_iter < 1) {
        h->err = "maximum_iterations must be >= 1";
        return BPB_ERR_ARG;
    }
    if (h->device < 0) {
        h->err = "host-only handle (device < 0): decoding needs a CUDA device, there is no CPU fallback";
        return BPB_ERR_CUDA;
    }
    return BPB_OK;
}

}  // namespace

// ======================================================================================================
extern "C" {

const char *bpb_version(void) { return "ldpc_b200 0.1 (sm_100a)"; }

const char *bpb_last_error(const bpb_decoder *h) {
    if (h) return h->err.c_str();
    std::lock_guard<std::mutex> lk(g_create_mu);
    return g_create_err.c_str();
}

int bpb_create(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, int device, bpb_decoder **out) {
    auto fail = [&](const std::string &msg, int code) {
        std::lock_guard<std::mutex> lk(g_create_mu);
        g_create_err = msg;
        return code;
    };
    if (!out) return fail("out is NULL", BPB_ERR_ARG);
    *out = nullptr;
    if (m < 1 || n < 1 || nnz < 0 || (nnz > 0 && (!rows || !cols))) return fail("bad matrix arguments", BPB_ERR_ARG);
    bpb_decoder *h = new (std::nothrow) bpb_decoder();
    if (!h) return fail("out of host memory", BPB_ERR_NOMEM);
    std::string err;
    int rc = bpb::build_host_graph(m, n, nnz, rows, cols, h->g, err);
    if (rc) {
        delete h;
        return fail(err, rc);
    }
    h->max_iter = n;  // maximum_iterations 0 means n in the shim (_bp_decoder.pyx:357)
    if (device < 0) {
        // host-only handle: only bpb_osd0_host (host-side by design) works; every decode call fails loudly
        h->device = -1;
        *out = h;
        return BPB_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1) {
        delete h;
        return fail(std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                        " (ldpc_b200 has no CPU fallback)",
                    BPB_ERR_CUDA);
    }
    if (device < 0 || device >= count) {
        delete h;
        return fail("device ordinal out of range", BPB_ERR_ARG);
    }
    h->device = device;
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&h->kev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->kev1);
    if (e != cudaSuccess) {
        std::string msg = std::string("CUDA init failed: ") + cudaGetErrorString(e);
        bpb_destroy(h);
        return fail(msg, BPB_ERR_CUDA);
    }
    *out = h;
    return BPB_OK;
}

void bpb_destroy(bpb_decoder *h) {
    if (!h) return;
    if (h->device < 0) {
        delete h;
        return;
    }
    cudaSetDevice(h->device);
    for (bpb::DeviceBuffer *b: all_buffers(h)) release(*b);
    for (int i = 0; i < 2; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->kev0) cudaEventDestroy(h->kev0);
    if (h->kev1) cudaEventDestroy(h->kev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int bpb_set_channel(bpb_decoder *h, const double *p, int n) {
    if (!h) return BPB_ERR_ARG;
    if (!p || n != h->g.n) {
        h->err = "Channel probabilities vector must have length equal to the number of bits";  // bp.hpp:106-109
        return BPB_ERR_ARG;
    }
    h->channel.assign(p, p + n);
    h->graph_dirty = true;
    return BPB_OK;
}

int bpb_set_max_iter(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v < 1) {
        h->err = "maximum_iterations must be >= 1";
        return BPB_ERR_ARG;
    }
    h->max_iter = v;
    return BPB_OK;
}

int bpb_set_method(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_PRODUCT_SUM && v != BPB_MINIMUM_SUM) {
        h->err = "invalid bp_method";
        return BPB_ERR_ARG;
    }
    h->method = v;
    return BPB_OK;
}

int bpb_set_schedule(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_SERIAL && v != BPB_PARALLEL) {
        h->err = "Invalid BP schedule";  // bp.hpp:171
        return v == 2 ? BPB_ERR_UNSUPPORTED : BPB_ERR_ARG;
    }
    if (h->schedule != v) h->graph_dirty = true;  // the shared-memory plan depends on the schedule
    h->schedule = v;
    return BPB_OK;
}

int bpb_set_ms_scaling_factor(bpb_decoder *h, double v) {
    if (!h) return BPB_ERR_ARG;
    h->ms_scaling = v;
    return BPB_OK;
}

int bpb_set_serial_schedule_order(bpb_decoder *h, const int32_t *order, int len) {
    if (!h) return BPB_ERR_ARG;
    if (!order) {
        h->serial_order.clear();
        h->graph_dirty = true;
        return BPB_OK;
    }
    if (len < 0) return BPB_ERR_ARG;
    for (int i = 0; i < len; i++)
        if (order[i] < 0 || order[i] >= h->g.n) {
            h->err = "serial_schedule_order entry out of range";
            return BPB_ERR_ARG;
        }
    h->serial_order.assign(order, order + len);
    h->graph_dirty = true;
    return BPB_OK;
}

int bpb_set_kernel(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_KERNEL_AUTO && v != BPB_KERNEL_STREAM && v != BPB_KERNEL_SMEM) {
        h->err = "invalid kernel family";
        return BPB_ERR_ARG;
    }
    h->kernel_pref = v;
    return BPB_OK;
}

int bpb_decode_batch_device(bpb_decoder *h, int input_type, const uint8_t *d_input, int64_t batch, uint8_t *d_decoding,
                            uint8_t *d_converged, int32_t *d_iterations, double *d_llr, void *cuda_stream) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!d_input || !d_decoding))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (input_type != BPB_INPUT_SYNDROME && input_type != BPB_INPUT_RECEIVED_VECTOR) {
        h->err = "invalid input type";
        return BPB_ERR_ARG;
    }
    if (batch == 0) return BPB_OK;
    BPB_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t) cuda_stream;
    if (h->graph_dirty) {
        // the blob upload runs on the handle's own stream and is synchronised inside
        if ((rc = upload_graph(h))) return rc;
    }
    const bpb::HostGraph &g = h->g;
    const int mwp = round_up((g.m + 31) / 32, 4);
    if ((rc = ensure(h, h->packed, (size_t) batch * mwp * 4))) return rc;
    uint32_t *d_packed = (uint32_t *) h->packed.ptr;
    const int pgrid = (int) std::min<int64_t>((batch + 7) / 8, (int64_t) h->sm_count * 16);
    if (input_type == BPB_INPUT_SYNDROME) {
        pack_syndromes_kernel<<<pgrid, 256, 0, st>>>(d_input, batch, g.m, mwp, d_packed);
    } else {
        const uint32_t *blob = (const uint32_t *) h->blob.ptr;
        pack_received_kernel<<<pgrid, 256, 0, st>>>(d_input, batch, g.m, g.n, mwp, blob, blob + (g.m + 1), d_packed);
    }
    BPB_CUDA(h, cudaGetLastError());
    h->launches += 1;
    // family: the on-chip kernels serve the parallel schedule of codes whose messages fit in shared memory;
    // everything else (serial schedule, large codes) streams its messages through HBM.
    const bool smem_able = h->smem_plan.ok;
    if (h->kernel_pref == BPB_KERNEL_SMEM && !smem_able) {
        h->err = "kernel family 'smem' requested but not available: " + h->smem_plan.why;
        return BPB_ERR_UNSUPPORTED;
    }
    // AUTO: the thread-group kernels win for the parallel schedule (3x the streaming family at n = 1000); for the
    // serial schedule the levelised streaming kernel measured slightly faster (6.5 vs 5.8 M decodes/s at n = 1000,
    // a level has only ~n/30 independent bits), so the on-chip serial kernel serves as its ramp-down second stage
    // and on request (kernel = smem).
    const bool use_smem = smem_able && (h->kernel_pref == BPB_KERNEL_SMEM ||
                                        (h->kernel_pref == BPB_KERNEL_AUTO && h->schedule == BPB_PARALLEL));
    if (use_smem)
        rc = launch_smem(h, d_packed, mwp, batch, d_decoding, d_converged, d_iterations, d_llr, st, nullptr, nullptr);
    else
        rc = launch_stream(h, d_packed, mwp, batch, d_decoding, d_converged, d_iterations, d_llr, st);
    if (rc) return rc;
    if (input_type == BPB_INPUT_RECEIVED_VECTOR) {
        const long long count = (long long) batch * g.n;
        const int xgrid = (int) std::min<long long>((count + 255) / 256, (long long) h->sm_count * 32);
        xor_received_kernel<<<xgrid, 256, 0, st>>>(d_decoding, d_input, count);
        BPB_CUDA(h, cudaGetLastError());
        h->launches += 1;
    }
    return BPB_OK;
}

int bpb_decode_batch(bpb_decoder *h, int input_type, const uint8_t *input, int64_t batch, uint8_t *decoding,
                     uint8_t *converged, int32_t *iterations, double *llr) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!input || !decoding))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (batch == 0) return BPB_OK;
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    const bpb::HostGraph &g = h->g;
    const int in_w = (input_type == BPB_INPUT_RECEIVED_VECTOR) ? g.n : g.m;
    // Chunked three-stage pipeline (H2D | kernels | D2H) over two staging slots.  The on-chip family keeps no
    // per-lane state in HBM, so small chunks cost nothing; the streaming family amortises its persistent-lane
    // ramp-down over large chunks.
    const bool smem_able = h->smem_plan.ok && (h->kernel_pref == BPB_KERNEL_SMEM ||
                                               (h->kernel_pref == BPB_KERNEL_AUTO && h->schedule == BPB_PARALLEL));
    const int64_t chunk_max = smem_able ? ((int64_t) 1 << 17) : ((int64_t) 1 << 20);
    int64_t c = 0;
    for (int64_t lo = 0; lo < batch; lo += chunk_max, ++c) {
        const int s = (int) (c & 1);
        const int64_t nb = std::min(chunk_max, batch - lo);
        const size_t cap = (size_t) std::min(chunk_max, batch);
        if ((rc = ensure(h, h->st_in[s], cap * in_w))) return rc;
        if ((rc = ensure(h, h->st_dec[s], cap * g.n))) return rc;
        if ((rc = ensure(h, h->st_conv[s], cap))) return rc;
        if ((rc = ensure(h, h->st_iters[s], cap * 4))) return rc;
        if (llr && (rc = ensure(h, h->st_llr[s], cap * g.n * 8))) return rc;
        if (c >= 2) BPB_CUDA(h, cudaStreamWaitEvent(h->s_in, h->ev_k[s], 0));  // slot's input consumed
        BPB_CUDA(h, cudaMemcpyAsync(h->st_in[s].ptr, input + lo * in_w, (size_t) nb * in_w, cudaMemcpyHostToDevice,
                                    h->s_in));
        BPB_CUDA(h, cudaEventRecord(h->ev_in[s], h->s_in));
        BPB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_in[s], 0));
        if (c >= 2) BPB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_out[s], 0));  // slot's outputs drained
        rc = bpb_decode_batch_device(h, input_type, (const uint8_t *) h->st_in[s].ptr, nb,
                                     (uint8_t *) h->st_dec[s].ptr, (uint8_t *) h->st_conv[s].ptr,
                                     (int32_t *) h->st_iters[s].ptr, llr ? (double *) h->st_llr[s].ptr : nullptr,
                                     h->stream);
        if (rc) return rc;
        BPB_CUDA(h, cudaEventRecord(h->ev_k[s], h->stream));
        BPB_CUDA(h, cudaStreamWaitEvent(h->s_out, h->ev_k[s], 0));
        BPB_CUDA(h, cudaMemcpyAsync(decoding + lo * g.n, h->st_dec[s].ptr, (size_t) nb * g.n, cudaMemcpyDeviceToHost,
                                    h->s_out));
        if (converged)
            BPB_CUDA(h, cudaMemcpyAsync(converged + lo, h->st_conv[s].ptr, (size_t) nb, cudaMemcpyDeviceToHost,
                                        h->s_out));
        if (iterations)
            BPB_CUDA(h, cudaMemcpyAsync(iterations + lo, h->st_iters[s].ptr, (size_t) nb * 4,
                                        cudaMemcpyDeviceToHost, h->s_out));
        if (llr)
            BPB_CUDA(h, cudaMemcpyAsync(llr + lo * g.n, h->st_llr[s].ptr, (size_t) nb * g.n * 8,
                                        cudaMemcpyDeviceToHost, h->s_out));
        BPB_CUDA(h, cudaEventRecord(h->ev_out[s], h->s_out));
    }
    BPB_CUDA(h, cudaStreamSynchronize(h->s_out));
    BPB_CUDA(h, cudaStreamSynchronize(h->stream));
    return BPB_OK;
}

int bpb_osd0_host(bpb_decoder *h, const uint8_t *syndromes, const double *llr, const uint8_t *converged,
                  int64_t batch, uint8_t *decoding, int threads) {
    if (!h) return BPB_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!syndromes || !llr || !decoding))) {
        h->err = "bad osd arguments";
        return BPB_ERR_ARG;
    }
    int rc = bpb::osd0_host(h->g, syndromes, llr, converged, batch, decoding, threads);
    if (rc) h->err = "osd0_host failed";
    return rc;
}

int bpb_bposd_decode_batch(bpb_decoder *h, const uint8_t *syndromes, int64_t batch, uint8_t *decoding,
                           uint8_t *converged, int32_t *iterations, uint8_t *bp_decoding, int threads) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!syndromes || !decoding))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (batch == 0) return BPB_OK;
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    const bpb::HostGraph &g = h->g;
    const int64_t chunk_max = (int64_t) 1 << 17;
    const size_t cap = (size_t) std::min(chunk_max, batch);
    if ((rc = ensure(h, h->osd_llr, cap * g.n * 8))) return rc;
    if ((rc = ensure(h, h->osd_fail_llr, cap * g.n * 8))) return rc;
    if ((rc = ensure(h, h->osd_fail_idx, cap * 4))) return rc;
    if ((rc = ensure(h, h->osd_count, 16))) return rc;
    std::vector<uint8_t> host_conv(cap);
    std::vector<uint32_t> fidx;
    std::vector<double> fllr;
    std::vector<uint8_t> fsyn, fdec;
    for (int64_t lo = 0; lo < batch; lo += chunk_max) {
        const int64_t nb = std::min(chunk_max, batch - lo);
        if ((rc = ensure(h, h->st_in[0], cap * g.m))) return rc;
        if ((rc = ensure(h, h->st_dec[0], cap * g.n))) return rc;
        if ((rc = ensure(h, h->st_conv[0], cap))) return rc;
        if ((rc = ensure(h, h->st_iters[0], cap * 4))) return rc;
        cudaStream_t st = h->stream;
        BPB_CUDA(h, cudaMemcpyAsync(h->st_in[0].ptr, syndromes + lo * g.m, (size_t) nb * g.m, cudaMemcpyHostToDevice, st));
        rc = bpb_decode_batch_device(h, BPB_INPUT_SYNDROME, (const uint8_t *) h->st_in[0].ptr, nb,
                                     (uint8_t *) h->st_dec[0].ptr, (uint8_t *) h->st_conv[0].ptr,
                                     (int32_t *) h->st_iters[0].ptr, (double *) h->osd_llr.ptr, st);
        if (rc) return rc;
        BPB_CUDA(h, cudaMemsetAsync(h->osd_count.ptr, 0, 16, st));
        const int cgrid = (int) std::min<int64_t>((nb + 7) / 8, (int64_t) h->sm_count * 16);
        compact_failures_kernel<<<cgrid, 256, 0, st>>>((const uint8_t *) h->st_conv[0].ptr, (const double *) h->osd_llr.ptr,
                                                      nb, g.n, (unsigned long long *) h->osd_count.ptr,
                                                      (uint32_t *) h->osd_fail_idx.ptr, (double *) h->osd_fail_llr.ptr);
        BPB_CUDA(h, cudaGetLastError());
        h->launches += 1;
        unsigned long long nfail = 0;
        BPB_CUDA(h, cudaMemcpyAsync(decoding + lo * g.n, h->st_dec[0].ptr, (size_t) nb * g.n, cudaMemcpyDeviceToHost, st));
        BPB_CUDA(h, cudaMemcpyAsync(host_conv.data(), h->st_conv[0].ptr, (size_t) nb, cudaMemcpyDeviceToHost, st));
        if (iterations)
            BPB_CUDA(h, cudaMemcpyAsync(iterations + lo, h->st_iters[0].ptr, (size_t) nb * 4, cudaMemcpyDeviceToHost, st));
        BPB_CUDA(h, cudaMemcpyAsync(&nfail, h->osd_count.ptr, 8, cudaMemcpyDeviceToHost, st));
        BPB_CUDA(h, cudaStreamSynchronize(st));
        if (converged) std::memcpy(converged + lo, host_conv.data(), (size_t) nb);
        if (bp_decoding) std::memcpy(bp_decoding + lo * g.n, decoding + lo * g.n, (size_t) nb * g.n);
        if (nfail) {
            fidx.resize(nfail);
            fllr.resize(nfail * (size_t) g.n);
            fsyn.resize(nfail * (size_t) g.m);
            fdec.assign(nfail * (size_t) g.n, 0);
            BPB_CUDA(h, cudaMemcpy(fidx.data(), h->osd_fail_idx.ptr, nfail * 4, cudaMemcpyDeviceToHost));
            BPB_CUDA(h, cudaMemcpy(fllr.data(), h->osd_fail_llr.ptr, nfail * (size_t) g.n * 8, cudaMemcpyDeviceToHost));
            for (size_t q = 0; q < nfail; q++)
                std::memcpy(&fsyn[q * g.m], syndromes + (lo + fidx[q]) * g.m, (size_t) g.m);
            rc = bpb::osd0_host(g, fsyn.data(), fllr.data(), nullptr, (int64_t) nfail, fdec.data(), threads);
            if (rc) {
                h->err = "osd0_host failed";
                return rc;
            }
            for (size_t q = 0; q < nfail; q++)
                std::memcpy(decoding + (lo + fidx[q]) * g.n, &fdec[q * g.n], (size_t) g.n);
        }
    }
    return BPB_OK;
}

int bpb_get_info(const bpb_decoder *h_, bpb_info *out) {
    bpb_decoder *h = const_cast<bpb_decoder *>(h_);
    if (!h || !out) return BPB_ERR_ARG;
    if (h->device < 0 && !h->channel.empty() && h->graph_dirty) {
        // host-only handle: the shared-memory plan is pure host work, build it so that it can be inspected
        if (h->max_smem_optin <= 0) h->max_smem_optin = 232448;
        compute_priors(h);
        build_smem_plan(h);
        h->graph_dirty = false;
    }
    std::memset(out, 0, sizeof(*out));
    out->m = h->g.m;
    out->n = h->g.n;
    out->nnz = h->g.nnz;
    out->max_row_degree = h->g.max_row_degree;
    out->max_col_degree = h->g.max_col_degree;
    out->device = h->device;
    out->sm_count = h->sm_count;
    out->kernel_family = h->last_family;
    out->grid = h->last_grid;
    out->block = h->last_block;
    out->launches = h->launches;
    int64_t ws = 0;
    for (bpb::DeviceBuffer *b: all_buffers(h)) ws += (int64_t) b->bytes;
    out->workspace_bytes = ws;
    if (h->device >= 0 && h->counter.ptr && h->last_family == BPB_KERNEL_STREAM) {
        unsigned long long words[4] = {0, 0, 0, 0};
        if (cudaMemcpy(words, h->counter.ptr, sizeof(words), cudaMemcpyDeviceToHost) == cudaSuccess) {
            out->stream_handed_off = (int64_t) words[1];
            out->stream_iterations = (int64_t) words[3];
        }
    }
    out->smem_family_available = h->smem_plan.ok ? 1 : 0;
    out->smem_bank_multiplicity = h->smem_plan.max_bank_multiplicity;
    out->smem_bytes_per_syndrome = (int) h->smem_plan.group_bytes;
    if (h->kernel_timed) {
        // CUDA-event time of the most recent message-update kernel (valid once that launch has finished)
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->kev0, h->kev1) == cudaSuccess) out->last_kernel_ms = ms;
    }
    return BPB_OK;
}

void *bpb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void bpb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
