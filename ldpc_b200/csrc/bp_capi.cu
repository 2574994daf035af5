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
#include <thread>

#include "bp_decoder.h"
#include "bp_edge_params.h"
#include "bp_pair_params.h"
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
    return {&h->blob,     &h->order_d,   &h->counter_s[0], &h->counter_s[1],   &h->msg_s[0], &h->msg_s[1],        &h->dec_w_s[0], &h->dec_w_s[1],      &h->syn_w_s[0], &h->syn_w_s[1],    &h->llr_tile_s[0], &h->llr_tile_s[1],
            &h->packed_s[0], &h->packed_s[1],   &h->smem_tab,  &h->handoff_s[0], &h->handoff_s[1],   &h->osd_llr,    &h->osd_fail_llr, &h->osd_fail_idx, &h->osd_count,  &h->st_in[0],  &h->st_in[1],   &h->st_dec[0],  &h->st_dec[1],
            &h->st_conv[0], &h->st_conv[1], &h->st_iters[0], &h->st_iters[1], &h->st_llr[0], &h->st_llr[1],
            &h->st_bp[0],   &h->st_bp[1],   &h->osd_conv,  &h->mc_thresh, &h->mc_err, &h->mc_syn, &h->mc_dec,
            &h->mc_conv,    &h->mc_its,     &h->mc_counts, &h->edge_msg_s[0], &h->edge_msg_s[1], &h->pair_tab, &h->rel_order, &h->rel_order_out, &h->rel_msg, &h->obs_tab,
            &h->b8_in[0], &h->b8_in[1], &h->b8_words[0], &h->b8_words[1], &h->b8_out[0], &h->b8_out[1],
            &h->b8_obs[0], &h->b8_obs[1], &h->soft_in, &h->soft_out, &h->soft_llr};
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

// ---- bit-packed I/O (stim's b8 layout: a row is ceil(bits / 8) bytes, bit k of the row is bit k % 8 of byte k / 8) ----

// b8 syndrome rows -> the kernels' packed words [B][mwp] (same bit order; the row stride changes and bits >= m are
// cleared: the convergence test looks at whole words)
__global__ void unpack_b8_rows_kernel(const uint8_t *__restrict__ in, long long batch, int m, int mb, int mwp,
                                      uint32_t *__restrict__ out) {
    const long long total = batch * mwp;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const long long b = t / mwp;
        const int w = (int) (t - b * mwp);
        const uint8_t *row = in + b * mb;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int byte = 4 * w + k;
            if (byte < mb) v |= (uint32_t) row[byte] << (8 * k);
        }
        const int first_bit = 32 * w;
        if (first_bit + 32 > m) v &= (first_bit >= m) ? 0u : ((1u << (m - first_bit)) - 1u);
        out[t] = v;
    }
}

// hard decisions [B][n] u8 -> b8 rows [B][ceil(n/8)]
__global__ void pack_b8_rows_kernel(const uint8_t *__restrict__ dec, long long batch, int n, int nb,
                                    uint8_t *__restrict__ out) {
    const long long total = batch * nb;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const long long b = t / nb;
        const int byte = (int) (t - b * nb);
        const uint8_t *row = dec + b * n + 8 * byte;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (8 * byte + k < n) v |= (uint32_t) (row[k] & 1u) << k;
        out[t] = (uint8_t) v;
    }
}

// predicted observable flips: obs = O x mod 2 for every row x of the decisions (what the reference's sinter driver
// computes per shot on the host, sinter_bposd_decoder.py:128-130), b8 rows [B][ceil(k/8)]
__global__ void observables_b8_kernel(const uint8_t *__restrict__ dec, long long batch, int n, int k, int kb,
                                      const uint32_t *__restrict__ obs_ptr, const uint32_t *__restrict__ obs_col,
                                      uint8_t *__restrict__ out) {
    const long long total = batch * kb;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const long long b = t / kb;
        const int byte = (int) (t - b * kb);
        const uint8_t *row = dec + b * n;
        uint32_t v = 0;
        for (int q = 0; q < 8; ++q) {
            const int o = 8 * byte + q;
            if (o >= k) break;
            uint32_t par = 0;
            for (uint32_t e = obs_ptr[o]; e < obs_ptr[o + 1]; ++e) par ^= row[obs_col[e]];
            v |= (par & 1u) << q;
        }
        out[t] = (uint8_t) v;
    }
}

using bpb::build_serial_batches;
using bpb::build_pair_plan;
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

// BP+OSD on the device: the batch indices of the syndromes BP did not solve (order is irrelevant: every entry is
// solved independently).  Optionally keeps a copy of the raw BP output rows for the caller.
__global__ void list_failures_kernel(const uint8_t *__restrict__ conv, long long batch, unsigned long long *count,
                                     unsigned long long *total, uint32_t *__restrict__ fail_idx) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long) gridDim.x * blockDim.x;
    const long long rounds = (batch + stride - 1) / stride;
    for (long long it = 0; it < rounds; ++it) {
        const long long b = it * stride + (long long) blockIdx.x * blockDim.x + threadIdx.x;
        const bool bad = b < batch && !conv[b];
        const uint32_t mask = __ballot_sync(0xffffffffu, bad);
        if (!mask) continue;
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(count, (unsigned long long) __popc(mask));
            atomicAdd(total, (unsigned long long) __popc(mask));  // running total over the chunks of one host call
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (bad) fail_idx[base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t) b;
    }
}

// ---- graph blob ---------------------------------------------------------------------------------------

int upload_graph(bpb_decoder *h) {
    const bpb::HostGraph &g = h->g;
    // Kernels of an earlier bpb_decode_batch_device call (asynchronous on the caller's stream) may still be reading
    // the tables rewritten below: drain the device first.  Only happens when the configuration changed.
    BPB_CUDA(h, cudaDeviceSynchronize());
    h->osd_plan = bpb::plan_osd_device(g, h->max_smem_optin);
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
    // SERIAL_RELATIVE: the configured order as it is (the kernel sorts its own copy per syndrome)
    rc = ensure(h, h->rel_order, h->serial_order.size() * sizeof(uint32_t));
    if (rc) return rc;
    BPB_CUDA(h, cudaMemcpy(h->rel_order.ptr, h->serial_order.data(), h->serial_order.size() * sizeof(uint32_t),
                           cudaMemcpyHostToDevice));
    rc = ensure(h, h->rel_order_out, h->serial_order.size() * sizeof(int32_t));
    if (rc) return rc;
    h->rel_order_valid = false;
    h->order_dirty = false;
    const int sb = bpb::serial_batch(g.max_row_degree, g.max_col_degree, g.regular);
    h->serial_batches = build_serial_batches(g, h->serial_order, sb);
    std::vector<uint32_t> upload = h->serial_batches;
    h->serial_entries = (int) h->serial_batches.size();
    h->serial_program_flags = false;
    if (g.regular && g.max_row_degree == 6 && g.max_col_degree == 3 && g.m < (1 << 20)) {
        // regular-code serial program (see bp_stream.cuh): {j, w_k for each of the three edges}, w_k = row | visited
        // flags << 20 | self << 28.  Flag f of an edge says that the f-th OTHER edge of its row belongs to a bit that
        // comes earlier in the program, i.e. has already been rewritten in the current sweep (levelisation keeps the
        // relative order of bits that share a check, so program order = reference order for that question).
        upload.assign(h->serial_batches.size() * 4, 0xffffffffu);
        std::vector<int> pos_of_bit((size_t) g.n, -1);
        bool covers_once = true;  // every bit at most once in the schedule (the flags assume it)
        for (size_t q = 0; q < h->serial_batches.size(); q++) {
            const uint32_t j = h->serial_batches[q];
            if (j == 0xffffffffu) continue;
            if (pos_of_bit[j] >= 0) covers_once = false;
            pos_of_bit[j] = (int) q;
        }
        for (int j = 0; j < g.n; j++)
            if (pos_of_bit[(size_t) j] < 0) covers_once = false;  // an unvisited bit keeps its prior for ever
        for (size_t q = 0; q < h->serial_batches.size(); q++) {
            const uint32_t j = h->serial_batches[q];
            if (j == 0xffffffffu) continue;
            upload[4 * q] = j;
            for (uint32_t e = g.col_ptr[j]; e < g.col_ptr[j + 1]; e++) {
                const uint32_t i = g.row_idx[e], self = g.csc2csr[e] - g.row_ptr[i];
                uint32_t flags = 0;
                int f = 0;
                for (uint32_t r = g.row_ptr[i]; r < g.row_ptr[i + 1]; r++) {
                    if (r == g.csc2csr[e]) continue;
                    if (pos_of_bit[g.col_idx[r]] >= 0 && pos_of_bit[g.col_idx[r]] < (int) q) flags |= 1u << f;
                    f++;
                }
                upload[4 * q + 1 + (e - g.col_ptr[j])] = i | (flags << 20) | (self << 28);
            }
        }
        h->serial_program_flags = covers_once;
    }
    rc = ensure(h, h->order_d, upload.size() * sizeof(uint32_t));
    if (rc) return rc;
    BPB_CUDA(h, cudaMemcpyAsync(h->order_d.ptr, upload.data(), upload.size() * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, h->stream));
    BPB_CUDA(h, cudaStreamSynchronize(h->stream));  // `upload` is pageable and dies at the end of this function
    build_smem_plan(h);
    if (h->smem_plan.ok) {
        rc = ensure(h, h->smem_tab, h->smem_plan.blob.size());
        if (rc) return rc;
        BPB_CUDA(h, cudaMemcpyAsync(h->smem_tab.ptr, h->smem_plan.blob.data(), h->smem_plan.blob.size(),
                                    cudaMemcpyHostToDevice, h->stream));
    }
    build_pair_plan(h);
    if (h->pair_plan.ok) {
        rc = ensure(h, h->pair_tab, h->pair_plan.blob.size());
        if (rc) return rc;
        BPB_CUDA(h, cudaMemcpyAsync(h->pair_tab.ptr, h->pair_plan.blob.data(), h->pair_plan.blob.size(),
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

bool pair_able(const bpb_decoder *h);
int launch_pair(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
                int32_t *d_iters, double *d_llr, cudaStream_t st, const uint32_t *index_list,
                const unsigned long long *batch_dev);

bool edge_able(const bpb_decoder *h);
int launch_edge(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
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
    if ((rc = ensure(h, h->msg_s[h->slot], warps * (size_t) g.nnz * 32 * sizeof(double) * (generic ? 2 : 1)))) return rc;
    if ((rc = ensure(h, h->dec_w_s[h->slot], warps * (size_t) n_pad * 4, true, st))) return rc;
    if ((rc = ensure(h, h->syn_w_s[h->slot], p.smem_syn ? 16 : warps * (size_t) m_pad * 4, true, st))) return rc;
    if (llr && (rc = ensure(h, h->llr_tile_s[h->slot], warps * (size_t) g.n * 32 * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->counter_s[h->slot], 64))) return rc;
    BPB_CUDA(h, cudaMemsetAsync(h->counter_s[h->slot].ptr, 0, 64, st));
    // second stage for the ramp-down (parallel schedule, when the thread-group kernels can take the code)
    // second stage for the ramp-down: the thread-group kernels when they can take the code, else (parallel schedule,
    // messages beyond shared memory, e.g. n = 10^4) the edge-parallel kernel with its messages in an L2-resident scratch
    const bool stage2_smem = h->smem_plan.ok && h->smem_plan.serial == (h->schedule == BPB_SERIAL);
    const bool stage2_edge = !stage2_smem && edge_able(h);
    const bool second_stage = (stage2_smem || stage2_edge) && h->max_iter > 16 && !std::getenv("BPB_NO_SECOND_STAGE");
    if (second_stage && (rc = ensure(h, h->handoff_s[h->slot], (size_t) warps * 32 * sizeof(uint32_t)))) return rc;
    p.iter_cap = second_stage ? 12 : h->max_iter + 1;
    p.handoff_count = (unsigned long long *) h->counter_s[h->slot].ptr + 1;
    p.handoff_list = (uint32_t *) h->handoff_s[h->slot].ptr;
    p.iter_total = (unsigned long long *) h->counter_s[h->slot].ptr + 3;

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
    p.counter = (unsigned long long *) h->counter_s[h->slot].ptr;
    p.msg = (double *) h->msg_s[h->slot].ptr;
    p.dec_w = (uint32_t *) h->dec_w_s[h->slot].ptr;
    p.syn_w_g = (uint32_t *) h->syn_w_s[h->slot].ptr;
    p.llr_tile = (double *) h->llr_tile_s[h->slot].ptr;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.order = (const uint32_t *) h->order_d.ptr;
    p.order_len = h->serial_entries;
    p.llr_last_only = h->llr_last_only ? 1 : 0;
    p.no_compaction = std::getenv("BPB_NO_COMPACTION") ? 1 : 0;
    p.compact_num = 2;
    p.compact_den = 1;
    if (const char *ov = std::getenv("BPB_COMPACT_RATIO")) {  // tuning: "num/den"
        int a = 0, b = 0;
        if (std::sscanf(ov, "%d/%d", &a, &b) == 2 && a > 0 && b > 0) {
            p.compact_num = a;
            p.compact_den = b;
        }
    }
    p.serial_no_init = (h->serial_program_flags && h->uniform_prior && !std::getenv("BPB_SERIAL_INIT")) ? 1 : 0;
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
        rc = stage2_smem ? launch_smem(h, d_packed, mwp, batch, d_dec, d_conv, d_iters, d_llr, st,
                                       (const uint32_t *) h->handoff_s[h->slot].ptr, (const unsigned long long *) h->counter_s[h->slot].ptr + 1)
                         : launch_edge(h, d_packed, mwp, batch, d_dec, d_conv, d_iters, d_llr, st,
                                       (const uint32_t *) h->handoff_s[h->slot].ptr, (const unsigned long long *) h->counter_s[h->slot].ptr + 1);
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
    if ((rc = ensure(h, h->counter_s[h->slot], 64))) return rc;
    // counter words: [0] streaming queue, [1] hand-off count, [2] thread-group queue
    if (!index_list) BPB_CUDA(h, cudaMemsetAsync(h->counter_s[h->slot].ptr, 0, 64, st));
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
    p.counter = (unsigned long long *) h->counter_s[h->slot].ptr + 2;
    p.index_list = index_list;
    p.batch_dev = batch_dev;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.llr_last_only = h->llr_last_only ? 1 : 0;
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

// ---- paired on-chip family: launch ------------------------------------------------------------------------------
// Thread groups per CTA and threads per group for a CTA of `maxt` threads.  Every thread keeps the decisions of its
// own columns in one register, two bits per column: a group must cover the columns in at most 16 rounds.
bool pair_geometry(const bpb_decoder *h, int maxt, int &G, int &T) {
    const bpb::PairPlan &pl = h->pair_plan;
    const bpb::HostGraph &g = h->g;
    if (!pl.ok || pl.blob.size() + pl.group_bytes > (size_t) h->max_smem_optin) return false;
    G = (int) (((size_t) h->max_smem_optin - pl.blob.size()) / pl.group_bytes);
    G = std::min(G, 15);  // named barriers 1..15
    T = maxt / G / 32 * 32;
    if (T < 32) {
        T = 32;
        G = maxt / 32;
    }
    // a thread per row / two columns is enough
    const int want = std::max(32, (int) align_up((uint32_t) std::max(g.m, (g.n + 1) / 2), 32));
    T = std::min(T, want);
    if (const char *ov = std::getenv("BPB_PAIR_GROUP_THREADS")) {  // tuning override: threads per group
        const int t_ov = std::atoi(ov);
        if (t_ov >= 32 && t_ov % 32 == 0 && t_ov <= maxt) {
            T = t_ov;
            G = std::min(G, maxt / T);
        }
    }
    return (g.n + T - 1) / T <= 16;
}

bool pair_able(const bpb_decoder *h) {
    int G, T;
    return h->schedule == BPB_PARALLEL && h->pair_plan.ok && pair_geometry(h, 512, G, T);
}

int launch_pair(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
                int32_t *d_iters, double *d_llr, cudaStream_t st, const uint32_t *index_list,
                const unsigned long long *batch_dev) {
    const bpb::HostGraph &g = h->g;
    const bpb::PairPlan &pl = h->pair_plan;
    const bool llr = d_llr != nullptr;
    bpb::PairKernel k = nullptr;
    int maxt = 512, G = 0, T = 0;
    if (const char *ov = std::getenv("BPB_PAIR_CTA_THREADS")) maxt = std::atoi(ov);  // tuning override: 512 | 640 | 768
    auto pick = [&](int cta) {
        return h->method == BPB_MINIMUM_SUM
                   ? bpb::pick_pair_ms(g.max_row_degree, g.max_col_degree, g.regular, llr, cta)
                   : bpb::pick_pair_ps(g.max_row_degree, g.max_col_degree, g.regular, llr, cta);
    };
    if (pair_able(h)) {
        k = pick(maxt);
        if (!k || !pair_geometry(h, maxt, G, T)) k = pick(maxt = 512);
    }
    if (!k || !pair_geometry(h, maxt, G, T)) {
        h->err = "paired on-chip kernel family not available for this code / schedule: " + pl.why;
        return BPB_ERR_UNSUPPORTED;
    }
    const size_t tab = pl.blob.size();
    const int block = G * T;
    const size_t smem_bytes = tab + (size_t) G * pl.group_bytes;
    BPB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes));
    int64_t grid64 = std::min<int64_t>(h->sm_count, (batch + 2 * G - 1) / (2 * G));
    if (grid64 < 1) grid64 = 1;
    int rc;
    if ((rc = ensure(h, h->counter_s[h->slot], 64))) return rc;
    // counter words: [0] streaming queue, [1] hand-off count, [2] thread-group queue
    if (!index_list) BPB_CUDA(h, cudaMemsetAsync(h->counter_s[h->slot].ptr, 0, 64, st));
    if (index_list) grid64 = h->sm_count;  // the count lives on the device
    bpb::PairParams p{};
    p.tab = (const uint32_t *) h->pair_tab.ptr;
    p.tab_bytes = (uint32_t) tab;
    p.off_row_deg = pl.off_row_deg;
    p.off_col_deg = pl.off_col_deg;
    p.off_col_row = pl.off_col_row;
    p.off_row_pos = pl.off_row_pos;
    p.off_col_pos = pl.off_col_pos;
    p.off_prior = pl.off_prior;
    p.group_bytes = pl.group_bytes;
    p.goff_msg = pl.goff_msg;
    p.goff_syn = pl.goff_syn;
    p.goff_acc = pl.goff_acc;
    p.goff_ctl = pl.goff_ctl;
    p.m = g.m;
    p.n = g.n;
    p.M = pl.M;
    p.N = pl.N;
    p.MW = (g.m + 31) / 32;
    p.msg_slots = pl.msg_slots;
    p.groups = G;
    p.T = T;
    p.max_iter = h->max_iter;
    p.ms_scaling = h->ms_scaling;
    p.uniform_prior = h->uniform_prior ? 1 : 0;
    p.prior0 = h->prior.empty() ? 0.0 : h->prior[0];
    p.synd_packed = d_packed;
    p.mwp = mwp;
    p.batch = batch;
    p.counter = (unsigned long long *) h->counter_s[h->slot].ptr + 2;
    p.index_list = index_list;
    p.batch_dev = batch_dev;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.llr_last_only = h->llr_last_only ? 1 : 0;
    if (!index_list) BPB_CUDA(h, cudaEventRecord(h->kev0, st));
    k<<<(int) grid64, block, smem_bytes, st>>>(p);
    BPB_CUDA(h, cudaGetLastError());
    if (!index_list) BPB_CUDA(h, cudaEventRecord(h->kev1, st));
    h->kernel_timed = true;
    h->launches += 1;
    h->last_family = BPB_KERNEL_PAIR;
    h->last_grid = (int) grid64;
    h->last_block = block;
    return BPB_OK;
}

// ---- edge-parallel family: launch ------------------------------------------------------------------------------
bool edge_able(const bpb_decoder *h) {
    return h->schedule == BPB_PARALLEL && h->g.max_row_degree <= 32 && h->g.max_col_degree <= 32 && h->g.nnz > 0;
}

int launch_edge(bpb_decoder *h, const uint32_t *d_packed, int mwp, int64_t batch, uint8_t *d_dec, uint8_t *d_conv,
                int32_t *d_iters, double *d_llr, cudaStream_t st, const uint32_t *index_list,
                const unsigned long long *batch_dev) {
    const bpb::HostGraph &g = h->g;
    if (!edge_able(h)) {
        h->err = "edge-parallel kernel family: parallel schedule and degrees <= 32 only";
        return BPB_ERR_UNSUPPORTED;
    }
    auto pow2_at_least = [](int x) {
        int p2 = 1;
        while (p2 < x) p2 <<= 1;
        return p2;
    };
    bpb::EdgeParams p{};
    p.G = pow2_at_least(std::max(1, g.max_row_degree));
    p.GV = pow2_at_least(std::max(1, g.max_col_degree));
    p.MW = (g.m + 31) / 32;
    const size_t fixed = (size_t) 2 * p.MW * 4 + (size_t) round_up(g.n, 16);
    const size_t msg_bytes = (size_t) g.nnz * 8;
    const bool msg_global = msg_bytes + fixed > (size_t) 100 * 1024;  // keep >= 2 CTAs per SM when messages are on chip
    const bool llr = d_llr != nullptr;
    bpb::EdgeKernel k = h->method == BPB_MINIMUM_SUM ? bpb::pick_edge_ms(llr, msg_global) : bpb::pick_edge_ps(llr, msg_global);
    const size_t smem_bytes = fixed + (msg_global ? 0 : msg_bytes);
    // threads per CTA: one lane per (row, slot) / (column, slot) when that fits; large codes loop
    int want = std::max(g.m * p.G, g.n * p.GV);
    int T = msg_global ? 1024 : 512;
    if (want < T) T = std::max(64, round_up(want, 32));
    if (const char *ov = std::getenv("BPB_EDGE_THREADS")) {
        const int t_ov = std::atoi(ov);
        if (t_ov >= 32 && t_ov % 32 == 0 && t_ov <= 1024) T = t_ov;
    }
    BPB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes));
    int occ = 0;
    BPB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, T, smem_bytes));
    if (occ < 1) {
        h->err = "edge kernel does not fit on an SM";
        return BPB_ERR_UNSUPPORTED;
    }
    int64_t grid64 = (int64_t) occ * h->sm_count;
    if (msg_global) {
        // keep the message scratch of all resident CTAs inside the L2 (about 100 MB of its 126 MB)
        const int64_t fit = std::max<int64_t>(1, ((int64_t) 100 << 20) / (int64_t) msg_bytes);
        grid64 = std::min(grid64, fit);
    }
    if (!index_list) grid64 = std::min<int64_t>(grid64, batch);
    if (grid64 < 1) grid64 = 1;
    int rc;
    if ((rc = ensure(h, h->counter_s[h->slot], 64))) return rc;
    if (!index_list) BPB_CUDA(h, cudaMemsetAsync(h->counter_s[h->slot].ptr, 0, 64, st));
    if (msg_global && (rc = ensure(h, h->edge_msg_s[h->slot], (size_t) grid64 * msg_bytes))) return rc;
    const uint32_t *blob = (const uint32_t *) h->blob.ptr;
    p.row_ptr = blob;
    p.col_idx = p.row_ptr + (g.m + 1);
    p.col_ptr = p.col_idx + g.nnz;
    p.csc2csr = p.col_ptr + (g.n + 1);
    p.row_idx = p.csc2csr + g.nnz;
    p.prior = reinterpret_cast<const double *>(blob + h->prior_off);
    p.m = g.m;
    p.n = g.n;
    p.nnz = g.nnz;
    p.max_iter = h->max_iter;
    p.ms_scaling = h->ms_scaling;
    p.uniform_prior = h->uniform_prior ? 1 : 0;
    p.prior0 = h->prior.empty() ? 0.0 : h->prior[0];
    p.synd_packed = d_packed;
    p.mwp = mwp;
    p.batch = batch;
    p.counter = (unsigned long long *) h->counter_s[h->slot].ptr + 2;  // word 2: the second-stage / thread-group queue
    p.index_list = index_list;
    p.batch_dev = batch_dev;
    p.msg_global = (double *) h->edge_msg_s[h->slot].ptr;
    p.out_dec = d_dec;
    p.out_conv = d_conv;
    p.out_iters = d_iters;
    p.out_llr = d_llr;
    p.llr_last_only = h->llr_last_only ? 1 : 0;
    if (!index_list) BPB_CUDA(h, cudaEventRecord(h->kev0, st));
    k<<<(int) grid64, T, smem_bytes, st>>>(p);
    BPB_CUDA(h, cudaGetLastError());
    if (!index_list) BPB_CUDA(h, cudaEventRecord(h->kev1, st));
    h->kernel_timed = true;
    h->launches += 1;
    h->last_family = BPB_KERNEL_EDGE;
    h->last_grid = (int) grid64;
    h->last_block = T;
    return BPB_OK;
}

constexpr int kInputPacked = 16;  // internal input type: bit-packed syndromes already on the device
int decode_device_core(bpb_decoder *h, int input_type, const uint8_t *d_input, int64_t batch, uint8_t *d_decoding,
                       uint8_t *d_converged, int32_t *d_iterations, double *d_llr, cudaStream_t cuda_stream);
int multi_device_decode(bpb_decoder *h, int input_type, const uint8_t *input, int64_t batch, uint8_t *decoding,
                        uint8_t *converged, int32_t *iterations, double *llr, uint8_t *bp_decoding, bool osd,
                        int threads);

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

const char *bpb_version(void) { return "ldpc_b200 0.2 (sm_100a)"; }

int bpb_libm_selfcheck(int samples) { return bpb::libm_selfcheck(samples > 0 ? samples : 20000); }

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
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
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
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void **) &h->host_counts, 64, cudaHostAllocDefault);
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
    for (bpb_decoder *c: h->children) bpb_destroy(c);
    h->children.clear();
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
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    for (int i = 0; i < 2; i++)
        if (h->pin_in[i]) cudaFreeHost(h->pin_in[i]);
    if (h->host_counts) cudaFreeHost(h->host_counts);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
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
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_channel(c, p, n);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_set_max_iter(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v < 1) {
        h->err = "maximum_iterations must be >= 1";
        return BPB_ERR_ARG;
    }
    h->max_iter = v;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_max_iter(c, v);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_set_method(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_PRODUCT_SUM && v != BPB_MINIMUM_SUM) {
        h->err = "invalid bp_method";
        return BPB_ERR_ARG;
    }
    h->method = v;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_method(c, v);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_set_schedule(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_SERIAL && v != BPB_PARALLEL && v != BPB_SERIAL_RELATIVE) {
        h->err = "Invalid BP schedule";  // bp.hpp:171
        return BPB_ERR_ARG;
    }
    if (h->schedule != v) h->graph_dirty = true;  // the shared-memory plan depends on the schedule
    h->schedule = v;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_schedule(c, v);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_set_ms_scaling_factor(bpb_decoder *h, double v) {
    if (!h) return BPB_ERR_ARG;
    h->ms_scaling = v;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_ms_scaling_factor(c, v);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_set_serial_schedule_order(bpb_decoder *h, const int32_t *order, int len) {
    if (!h) return BPB_ERR_ARG;
    if (!order) {
        h->serial_order.clear();
        h->graph_dirty = true;
        for (bpb_decoder *c: h->children) bpb_set_serial_schedule_order(c, nullptr, 0);
        return BPB_OK;
    }
    if (len < 0) return BPB_ERR_ARG;
    for (int i = 0; i < len; i++)
        if (order[i] < 0 || order[i] >= h->g.n) {
            h->err = "serial_schedule_order entry out of range";
            return BPB_ERR_ARG;
        }
    h->serial_order.assign(order, order + len);
    // SERIAL_RELATIVE only reads the raw schedule (no levelisation, no tables depend on it): re-upload just that
    if (h->schedule == BPB_SERIAL_RELATIVE && !h->graph_dirty && h->rel_order.bytes >= (size_t) len * 4)
        h->order_dirty = true;
    else
        h->graph_dirty = true;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_serial_schedule_order(c, order, len);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_get_last_schedule_order(bpb_decoder *h, int32_t *out, int len) {
    if (!h || !out) return BPB_ERR_ARG;
    if (h->device < 0 || !h->rel_order_valid || h->schedule != BPB_SERIAL_RELATIVE) {
        h->err = "no SERIAL_RELATIVE decode has run on this handle since the last configuration change";
        return BPB_ERR_ARG;
    }
    if (len != (int) h->serial_order.size()) {
        h->err = "schedule length mismatch";
        return BPB_ERR_ARG;
    }
    BPB_CUDA(h, cudaSetDevice(h->device));
    BPB_CUDA(h, cudaDeviceSynchronize());
    BPB_CUDA(h, cudaMemcpy(out, h->rel_order_out.ptr, sizeof(int32_t) * (size_t) len, cudaMemcpyDeviceToHost));
    return BPB_OK;
}

int bpb_set_kernel(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_KERNEL_AUTO && v != BPB_KERNEL_STREAM && v != BPB_KERNEL_SMEM && v != BPB_KERNEL_EDGE && v != BPB_KERNEL_PAIR) {
        h->err = "invalid kernel family";
        return BPB_ERR_ARG;
    }
    h->kernel_pref = v;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_kernel(c, v);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
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
    return decode_device_core(h, input_type, d_input, batch, d_decoding, d_converged, d_iterations, d_llr,
                              (cudaStream_t) cuda_stream);
}

}  // extern "C"

namespace {

// input_type kInputPacked: d_input is already the bit-packed syndrome array [batch][mwp] (on-device producers)
int decode_device_core(bpb_decoder *h, int input_type, const uint8_t *d_input, int64_t batch, uint8_t *d_decoding,
                       uint8_t *d_converged, int32_t *d_iterations, double *d_llr, cudaStream_t cuda_stream) {
    int rc = BPB_OK;
    if (batch == 0) return BPB_OK;
    BPB_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream;
    if (h->graph_dirty) {
        // drains the device, uploads on the handle's own stream and synchronises inside
        if ((rc = upload_graph(h))) return rc;
    }
    // The handle's workspaces (packed syndromes, counters, message tiles) are shared by consecutive calls: a call on
    // another stream waits for the previous call's work.
    // (the dual-stream host pipeline gives each of its two streams its own counters / packed buffer instead)
    if (h->have_last && h->last_stream != st && !h->pipeline_dual) BPB_CUDA(h, cudaStreamWaitEvent(st, h->ev_last, 0));
    const bpb::HostGraph &g = h->g;
    const int mwp = round_up((g.m + 31) / 32, 4);
    const uint32_t *d_packed = reinterpret_cast<const uint32_t *>(d_input);
    if (input_type != kInputPacked) {
        if ((rc = ensure(h, h->packed_s[h->slot], (size_t) batch * mwp * 4))) return rc;
        d_packed = (const uint32_t *) h->packed_s[h->slot].ptr;
    }
    h->last_packed = d_packed;
    uint32_t *d_pack_out = (uint32_t *) h->packed_s[h->slot].ptr;
    const int pgrid = (int) std::min<int64_t>((batch + 7) / 8, (int64_t) h->sm_count * 16);
    if (input_type == kInputPacked) {
    } else if (input_type == BPB_INPUT_SYNDROME) {
        pack_syndromes_kernel<<<pgrid, 256, 0, st>>>(d_input, batch, g.m, mwp, d_pack_out);
    } else {
        const uint32_t *blob = (const uint32_t *) h->blob.ptr;
        pack_received_kernel<<<pgrid, 256, 0, st>>>(d_input, batch, g.m, g.n, mwp, blob, blob + (g.m + 1), d_pack_out);
    }
    BPB_CUDA(h, cudaGetLastError());
    if (input_type != kInputPacked) h->launches += 1;
    if (h->schedule == BPB_SERIAL_RELATIVE) {
        // data-dependent schedule: its own kernel (one warp per syndrome, bp_relative.cu), whatever the family preference
        if (h->order_dirty) {
            BPB_CUDA(h, cudaStreamSynchronize(st));  // earlier decodes may still read the schedule
            BPB_CUDA(h, cudaMemcpy(h->rel_order.ptr, h->serial_order.data(), h->serial_order.size() * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice));
            h->order_dirty = false;
        }
        if ((rc = ensure(h, h->counter_s[h->slot], 64))) return rc;
        BPB_CUDA(h, cudaMemsetAsync(h->counter_s[h->slot].ptr, 0, 64, st));
        int grid = 0;
        BPB_CUDA(h, cudaEventRecord(h->kev0, st));
        const int e = bpb::launch_relative_kernel(g, h->sm_count, h->max_smem_optin, (const uint32_t *) h->blob.ptr,
                                                  h->prior_off, h->method, h->max_iter, h->ms_scaling,
                                                  (const uint32_t *) h->rel_order.ptr, (int) h->serial_order.size(),
                                                  d_packed, mwp, batch, (unsigned long long *) h->counter_s[h->slot].ptr + 2,
                                                  &h->rel_msg, d_decoding, d_converged, d_iterations, d_llr,
                                                  h->llr_last_only ? 1 : 0, (int32_t *) h->rel_order_out.ptr, st, &grid);
        if (e == -1) {
            h->err = "SERIAL_RELATIVE: the code does not fit the kernel (column degree > 32 or n beyond shared memory)";
            return BPB_ERR_UNSUPPORTED;
        }
        if (e) {
            h->err = std::string("bp_relative_kernel launch: ") + cudaGetErrorString((cudaError_t) e);
            return BPB_ERR_CUDA;
        }
        BPB_CUDA(h, cudaEventRecord(h->kev1, st));
        h->kernel_timed = true;
        h->rel_order_valid = true;
        h->launches += 1;
        h->last_family = BPB_KERNEL_EDGE;  // a cooperative (warp per syndrome) kernel
        h->last_grid = grid;
        h->last_block = 32;
        if (input_type == BPB_INPUT_RECEIVED_VECTOR) {
            const long long count = (long long) batch * g.n;
            const int xgrid = (int) std::min<long long>((count + 255) / 256, (long long) h->sm_count * 32);
            xor_received_kernel<<<xgrid, 256, 0, st>>>(d_decoding, d_input, count);
            BPB_CUDA(h, cudaGetLastError());
            h->launches += 1;
        }
        BPB_CUDA(h, cudaEventRecord(h->ev_last, st));
        h->last_stream = st;
        h->have_last = true;
        return BPB_OK;
    }
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
    if (h->kernel_pref == BPB_KERNEL_EDGE && !edge_able(h)) {
        h->err = "kernel family 'edge' serves the parallel schedule with degrees <= 32 only";
        return BPB_ERR_UNSUPPORTED;
    }
    // AUTO, parallel schedule, small batch, a code the on-chip families cannot hold: one CTA per syndrome with a lane
    // per edge and L2-resident messages instead of a streaming-kernel launch that would leave the lanes of every warp
    // but one idle.  (For codes that do fit, the thread-group kernels are faster at every batch size, also for a
    // single syndrome: 44 vs 89 us at n = 1000, profiles/r2_latency_c2.jsonl.)
    const bool small_batch = batch <= (int64_t) h->sm_count * 2;
    const bool use_edge = h->kernel_pref == BPB_KERNEL_EDGE ||
                          (h->kernel_pref == BPB_KERNEL_AUTO && edge_able(h) && small_batch && !smem_able &&
                           !pair_able(h) && !std::getenv("BPB_NO_EDGE_AUTO"));
    if (h->kernel_pref == BPB_KERNEL_PAIR && !pair_able(h)) {
        h->err = "kernel family 'pair' serves the parallel schedule of codes whose messages fit in shared memory twice: " +
                 h->pair_plan.why;
        return BPB_ERR_UNSUPPORTED;
    }
    // AUTO: the paired kernels (two syndromes per thread group) when they can take the code: config 2 (min-sum)
    // 28.7 vs 22.4 M decodes/s for one-syndrome groups, config 3 (product-sum, arithmetic-bound) 4.83 vs 4.59
    const bool use_pair = !use_edge && pair_able(h) &&
                          (h->kernel_pref == BPB_KERNEL_PAIR ||
                           (h->kernel_pref == BPB_KERNEL_AUTO && !std::getenv("BPB_NO_PAIR_AUTO")));
    const bool use_smem = !use_edge && !use_pair && smem_able &&
                          (h->kernel_pref == BPB_KERNEL_SMEM ||
                           (h->kernel_pref == BPB_KERNEL_AUTO && h->schedule == BPB_PARALLEL));
    if (use_edge)
        rc = launch_edge(h, d_packed, mwp, batch, d_decoding, d_converged, d_iterations, d_llr, st, nullptr, nullptr);
    else if (use_pair)
        rc = launch_pair(h, d_packed, mwp, batch, d_decoding, d_converged, d_iterations, d_llr, st, nullptr, nullptr);
    else if (use_smem)
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
    BPB_CUDA(h, cudaEventRecord(h->ev_last, st));
    h->last_stream = st;
    h->have_last = true;
    return BPB_OK;
}

bool osd_on_device(const bpb_decoder *h) {
    return h->osd_location != BPB_OSD_HOST && h->osd_plan.warps_per_cta > 0;
}

// BP + OSD-0 for one chunk that is already in device memory; everything is enqueued on `st`.
// osd_count words: [0] failures of this chunk, [1] OSD work queue, [2] running total of the host call.
int enqueue_bposd_device(bpb_decoder *h, const uint8_t *d_syn, int64_t nb, uint8_t *d_dec, uint8_t *d_conv,
                         int32_t *d_its, uint8_t *d_bp_dec, cudaStream_t st, int input_type = BPB_INPUT_SYNDROME) {
    const bpb::HostGraph &g = h->g;
    int rc;
    if ((rc = ensure(h, h->osd_llr, (size_t) nb * g.n * 8))) return rc;
    if ((rc = ensure(h, h->osd_fail_idx, (size_t) nb * 4))) return rc;
    if ((rc = ensure(h, h->osd_count, 64, true, st))) return rc;
    if (!d_conv) {
        if ((rc = ensure(h, h->osd_conv, (size_t) nb))) return rc;
        d_conv = (uint8_t *) h->osd_conv.ptr;
    }
    h->llr_last_only = true;  // only syndromes that ran all maximum_iterations need their posterior LLRs
    rc = decode_device_core(h, input_type, d_syn, nb, d_dec, d_conv, d_its, (double *) h->osd_llr.ptr, st);
    h->llr_last_only = false;
    if (rc) return rc;
    if (d_bp_dec) BPB_CUDA(h, cudaMemcpyAsync(d_bp_dec, d_dec, (size_t) nb * g.n, cudaMemcpyDeviceToDevice, st));
    unsigned long long *cnt = (unsigned long long *) h->osd_count.ptr;
    BPB_CUDA(h, cudaMemsetAsync(cnt, 0, 16, st));
    const int lgrid = (int) std::min<int64_t>((nb + 255) / 256, (int64_t) h->sm_count * 8);
    list_failures_kernel<<<lgrid, 256, 0, st>>>(d_conv, nb, cnt, cnt + 2, (uint32_t *) h->osd_fail_idx.ptr);
    BPB_CUDA(h, cudaGetLastError());
    const uint32_t *blob = (const uint32_t *) h->blob.ptr;
    const int mwp = round_up((g.m + 31) / 32, 4);
    const int e = bpb::launch_osd0_kernel(h->osd_plan, g, h->sm_count, blob, blob + (g.m + 1),
                                          h->last_packed, mwp, (const double *) h->osd_llr.ptr,
                                          (const uint32_t *) h->osd_fail_idx.ptr, cnt, cnt + 1, d_dec, nb, st);
    if (e) {
        h->err = std::string("osd0_kernel launch: ") + cudaGetErrorString((cudaError_t) e);
        return BPB_ERR_CUDA;
    }
    h->launches += 2;
    BPB_CUDA(h, cudaEventRecord(h->ev_last, st));
    return BPB_OK;
}

// memcpy on a few threads (one thread moves ~10 GB/s, the H2D link takes 25+)
void parallel_copy(uint8_t *dst, const uint8_t *src, size_t bytes) {
    const size_t min_part = (size_t) 4 << 20;
    int parts = (int) std::min<size_t>(4, std::max<size_t>(1, bytes / min_part));
    const unsigned hw = std::thread::hardware_concurrency();
    if (hw && (unsigned) parts > hw) parts = (int) hw;
    if (parts <= 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 1; t < parts; t++) {
        const size_t a = bytes * t / parts, b = bytes * (t + 1) / parts;
        pool.emplace_back([=] { std::memcpy(dst + a, src + a, b - a); });
    }
    std::memcpy(dst, src, bytes / parts);
    for (auto &th: pool) th.join();
}

// Rows per chunk of the host pipelines from a byte budget per staging slot (the fixed row counts of round 1 asked
// for tens of GB at n = 10^4 with LLRs).
int64_t chunk_rows(int64_t want, size_t row_bytes, int64_t batch) {
    const size_t budget = (size_t) 2 << 30;
    int64_t rows = (int64_t) (budget / std::max<size_t>(row_bytes, 1));
    rows = std::max<int64_t>(rows, 1024);
    rows = std::min<int64_t>(rows, want);
    return std::max<int64_t>(1, std::min<int64_t>(rows, batch));
}

// Chunked three-stage pipeline (H2D | kernels | D2H) over two staging slots.  The on-chip family keeps no per-lane
// state in HBM, so small chunks cost nothing; the streaming family amortises its persistent-lane ramp-down over large
// chunks.  with_osd: BP + OSD-0 with the elimination on the device.
int host_pipeline(bpb_decoder *h, int input_type, const uint8_t *input, int64_t batch, uint8_t *decoding,
                  uint8_t *converged, int32_t *iterations, double *llr, uint8_t *bp_decoding, bool with_osd) {
    const bpb::HostGraph &g = h->g;
    int rc = BPB_OK;
    const int in_w = (input_type == BPB_INPUT_RECEIVED_VECTOR) ? g.n : g.m;
    const bool smem_able = h->smem_plan.ok && (h->kernel_pref == BPB_KERNEL_SMEM || h->kernel_pref == BPB_KERNEL_EDGE ||
                                               h->kernel_pref == BPB_KERNEL_PAIR ||
                                               (h->kernel_pref == BPB_KERNEL_AUTO && h->schedule == BPB_PARALLEL));
    const size_t row_bytes = (size_t) in_w + (size_t) g.n * (bp_decoding ? 2 : 1) + 5 +
                             ((llr || with_osd) ? (size_t) g.n * 8 : 0);
    const size_t io_bytes = (size_t) in_w + (size_t) g.n * (bp_decoding ? 2 : 1) + 5 + (llr ? (size_t) g.n * 8 : 0);
    // on-chip families: the first chunk's H2D and the last chunk's D2H are not hidden behind a kernel, so keep chunks
    // small (2^16 rows = 1/16 of the 2^20 workload); each chunk still gives every SM hundreds of syndromes
    // (2^16 rows of n = 1000, but at least 32 MB of input + output per chunk: small codes pay per-chunk overheads)
    int64_t want_rows = smem_able ? std::min<int64_t>((int64_t) 1 << 18,
                                                      std::max<int64_t>((int64_t) 1 << 16,
                                                                        ((int64_t) 32 << 20) / (int64_t) io_bytes))
                                  : ((int64_t) 1 << 20);
    if (const char *ov = std::getenv("BPB_CHUNK_ROWS")) want_rows = std::max<int64_t>(1024, std::atoll(ov));  // tuning
    const int64_t chunk_max = chunk_rows(want_rows, row_bytes, batch);
    const size_t cap = (size_t) chunk_max;
    // Pageable input (e.g. a plain numpy array): cudaMemcpyAsync would go through the driver's single staging buffer
    // at a few GB/s and block.  Stage it ourselves: a few host threads copy the chunk into a pinned slot buffer while
    // the GPU works on the previous chunk, then a true asynchronous H2D follows.
    bool stage_input = false;
    {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, input) != cudaSuccess) cudaGetLastError();
        else stage_input = (attr.type == cudaMemoryTypeUnregistered);
    }
    if (stage_input) {
        for (int s = 0; s < 2; s++) {
            if (h->pin_in_bytes[s] >= cap * (size_t) in_w) continue;
            if (h->pin_in[s]) cudaFreeHost(h->pin_in[s]);
            h->pin_in[s] = nullptr;
            h->pin_in_bytes[s] = 0;
            if (cudaHostAlloc(&h->pin_in[s], cap * (size_t) in_w, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                stage_input = false;  // cannot pin that much: let the driver stage
                break;
            }
            h->pin_in_bytes[s] = cap * (size_t) in_w;
        }
    }
    cudaError_t ce = cudaSuccess;
    auto fail_cuda = [&](const char *what) {
        h->err = std::string(what) + ": " + cudaGetErrorString(ce);
        rc = BPB_ERR_CUDA;
    };
#define PIPE_CUDA(call)                 \
    if (rc == BPB_OK) {                 \
        ce = (call);                    \
        if (ce != cudaSuccess) fail_cuda(#call); \
    }
    if (with_osd) {
        if ((rc = ensure(h, h->osd_count, 64, true, h->stream))) return rc;
        PIPE_CUDA(cudaMemsetAsync((unsigned long long *) h->osd_count.ptr + 2, 0, 8, h->stream));
    }
    // The kernels of consecutive chunks may run concurrently: two compute streams, one per staging slot, each with its
    // own work-queue counters, packed syndromes and (streaming family) message tiles.  The CTAs of chunk c+1 start on
    // the SMs that chunk c's CTAs leave, which hides the ramp-down of every chunk: 50-iteration non-convergers that keep
    // a few SMs busy for ~0.2 ms in the on-chip families, the half-empty warps of the last iterations in the streaming
    // family.  BP+OSD shares its failure lists between chunks and stays on one stream.
    // (streaming family: a second set of message tiles only when one set is modest, 2368 warps x E x 256 bytes)
    const bool tiles_ok = smem_able || (double) g.nnz * 256.0 * 16.0 * (double) h->sm_count < 40e9;
    const bool dual = !with_osd && h->schedule != BPB_SERIAL_RELATIVE && h->kernel_pref != BPB_KERNEL_EDGE && tiles_ok &&
                      batch > (int64_t) h->sm_count * 2 && !std::getenv("BPB_NO_DUAL_STREAM");
    if (dual && h->have_last) {
        PIPE_CUDA(cudaStreamWaitEvent(h->stream, h->ev_last, 0));
        PIPE_CUDA(cudaStreamWaitEvent(h->stream2, h->ev_last, 0));
    }
    h->pipeline_dual = dual;
    int64_t c = 0;
    for (int64_t lo = 0; lo < batch && rc == BPB_OK; lo += chunk_max, ++c) {
        const int s = (int) (c & 1);
        const cudaStream_t cs = (dual && s) ? h->stream2 : h->stream;
        h->slot = dual ? s : 0;
        const int64_t nb = std::min(chunk_max, batch - lo);
        if ((rc = ensure(h, h->st_in[s], cap * in_w))) break;
        if ((rc = ensure(h, h->st_dec[s], cap * g.n))) break;
        if ((rc = ensure(h, h->st_conv[s], cap))) break;
        if ((rc = ensure(h, h->st_iters[s], cap * 4))) break;
        if (llr && (rc = ensure(h, h->st_llr[s], cap * g.n * 8))) break;
        if (bp_decoding && (rc = ensure(h, h->st_bp[s], cap * g.n))) break;
        if (c >= 2) PIPE_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_k[s], 0));  // slot's input consumed
        const uint8_t *src = input + lo * in_w;
        if (stage_input) {
            if (c >= 2) PIPE_CUDA(cudaEventSynchronize(h->ev_in[s]));  // the slot's previous H2D has left the buffer
            parallel_copy((uint8_t *) h->pin_in[s], src, (size_t) nb * in_w);
            src = (const uint8_t *) h->pin_in[s];
        }
        PIPE_CUDA(cudaMemcpyAsync(h->st_in[s].ptr, src, (size_t) nb * in_w, cudaMemcpyHostToDevice, h->s_in));
        PIPE_CUDA(cudaEventRecord(h->ev_in[s], h->s_in));
        PIPE_CUDA(cudaStreamWaitEvent(cs, h->ev_in[s], 0));
        if (c >= 2) PIPE_CUDA(cudaStreamWaitEvent(cs, h->ev_out[s], 0));  // slot's outputs drained
        if (rc) break;
        if (with_osd)
            rc = enqueue_bposd_device(h, (const uint8_t *) h->st_in[s].ptr, nb, (uint8_t *) h->st_dec[s].ptr,
                                      (uint8_t *) h->st_conv[s].ptr, (int32_t *) h->st_iters[s].ptr,
                                      bp_decoding ? (uint8_t *) h->st_bp[s].ptr : nullptr, cs);
        else
            rc = bpb_decode_batch_device(h, input_type, (const uint8_t *) h->st_in[s].ptr, nb,
                                         (uint8_t *) h->st_dec[s].ptr, (uint8_t *) h->st_conv[s].ptr,
                                         (int32_t *) h->st_iters[s].ptr, llr ? (double *) h->st_llr[s].ptr : nullptr,
                                         cs);
        if (rc) break;
        PIPE_CUDA(cudaEventRecord(h->ev_k[s], cs));
        PIPE_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_k[s], 0));
        PIPE_CUDA(cudaMemcpyAsync(decoding + lo * g.n, h->st_dec[s].ptr, (size_t) nb * g.n, cudaMemcpyDeviceToHost,
                                  h->s_out));
        if (converged)
            PIPE_CUDA(cudaMemcpyAsync(converged + lo, h->st_conv[s].ptr, (size_t) nb, cudaMemcpyDeviceToHost, h->s_out));
        if (iterations)
            PIPE_CUDA(cudaMemcpyAsync(iterations + lo, h->st_iters[s].ptr, (size_t) nb * 4, cudaMemcpyDeviceToHost,
                                      h->s_out));
        if (llr)
            PIPE_CUDA(cudaMemcpyAsync(llr + lo * g.n, h->st_llr[s].ptr, (size_t) nb * g.n * 8, cudaMemcpyDeviceToHost,
                                      h->s_out));
        if (bp_decoding)
            PIPE_CUDA(cudaMemcpyAsync(bp_decoding + lo * g.n, h->st_bp[s].ptr, (size_t) nb * g.n,
                                      cudaMemcpyDeviceToHost, h->s_out));
        PIPE_CUDA(cudaEventRecord(h->ev_out[s], h->s_out));
    }
#undef PIPE_CUDA
    // also on the error path: no copy into the caller's memory may be in flight when this returns
    cudaStreamSynchronize(h->s_in);
    cudaError_t e1 = cudaStreamSynchronize(h->stream);
    if (dual) {
        const cudaError_t e1b = cudaStreamSynchronize(h->stream2);
        if (e1 == cudaSuccess) e1 = e1b;
    }
    h->pipeline_dual = false;
    h->slot = 0;
    h->have_last = false;  // everything this call enqueued has finished
    cudaError_t e2 = cudaStreamSynchronize(h->s_out);
    if (rc == BPB_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        h->err = std::string("pipeline synchronise: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
        rc = BPB_ERR_CUDA;
    }
    if (rc == BPB_OK && with_osd) {
        if (cudaMemcpy(h->host_counts, (unsigned long long *) h->osd_count.ptr + 2, 8, cudaMemcpyDeviceToHost) ==
            cudaSuccess)
            h->osd_device_solved += (int64_t) h->host_counts[0];
    }
    return rc;
}

}  // namespace

extern "C" {

int bpb_decode_batch(bpb_decoder *h, int input_type, const uint8_t *input, int64_t batch, uint8_t *decoding,
                     uint8_t *converged, int32_t *iterations, double *llr) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!input || !decoding))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (batch == 0) return BPB_OK;
    if (!h->children.empty())
        return multi_device_decode(h, input_type, input, batch, decoding, converged, iterations, llr, nullptr, false, 0);
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    return host_pipeline(h, input_type, input, batch, decoding, converged, iterations, llr, nullptr, false);
}

int bpb_osd0_host(bpb_decoder *h, const uint8_t *syndromes, const double *llr, const uint8_t *converged,
                  int64_t batch, uint8_t *decoding, int threads) {
    if (!h) return BPB_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!syndromes || !llr || !decoding))) {
        h->err = "bad osd arguments";
        return BPB_ERR_ARG;
    }
    int64_t outside = 0;
    int rc = bpb::osd0_host(h->g, syndromes, llr, converged, batch, decoding, threads, &outside);
    if (rc) h->err = "osd0_host failed";
    h->osd_host_inconsistent += outside;
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
    if (!h->children.empty())
        return multi_device_decode(h, BPB_INPUT_SYNDROME, syndromes, batch, decoding, converged, iterations, nullptr,
                                   bp_decoding, true, threads);
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    const bpb::HostGraph &g = h->g;
    if (h->osd_location == BPB_OSD_DEVICE && h->osd_plan.warps_per_cta == 0) {
        h->err = "OSD-0 on the device requested but this code does not fit the kernel (m > 1024 or matrix beyond shared memory)";
        return BPB_ERR_UNSUPPORTED;
    }
    if (osd_on_device(h))
        return host_pipeline(h, BPB_INPUT_SYNDROME, syndromes, batch, decoding, converged, iterations, nullptr,
                             bp_decoding, true);
    // host elimination: only the LLR rows of the non-converged syndromes cross PCIe
    const int64_t chunk_max = chunk_rows((int64_t) 1 << 17, (size_t) g.n * 16 + g.m + g.n + 5, batch);
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
            int64_t outside = 0;
            rc = bpb::osd0_host(g, fsyn.data(), fllr.data(), nullptr, (int64_t) nfail, fdec.data(), threads, &outside);
            h->osd_host_inconsistent += outside;
            if (rc) {
                h->err = "osd0_host failed";
                return rc;
            }
            h->osd_host_solved += (int64_t) nfail;
            for (size_t q = 0; q < nfail; q++)
                std::memcpy(decoding + (lo + fidx[q]) * g.n, &fdec[q * g.n], (size_t) g.n);
        }
    }
    return BPB_OK;
}

int bpb_bposd_decode_batch_device(bpb_decoder *h, const uint8_t *d_syndromes, int64_t batch, uint8_t *d_decoding,
                                  uint8_t *d_converged, int32_t *d_iterations, uint8_t *d_bp_decoding,
                                  void *cuda_stream) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!d_syndromes || !d_decoding))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (batch == 0) return BPB_OK;
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    if (h->osd_plan.warps_per_cta == 0 || h->osd_location == BPB_OSD_HOST) {
        h->err = "OSD-0 on the device is not available for this code / configuration";
        return BPB_ERR_UNSUPPORTED;
    }
    return enqueue_bposd_device(h, d_syndromes, batch, d_decoding, d_converged, d_iterations, d_bp_decoding,
                                (cudaStream_t) cuda_stream);
}

int bpb_set_observables(bpb_decoder *h, int k, int64_t nnz, const int32_t *rows, const int32_t *cols) {
    if (!h) return BPB_ERR_ARG;
    if (k < 0 || nnz < 0 || (nnz > 0 && (!rows || !cols))) {
        h->err = "bad observables matrix";
        return BPB_ERR_ARG;
    }
    std::vector<uint32_t> ptr((size_t) k + 1, 0u), col((size_t) nnz);
    for (int64_t e = 0; e < nnz; e++) {
        if (rows[e] < 0 || rows[e] >= k || cols[e] < 0 || cols[e] >= h->g.n) {
            h->err = "observables matrix entry out of range";
            return BPB_ERR_ARG;
        }
        ptr[(size_t) rows[e] + 1]++;
    }
    for (int o = 0; o < k; o++) ptr[(size_t) o + 1] += ptr[(size_t) o];
    std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < nnz; e++) col[fill[(size_t) rows[e]]++] = (uint32_t) cols[e];
    h->obs_k = k;
    h->obs_ptr = ptr;
    h->obs_col = col;
    h->obs_dirty = true;
    for (bpb_decoder *c: h->children) {
        const int rc_c = bpb_set_observables(c, k, nnz, rows, cols);
        if (rc_c) {
            h->err = c->err;
            return rc_c;
        }
    }
    return BPB_OK;
}

int bpb_decode_batch_b8(bpb_decoder *h, int with_osd, const uint8_t *syndromes_b8, int64_t batch, uint8_t *decoding_b8,
                        uint8_t *observables_b8, uint8_t *converged, int32_t *iterations) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!syndromes_b8 || (!decoding_b8 && !observables_b8)))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (observables_b8 && h->obs_k <= 0) {
        h->err = "observables requested but bpb_set_observables has not been called";
        return BPB_ERR_ARG;
    }
    if (batch == 0) return BPB_OK;
    const bpb::HostGraph &g = h->g;
    const int mb = (g.m + 7) / 8, nb8 = (g.n + 7) / 8, kb = (h->obs_k + 7) / 8;
    if (!h->children.empty()) {
        // contiguous slices of the shots, one host thread per device
        const int k = (int) h->children.size();
        std::vector<int> rcs((size_t) k, BPB_OK);
        auto work = [&](int r) {
            const int64_t lo = batch * r / k, hi = batch * (r + 1) / k;
            if (hi > lo)
                rcs[(size_t) r] = bpb_decode_batch_b8(h->children[(size_t) r], with_osd, syndromes_b8 + lo * mb, hi - lo,
                                                      decoding_b8 ? decoding_b8 + lo * nb8 : nullptr,
                                                      observables_b8 ? observables_b8 + lo * kb : nullptr,
                                                      converged ? converged + lo : nullptr,
                                                      iterations ? iterations + lo : nullptr);
        };
        std::vector<std::thread> pool;
        for (int r = 1; r < k; r++) pool.emplace_back(work, r);
        work(0);
        for (auto &t: pool) t.join();
        for (int r = 0; r < k; r++)
            if (rcs[(size_t) r]) {
                h->err = "device slice " + std::to_string(r) + ": " + h->children[(size_t) r]->err;
                return rcs[(size_t) r];
            }
        return BPB_OK;
    }
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    if (with_osd && !osd_on_device(h)) {
        h->err = "bpb_decode_batch_b8 with OSD-0 needs the device OSD-0 kernel, which is not available for this code";
        return BPB_ERR_UNSUPPORTED;
    }
    if (observables_b8 && h->obs_dirty) {
        const size_t words = h->obs_ptr.size() + h->obs_col.size();
        if ((rc = ensure(h, h->obs_tab, words * 4))) return rc;
        BPB_CUDA(h, cudaDeviceSynchronize());
        BPB_CUDA(h, cudaMemcpy(h->obs_tab.ptr, h->obs_ptr.data(), h->obs_ptr.size() * 4, cudaMemcpyHostToDevice));
        if (!h->obs_col.empty())
            BPB_CUDA(h, cudaMemcpy((uint32_t *) h->obs_tab.ptr + h->obs_ptr.size(), h->obs_col.data(),
                                   h->obs_col.size() * 4, cudaMemcpyHostToDevice));
        h->obs_dirty = false;
    }
    const int mwp = round_up((g.m + 31) / 32, 4);
    // rows per chunk: the unpacked decisions of a chunk stay on the device (n bytes per row), so bound those
    const int64_t chunk_max = std::max<int64_t>(1024, std::min<int64_t>({batch, (int64_t) 1 << 18,
                                                                         ((int64_t) 512 << 20) / std::max(g.n, 1)}));
    const size_t cap = (size_t) chunk_max;
    cudaError_t ce = cudaSuccess;
#define B8_CUDA(call)                                                             \
    if (rc == BPB_OK) {                                                           \
        ce = (call);                                                              \
        if (ce != cudaSuccess) {                                                  \
            h->err = std::string(#call) + ": " + cudaGetErrorString(ce);          \
            rc = BPB_ERR_CUDA;                                                    \
        }                                                                         \
    }
    if (with_osd) {
        if ((rc = ensure(h, h->osd_count, 64, true, h->stream))) return rc;
        B8_CUDA(cudaMemsetAsync((unsigned long long *) h->osd_count.ptr + 2, 0, 8, h->stream));
    }
    int64_t c = 0;
    for (int64_t lo = 0; lo < batch && rc == BPB_OK; lo += chunk_max, ++c) {
        const int s = (int) (c & 1);
        const int64_t nb = std::min(chunk_max, batch - lo);
        if ((rc = ensure(h, h->b8_in[s], cap * mb))) break;
        if ((rc = ensure(h, h->b8_words[s], cap * mwp * 4))) break;
        if ((rc = ensure(h, h->st_dec[s], cap * g.n))) break;
        if ((rc = ensure(h, h->st_conv[s], cap))) break;
        if ((rc = ensure(h, h->st_iters[s], cap * 4))) break;
        if (decoding_b8 && (rc = ensure(h, h->b8_out[s], cap * nb8))) break;
        if (observables_b8 && (rc = ensure(h, h->b8_obs[s], cap * std::max(kb, 1)))) break;
        if (c >= 2) B8_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_k[s], 0));  // slot's input consumed
        B8_CUDA(cudaMemcpyAsync(h->b8_in[s].ptr, syndromes_b8 + lo * mb, (size_t) nb * mb, cudaMemcpyHostToDevice, h->s_in));
        B8_CUDA(cudaEventRecord(h->ev_in[s], h->s_in));
        B8_CUDA(cudaStreamWaitEvent(h->stream, h->ev_in[s], 0));
        if (c >= 2) B8_CUDA(cudaStreamWaitEvent(h->stream, h->ev_out[s], 0));  // slot's outputs drained
        if (rc) break;
        const int ugrid = (int) std::min<int64_t>((nb * mwp + 255) / 256, (int64_t) h->sm_count * 16);
        unpack_b8_rows_kernel<<<ugrid, 256, 0, h->stream>>>((const uint8_t *) h->b8_in[s].ptr, nb, g.m, mb, mwp,
                                                            (uint32_t *) h->b8_words[s].ptr);
        B8_CUDA(cudaGetLastError());
        h->launches += 1;
        if (rc) break;
        if (with_osd)
            rc = enqueue_bposd_device(h, (const uint8_t *) h->b8_words[s].ptr, nb, (uint8_t *) h->st_dec[s].ptr,
                                      (uint8_t *) h->st_conv[s].ptr, (int32_t *) h->st_iters[s].ptr, nullptr,
                                      h->stream, kInputPacked);
        else
            rc = decode_device_core(h, kInputPacked, (const uint8_t *) h->b8_words[s].ptr, nb,
                                    (uint8_t *) h->st_dec[s].ptr, (uint8_t *) h->st_conv[s].ptr,
                                    (int32_t *) h->st_iters[s].ptr, nullptr, h->stream);
        if (rc) break;
        if (decoding_b8) {
            const int pgrid = (int) std::min<int64_t>((nb * nb8 + 255) / 256, (int64_t) h->sm_count * 16);
            pack_b8_rows_kernel<<<pgrid, 256, 0, h->stream>>>((const uint8_t *) h->st_dec[s].ptr, nb, g.n, nb8,
                                                              (uint8_t *) h->b8_out[s].ptr);
            B8_CUDA(cudaGetLastError());
            h->launches += 1;
        }
        if (observables_b8) {
            const int ogrid = (int) std::min<int64_t>((nb * kb + 255) / 256, (int64_t) h->sm_count * 16);
            const uint32_t *optr = (const uint32_t *) h->obs_tab.ptr;
            observables_b8_kernel<<<ogrid, 256, 0, h->stream>>>((const uint8_t *) h->st_dec[s].ptr, nb, g.n, h->obs_k, kb,
                                                                optr, optr + h->obs_ptr.size(),
                                                                (uint8_t *) h->b8_obs[s].ptr);
            B8_CUDA(cudaGetLastError());
            h->launches += 1;
        }
        B8_CUDA(cudaEventRecord(h->ev_k[s], h->stream));
        B8_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_k[s], 0));
        if (decoding_b8)
            B8_CUDA(cudaMemcpyAsync(decoding_b8 + lo * nb8, h->b8_out[s].ptr, (size_t) nb * nb8, cudaMemcpyDeviceToHost, h->s_out));
        if (observables_b8)
            B8_CUDA(cudaMemcpyAsync(observables_b8 + lo * kb, h->b8_obs[s].ptr, (size_t) nb * kb, cudaMemcpyDeviceToHost, h->s_out));
        if (converged)
            B8_CUDA(cudaMemcpyAsync(converged + lo, h->st_conv[s].ptr, (size_t) nb, cudaMemcpyDeviceToHost, h->s_out));
        if (iterations)
            B8_CUDA(cudaMemcpyAsync(iterations + lo, h->st_iters[s].ptr, (size_t) nb * 4, cudaMemcpyDeviceToHost, h->s_out));
        B8_CUDA(cudaEventRecord(h->ev_out[s], h->s_out));
    }
#undef B8_CUDA
    // also on the error path: no copy into the caller's memory may be in flight when this returns
    cudaStreamSynchronize(h->s_in);
    const cudaError_t e1 = cudaStreamSynchronize(h->stream);
    const cudaError_t e2 = cudaStreamSynchronize(h->s_out);
    if (rc == BPB_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        h->err = std::string("pipeline synchronise: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
        rc = BPB_ERR_CUDA;
    }
    h->have_last = false;
    return rc;
}

int bpb_soft_info_decode_batch(bpb_decoder *h, const double *soft_syndromes, int64_t batch, double cutoff, double sigma,
                               uint8_t *decoding, uint8_t *converged, int32_t *iterations, double *llr,
                               double *soft_out) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (batch < 0 || (batch > 0 && (!soft_syndromes || !decoding))) {
        h->err = "bad decode arguments";
        return BPB_ERR_ARG;
    }
    if (!(sigma > 0)) {
        h->err = "The sigma value must be a float greater than 0.";  // _bp_decoder.pyx:748-749
        return BPB_ERR_ARG;
    }
    if (!h->children.empty()) {
        h->err = "bpb_soft_info_decode_batch is not split over devices; use a single-device handle";
        return BPB_ERR_UNSUPPORTED;
    }
    if (batch == 0) return BPB_OK;
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    const bpb::HostGraph &g = h->g;
    cudaStream_t st = h->stream;
    if (h->order_dirty) {
        BPB_CUDA(h, cudaStreamSynchronize(st));
        BPB_CUDA(h, cudaMemcpy(h->rel_order.ptr, h->serial_order.data(), h->serial_order.size() * sizeof(uint32_t),
                               cudaMemcpyHostToDevice));
        h->order_dirty = false;
    }
    const size_t row_bytes = (size_t) g.m * 16 + (size_t) g.n * 9 + 5;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, ((int64_t) 1 << 30) / (int64_t) row_bytes));
    const size_t cap = (size_t) chunk;
    if ((rc = ensure(h, h->soft_in, cap * g.m * 8))) return rc;
    if (soft_out && (rc = ensure(h, h->soft_out, cap * g.m * 8))) return rc;
    if (llr && (rc = ensure(h, h->soft_llr, cap * g.n * 8))) return rc;
    if ((rc = ensure(h, h->st_dec[0], cap * g.n))) return rc;
    if ((rc = ensure(h, h->st_conv[0], cap))) return rc;
    if ((rc = ensure(h, h->st_iters[0], cap * 4))) return rc;
    if ((rc = ensure(h, h->counter_s[0], 64))) return rc;
    for (int64_t lo = 0; lo < batch; lo += chunk) {
        const int64_t nb = std::min(chunk, batch - lo);
        BPB_CUDA(h, cudaMemcpyAsync(h->soft_in.ptr, soft_syndromes + lo * g.m, (size_t) nb * g.m * 8,
                                    cudaMemcpyHostToDevice, st));
        BPB_CUDA(h, cudaMemsetAsync(h->counter_s[0].ptr, 0, 64, st));
        const int e = bpb::launch_softinfo_kernel(g, h->sm_count, h->max_smem_optin, (const uint32_t *) h->blob.ptr,
                                                  h->prior_off, h->max_iter, h->ms_scaling, cutoff, sigma,
                                                  (const uint32_t *) h->rel_order.ptr, (int) h->serial_order.size(),
                                                  (const double *) h->soft_in.ptr, nb,
                                                  (unsigned long long *) h->counter_s[0].ptr + 2, &h->rel_msg,
                                                  (uint8_t *) h->st_dec[0].ptr, (uint8_t *) h->st_conv[0].ptr,
                                                  (int32_t *) h->st_iters[0].ptr,
                                                  llr ? (double *) h->soft_llr.ptr : nullptr,
                                                  soft_out ? (double *) h->soft_out.ptr : nullptr, st);
        if (e == -1) {
            h->err = "soft-information decoding: the code does not fit the kernel (column degree > 32 or n, m beyond shared memory)";
            return BPB_ERR_UNSUPPORTED;
        }
        if (e) {
            h->err = std::string("bp_softinfo_kernel launch: ") + cudaGetErrorString((cudaError_t) e);
            return BPB_ERR_CUDA;
        }
        h->launches += 1;
        h->last_family = BPB_KERNEL_EDGE;
        BPB_CUDA(h, cudaMemcpyAsync(decoding + lo * g.n, h->st_dec[0].ptr, (size_t) nb * g.n, cudaMemcpyDeviceToHost, st));
        if (converged)
            BPB_CUDA(h, cudaMemcpyAsync(converged + lo, h->st_conv[0].ptr, (size_t) nb, cudaMemcpyDeviceToHost, st));
        if (iterations)
            BPB_CUDA(h, cudaMemcpyAsync(iterations + lo, h->st_iters[0].ptr, (size_t) nb * 4, cudaMemcpyDeviceToHost, st));
        if (llr)
            BPB_CUDA(h, cudaMemcpyAsync(llr + lo * g.n, h->soft_llr.ptr, (size_t) nb * g.n * 8, cudaMemcpyDeviceToHost, st));
        if (soft_out)
            BPB_CUDA(h, cudaMemcpyAsync(soft_out + lo * g.m, h->soft_out.ptr, (size_t) nb * g.m * 8, cudaMemcpyDeviceToHost, st));
        BPB_CUDA(h, cudaStreamSynchronize(st));
    }
    h->have_last = false;
    return BPB_OK;
}

int bpb_mc_bsc(bpb_decoder *h, uint64_t seed, int64_t first_run, int64_t runs, const double *flip_prob, int with_osd,
               int64_t counts[5]) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (runs < 0 || first_run < 0 || !counts) {
        h->err = "bad Monte-Carlo arguments";
        return BPB_ERR_ARG;
    }
    for (int i = 0; i < 5; i++) counts[i] = 0;
    if (runs == 0) return BPB_OK;
    const bpb::HostGraph &g = h->g;
    if (!h->children.empty()) {
        // runs split contiguously over the devices; run numbers (hence the drawn errors) do not depend on the split
        const int k = (int) h->children.size();
        std::vector<int> rcs((size_t) k, BPB_OK);
        std::vector<std::vector<int64_t>> part((size_t) k, std::vector<int64_t>(5, 0));
        auto work = [&](int r) {
            const int64_t lo = runs * r / k, hi = runs * (r + 1) / k;
            if (hi > lo)
                rcs[(size_t) r] = bpb_mc_bsc(h->children[(size_t) r], seed, first_run + lo, hi - lo, flip_prob, with_osd,
                                             part[(size_t) r].data());
        };
        std::vector<std::thread> pool;
        for (int r = 1; r < k; r++) pool.emplace_back(work, r);
        work(0);
        for (auto &t: pool) t.join();
        for (int r = 0; r < k; r++) {
            if (rcs[(size_t) r]) {
                h->err = "device slice " + std::to_string(r) + ": " + h->children[(size_t) r]->err;
                return rcs[(size_t) r];
            }
            for (int i = 0; i < 5; i++) counts[i] += part[(size_t) r][(size_t) i];
        }
        return BPB_OK;
    }
    BPB_CUDA(h, cudaSetDevice(h->device));
    if (h->graph_dirty && (rc = upload_graph(h))) return rc;
    if (with_osd && !osd_on_device(h)) {
        h->err = "bpb_mc_bsc with OSD-0 needs the device OSD-0 kernel, which is not available for this code";
        return BPB_ERR_UNSUPPORTED;
    }
    std::vector<unsigned long long> thresh((size_t) g.n);
    for (int j = 0; j < g.n; j++) {
        double p = flip_prob ? flip_prob[j] : h->channel[(size_t) j];
        if (!(p >= 0.0)) p = 0.0;
        if (p > 1.0) p = 1.0;
        thresh[(size_t) j] = (unsigned long long) (p * 4294967296.0);
    }
    cudaStream_t st = h->stream;
    if ((rc = ensure(h, h->mc_thresh, (size_t) g.n * 8))) return rc;
    BPB_CUDA(h, cudaMemcpyAsync(h->mc_thresh.ptr, thresh.data(), (size_t) g.n * 8, cudaMemcpyHostToDevice, st));
    const int nw = (g.n + 31) / 32, mwp = round_up((g.m + 31) / 32, 4);
    const size_t row_bytes = (size_t) g.n + (size_t) nw * 4 + (size_t) mwp * 4 + 5 + (with_osd ? (size_t) g.n * 8 : 0);
    const int64_t chunk = chunk_rows((int64_t) 1 << 20, row_bytes, runs);
    if ((rc = ensure(h, h->mc_err, (size_t) chunk * nw * 4))) return rc;
    if ((rc = ensure(h, h->mc_syn, (size_t) chunk * mwp * 4))) return rc;
    if ((rc = ensure(h, h->mc_dec, (size_t) chunk * g.n))) return rc;
    if ((rc = ensure(h, h->mc_conv, (size_t) chunk))) return rc;
    if ((rc = ensure(h, h->mc_its, (size_t) chunk * 4))) return rc;
    if ((rc = ensure(h, h->mc_counts, 64))) return rc;
    BPB_CUDA(h, cudaMemsetAsync(h->mc_counts.ptr, 0, 64, st));
    const uint32_t *blob = (const uint32_t *) h->blob.ptr;
    for (int64_t lo = 0; lo < runs; lo += chunk) {
        const int64_t nb = std::min(chunk, runs - lo);
        int e = bpb::launch_mc_generate(blob, blob + (g.m + 1), (const unsigned long long *) h->mc_thresh.ptr, g.m, g.n,
                                        nw, mwp, seed, (unsigned long long) (first_run + lo), nb,
                                        (uint32_t *) h->mc_err.ptr, (uint32_t *) h->mc_syn.ptr, h->sm_count, st);
        if (e) {
            h->err = std::string("mc_generate_kernel: ") + cudaGetErrorString((cudaError_t) e);
            return BPB_ERR_CUDA;
        }
        if (with_osd)
            rc = enqueue_bposd_device(h, (const uint8_t *) h->mc_syn.ptr, nb, (uint8_t *) h->mc_dec.ptr,
                                      (uint8_t *) h->mc_conv.ptr, (int32_t *) h->mc_its.ptr, nullptr, st, kInputPacked);
        else
            rc = decode_device_core(h, kInputPacked, (const uint8_t *) h->mc_syn.ptr, nb, (uint8_t *) h->mc_dec.ptr,
                                    (uint8_t *) h->mc_conv.ptr, (int32_t *) h->mc_its.ptr, nullptr, st);
        if (rc) return rc;
        e = bpb::launch_mc_score((const uint8_t *) h->mc_dec.ptr, (const uint32_t *) h->mc_err.ptr,
                                 (const uint8_t *) h->mc_conv.ptr, (const int32_t *) h->mc_its.ptr, g.n, nw, nb,
                                 (unsigned long long *) h->mc_counts.ptr, h->sm_count, st);
        if (e) {
            h->err = std::string("mc_score_kernel: ") + cudaGetErrorString((cudaError_t) e);
            return BPB_ERR_CUDA;
        }
        h->launches += 2;
    }
    unsigned long long host[5] = {0, 0, 0, 0, 0};
    BPB_CUDA(h, cudaMemcpyAsync(host, h->mc_counts.ptr, sizeof(host), cudaMemcpyDeviceToHost, st));
    BPB_CUDA(h, cudaStreamSynchronize(st));
    counts[0] = (int64_t) host[4];
    counts[1] = (int64_t) host[0];
    counts[2] = (int64_t) host[1];
    counts[3] = (int64_t) host[2];
    counts[4] = (int64_t) host[3];
    return BPB_OK;
}

int bpb_set_devices(bpb_decoder *h, const int *ids, int count) {
    if (!h) return BPB_ERR_ARG;
    if (h->device < 0) {
        h->err = "host-only handle";
        return BPB_ERR_CUDA;
    }
    if (count < 0 || (count > 0 && !ids)) {
        h->err = "bad device list";
        return BPB_ERR_ARG;
    }
    for (bpb_decoder *c: h->children) bpb_destroy(c);
    h->children.clear();
    if (count == 0) return BPB_OK;
    const bpb::HostGraph &g = h->g;
    std::vector<int32_t> rows((size_t) g.nnz), cols((size_t) g.nnz);
    for (int i = 0; i < g.m; i++)
        for (uint32_t e = g.row_ptr[(size_t) i]; e < g.row_ptr[(size_t) i + 1]; e++) {
            rows[e] = i;
            cols[e] = (int32_t) g.col_idx[e];
        }
    for (int k = 0; k < count; k++) {
        bpb_decoder *c = nullptr;
        int rc = bpb_create(g.m, g.n, g.nnz, rows.data(), cols.data(), ids[k], &c);
        if (rc == BPB_OK && !h->channel.empty()) rc = bpb_set_channel(c, h->channel.data(), g.n);
        if (rc == BPB_OK) {
            c->max_iter = h->max_iter;
            c->method = h->method;
            c->schedule = h->schedule;
            c->ms_scaling = h->ms_scaling;
            c->serial_order = h->serial_order;
            c->kernel_pref = h->kernel_pref;
            c->osd_location = h->osd_location;
            c->obs_k = h->obs_k;
            c->obs_ptr = h->obs_ptr;
            c->obs_col = h->obs_col;
            c->obs_dirty = h->obs_k > 0;
            c->graph_dirty = true;
            h->children.push_back(c);
        } else {
            h->err = std::string("bpb_set_devices: device ") + std::to_string(ids[k]) + ": " +
                     (c ? c->err : std::string(bpb_last_error(nullptr)));
            if (c) bpb_destroy(c);
            for (bpb_decoder *d: h->children) bpb_destroy(d);
            h->children.clear();
            return rc;
        }
    }
    return BPB_OK;
}

int bpb_set_osd_location(bpb_decoder *h, int v) {
    if (!h) return BPB_ERR_ARG;
    if (v != BPB_OSD_AUTO && v != BPB_OSD_HOST && v != BPB_OSD_DEVICE) {
        h->err = "invalid OSD location";
        return BPB_ERR_ARG;
    }
    h->osd_location = v;
    for (bpb_decoder *c: h->children) c->osd_location = v;
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
        build_pair_plan(h);
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
    if (h->device >= 0 && h->counter_s[h->slot].ptr && h->last_family == BPB_KERNEL_STREAM) {
        unsigned long long words[4] = {0, 0, 0, 0};
        if (cudaMemcpy(words, h->counter_s[h->slot].ptr, sizeof(words), cudaMemcpyDeviceToHost) == cudaSuccess) {
            out->stream_handed_off = (int64_t) words[1];
            out->stream_iterations = (int64_t) words[3];
        }
    }
    out->osd_device_available = (h->device >= 0 && !h->graph_dirty) ? (h->osd_plan.warps_per_cta > 0)
                                                                    : (bpb::plan_osd_device(h->g, h->max_smem_optin > 0 ? h->max_smem_optin : 232448).warps_per_cta > 0);
    out->osd_device_solved = h->osd_device_solved;
    out->osd_host_solved = h->osd_host_solved;
    out->osd_host_inconsistent = h->osd_host_inconsistent;
    for (const bpb_decoder *c: h->children) {
        out->osd_device_solved += c->osd_device_solved;
        out->osd_host_solved += c->osd_host_solved;
        out->launches += c->launches;
    }
    out->smem_family_available = h->smem_plan.ok ? 1 : 0;
    out->smem_bank_multiplicity = h->smem_plan.max_bank_multiplicity;
    out->smem_bytes_per_syndrome = (int) h->smem_plan.group_bytes;
    out->pair_family_available = h->pair_plan.ok ? 1 : 0;
    out->pair_bank_multiplicity = h->pair_plan.max_bank_multiplicity;
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

namespace {

// bpb_set_devices: contiguous slices of one host batch, one host thread per device, each through that device's own
// pipeline; outputs land in disjoint ranges of the caller's arrays.
int multi_device_decode(bpb_decoder *h, int input_type, const uint8_t *input, int64_t batch, uint8_t *decoding,
                        uint8_t *converged, int32_t *iterations, double *llr, uint8_t *bp_decoding, bool osd,
                        int threads) {
    const int k = (int) h->children.size();
    const bpb::HostGraph &g = h->g;
    const int in_w = (input_type == BPB_INPUT_RECEIVED_VECTOR) ? g.n : g.m;
    std::vector<int> rcs((size_t) k, BPB_OK);
    auto work = [&](int r) {
        const int64_t lo = batch * r / k, hi = batch * (r + 1) / k;
        if (hi <= lo) return;
        bpb_decoder *c = h->children[(size_t) r];
        const int64_t nb = hi - lo;
        if (osd)
            rcs[(size_t) r] = bpb_bposd_decode_batch(c, input + lo * in_w, nb, decoding + lo * g.n,
                                                     converged ? converged + lo : nullptr,
                                                     iterations ? iterations + lo : nullptr,
                                                     bp_decoding ? bp_decoding + lo * g.n : nullptr,
                                                     threads > 0 ? std::max(1, threads / k) : 0);
        else
            rcs[(size_t) r] = bpb_decode_batch(c, input_type, input + lo * in_w, nb, decoding + lo * g.n,
                                               converged ? converged + lo : nullptr,
                                               iterations ? iterations + lo : nullptr, llr ? llr + lo * g.n : nullptr);
    };
    std::vector<std::thread> pool;
    for (int r = 1; r < k; r++) pool.emplace_back(work, r);
    work(0);
    for (auto &t: pool) t.join();
    for (int r = 0; r < k; r++)
        if (rcs[(size_t) r]) {
            h->err = "device slice " + std::to_string(r) + ": " + h->children[(size_t) r]->err;
            return rcs[(size_t) r];
        }
    h->last_family = h->children[0]->last_family;
    h->last_grid = h->children[0]->last_grid;
    h->last_block = h->children[0]->last_block;
    return BPB_OK;
}

}  // namespace
