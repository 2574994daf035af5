// mc_device.cu -- on-device syndrome generation and logical-failure reduction.
//
// The reference's Monte-Carlo driver draws a BSC error, computes its syndrome, decodes it and compares the decoding
// with the error, one run at a time on the host (reference src_python/ldpc/monte_carlo_simulation/mcs.py:124-139 with
// generate_bsc_error, src_python/ldpc/noise_models/bsc.py:23, and GF2Sparse::mulvec, src_cpp/gf2sparse.hpp:177-214).
// Here the three host steps around the decoder become two kernels, so an error-rate sweep moves no per-syndrome data
// over PCIe: only five counters come back.
//
//   mc_generate_kernel : run r draws e_j ~ Bernoulli(p_j) from Philox4x32-10 keyed by the seed with counter
//                        (r, j / 4) -- bit j uses output word j % 4, so the error of run r depends on (seed, r, j)
//                        only: independent of chunking, batch size and of how runs are sharded over devices -- then
//                        s = H e.  One warp per run; outputs are bit-packed (errors [B][nw], syndromes [B][mwp]);
//   mc_score_kernel    : compares decoding[b] with e_b (the reference's `not np.array_equal(decoding, error)`,
//                        mcs.py:135) and reduces {failures, converged, sum of iterations, converged-but-wrong}.
#include <algorithm>

#include "bp_decoder.h"

namespace bpb {

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

}  // namespace

// thresh[j] = floor(p_j * 2^32) as u64 (2^32 for p_j = 1): bit j is flipped iff its 32-bit draw is below it.
__global__ void __launch_bounds__(256) mc_generate_kernel(const uint32_t *__restrict__ row_ptr,
                                                          const uint32_t *__restrict__ col_idx,
                                                          const unsigned long long *__restrict__ thresh, int m, int n,
                                                          int nw, int mwp, unsigned long long seed,
                                                          unsigned long long first_run, long long batch,
                                                          uint32_t *__restrict__ err_packed,
                                                          uint32_t *__restrict__ synd_packed) {
    extern __shared__ uint32_t mc_sm[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    uint32_t *ew = mc_sm + (size_t) wib * nw;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    const uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
    for (long long b = warp; b < batch; b += nwarps) {
        const unsigned long long run = first_run + (unsigned long long) b;
        for (int w = 0; w < nw; ++w) {
            const int j = w * 32 + lane;
            uint32_t r[4];
            philox4x32_10((uint32_t) run, (uint32_t) (run >> 32), (uint32_t) (j >> 2), 0u, k0, k1, r);
            const bool flip = j < n && (unsigned long long) r[j & 3] < __ldg(thresh + j);
            const uint32_t word = __ballot_sync(0xffffffffu, flip);
            if (lane == 0) {
                ew[w] = word;
                err_packed[b * nw + w] = word;
            }
        }
        __syncwarp();
        for (int w = 0; w < mwp; ++w) {
            const int i = w * 32 + lane;
            uint32_t bit = 0;
            if (i < m)
                for (uint32_t e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
                    const uint32_t c = col_idx[e];
                    bit ^= (ew[c >> 5] >> (c & 31)) & 1u;
                }
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == 0) synd_packed[b * mwp + w] = word;
        }
        __syncwarp();
    }
}

// counts: [0] runs whose decoding differs from the error, [1] BP converged, [2] sum of BP iterations,
//         [3] converged but wrong (undetected), [4] runs scored
__global__ void __launch_bounds__(256) mc_score_kernel(const uint8_t *__restrict__ dec,
                                                       const uint32_t *__restrict__ err_packed,
                                                       const uint8_t *__restrict__ conv,
                                                       const int32_t *__restrict__ iters, int n, int nw,
                                                       long long batch, unsigned long long *counts) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long) gridDim.x * blockDim.x) >> 5;
    unsigned long long fails = 0, convs = 0, its = 0, wrong = 0, runs = 0;
    for (long long b = warp; b < batch; b += nwarps) {
        const uint8_t *row = dec + b * n;
        bool diff = false;
        for (int w = 0; w < nw; ++w) {
            const int j = w * 32 + lane;
            const uint32_t bit = (j < n) ? (row[j] != 0) : 0u;
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            diff |= word != err_packed[b * nw + w];
        }
        if (lane == 0) {
            const bool cv = conv[b] != 0;
            fails += diff;
            convs += cv;
            its += (unsigned long long) iters[b];
            wrong += (cv && diff);
            runs += 1;
        }
    }
    if (lane == 0 && runs) {
        atomicAdd(counts + 0, fails);
        atomicAdd(counts + 1, convs);
        atomicAdd(counts + 2, its);
        atomicAdd(counts + 3, wrong);
        atomicAdd(counts + 4, runs);
    }
}

int launch_mc_generate(const uint32_t *d_row_ptr, const uint32_t *d_col_idx, const unsigned long long *d_thresh, int m,
                       int n, int nw, int mwp, unsigned long long seed, unsigned long long first_run, int64_t batch,
                       uint32_t *d_err, uint32_t *d_syn, int sm_count, cudaStream_t st) {
    const int block = 256;
    const size_t smem = (size_t) (block / 32) * nw * 4;
    const int grid = (int) std::min<int64_t>((batch + 7) / 8, (int64_t) sm_count * 8);
    mc_generate_kernel<<<grid < 1 ? 1 : grid, block, smem, st>>>(d_row_ptr, d_col_idx, d_thresh, m, n, nw, mwp, seed,
                                                                first_run, batch, d_err, d_syn);
    return (int) cudaGetLastError();
}

int launch_mc_score(const uint8_t *d_dec, const uint32_t *d_err, const uint8_t *d_conv, const int32_t *d_iters, int n,
                    int nw, int64_t batch, unsigned long long *d_counts, int sm_count, cudaStream_t st) {
    const int grid = (int) std::min<int64_t>((batch + 7) / 8, (int64_t) sm_count * 8);
    mc_score_kernel<<<grid < 1 ? 1 : grid, 256, 0, st>>>(d_dec, d_err, d_conv, d_iters, n, nw, batch, d_counts);
    return (int) cudaGetLastError();
}

}  // namespace bpb
