// bp_decoder.h -- internal state behind the opaque bpb_decoder handle (include/bp_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bp_b200.h"

namespace bpb {

// Host copy of H in the reference's traversal order (sparse_matrix_base.hpp:423-482):
// CSR with ascending columns per row, CSC with ascending rows per column.
struct HostGraph {
    int m = 0, n = 0, nnz = 0;
    std::vector<uint32_t> row_ptr, col_idx;            // CSR
    std::vector<uint32_t> col_ptr, row_idx, csc2csr;   // CSC; csc2csr: CSC position -> CSR edge id
    int max_row_degree = 0, max_col_degree = 0;
    bool regular = false;  // every row has max_row_degree entries and every column max_col_degree
};

int build_host_graph(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, HostGraph &g,
                     std::string &err);

// Bit-packed OSD-0 on the host (osd_host.cpp).
int osd0_host(const HostGraph &g, const uint8_t *syndromes, const double *llr, const uint8_t *converged,
              int64_t batch, uint8_t *decoding, int threads, int64_t *inconsistent = nullptr);

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

// Shared-memory plan of the on-chip kernel family (bp_smem.cuh): table blob + per-group areas.
struct SmemPlan {
    bool ok = false;
    std::string why;            // why the family is not usable for this code (when !ok)
    std::vector<uint8_t> blob;  // tables in their final shared-memory byte layout
    uint32_t off_row_deg = 0, off_col_deg = 0, off_col_row = 0, off_row_pos = 0, off_col_pos = 0, off_prior = 0;
    bool serial = false;        // built for the serial schedule (extra tables: self slots, levels)
    uint32_t off_col_self = 0, off_lev_ptr = 0, off_lev_bits = 0;
    int n_levels = 0, mean_level = 0;
    int max_bank_multiplicity = 0;  // 1 = both passes conflict-free (verified by build_smem_plan)
    int msg_doubles = 0;        // length of one group's message array (16 * largest colour class)
    uint32_t group_bytes = 0, goff_msg = 0, goff_dec = 0, goff_syn = 0, goff_ctl = 0;
    int M = 0, N = 0;
};

// Shared-memory plan of the paired on-chip kernel family (bp_pair.cuh): two syndromes per thread group, double2 slots.
struct PairPlan {
    bool ok = false;
    std::string why;
    std::vector<uint8_t> blob;
    uint32_t off_row_deg = 0, off_col_deg = 0, off_col_row = 0, off_row_pos = 0, off_col_pos = 0, off_prior = 0;
    int max_bank_multiplicity = 0;  // 1 = both passes conflict-free (lanes of a quarter-warp on different bank quads)
    int msg_slots = 0;              // length of one group's double2 message array (8 * largest colour class)
    uint32_t group_bytes = 0, goff_msg = 0, goff_syn = 0, goff_acc = 0, goff_ctl = 0;
    int M = 0, N = 0;
};

// Device OSD-0 (osd_device.cu): per-warp shared-memory plan; warps_per_cta == 0 when the code does not fit.
struct OsdDevicePlan {
    int warps_per_cta = 0;
    int depth = 0, ws = 0, rows_per_lane = 0;
    uint32_t warp_bytes = 0, off_a = 0, off_b = 0, off_inv = 0, off_piv = 0;
};
OsdDevicePlan plan_osd_device(const HostGraph &g, int max_smem_optin);
int launch_osd0_kernel(const OsdDevicePlan &pl, const HostGraph &g, int sm_count, const uint32_t *d_row_ptr,
                       const uint32_t *d_col_idx, const uint32_t *d_packed, int mwp, const double *d_llr,
                       const uint32_t *d_fail_idx, const unsigned long long *d_count, unsigned long long *d_counter,
                       uint8_t *d_dec, int64_t max_items, cudaStream_t st);

// SERIAL_RELATIVE schedule (bp_relative.cu): one warp per syndrome; returns a cudaError_t value, -1 = does not fit
int launch_relative_kernel(const HostGraph &g, int sm_count, int max_smem_optin, const uint32_t *d_blob,
                           uint32_t prior_off, int method, int max_iter, double ms_scaling, const uint32_t *d_order0,
                           int order_len, const uint32_t *d_packed, int mwp, int64_t batch,
                           unsigned long long *d_counter, DeviceBuffer *scratch, uint8_t *d_dec, uint8_t *d_conv,
                           int32_t *d_iters, double *d_llr, int llr_last_only, int32_t *d_order_out, cudaStream_t st,
                           int *grid_out);

// soft-information serial min-sum (bp_relative.cu): one warp per soft syndrome; cudaError_t value, -1 = does not fit
int launch_softinfo_kernel(const HostGraph &g, int sm_count, int max_smem_optin, const uint32_t *d_blob,
                           uint32_t prior_off, int max_iter, double ms_scaling, double cutoff, double sigma,
                           const uint32_t *d_order0, int order_len, const double *d_soft, int64_t batch,
                           unsigned long long *d_counter, DeviceBuffer *scratch, uint8_t *d_dec, uint8_t *d_conv,
                           int32_t *d_iters, double *d_llr, double *d_soft_out, cudaStream_t st);

// On-device BSC sampling and scoring (mc_device.cu)
int launch_mc_generate(const uint32_t *d_row_ptr, const uint32_t *d_col_idx, const unsigned long long *d_thresh, int m,
                       int n, int nw, int mwp, unsigned long long seed, unsigned long long first_run, int64_t batch,
                       uint32_t *d_err, uint32_t *d_syn, int sm_count, cudaStream_t st);
int launch_mc_score(const uint8_t *d_dec, const uint32_t *d_err, const uint8_t *d_conv, const int32_t *d_iters, int n,
                    int nw, int64_t batch, unsigned long long *d_counts, int sm_count, cudaStream_t st);

}  // namespace bpb

struct bpb_decoder;

namespace bpb {
// bp_plan.cpp (host-only planning)
void compute_priors(bpb_decoder *h);
std::vector<uint32_t> build_serial_batches(const HostGraph &g, const std::vector<uint32_t> &order, int sb);
void build_smem_plan(bpb_decoder *h);
void build_pair_plan(bpb_decoder *h);
int libm_selfcheck(int samples);
int place_messages(const HostGraph &g, int lanes_per_phase, std::vector<uint32_t> &slot_of_edge);
}  // namespace bpb

struct bpb_decoder {
    bpb::HostGraph g;
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    // parameters (reference BpDecoder members, bp.hpp:55-74)
    std::vector<double> channel;
    std::vector<double> prior;  // log((1-p)/p) computed on the host with libm like bp.hpp:150-151
    bool uniform_prior = true;
    int max_iter = 0;
    int method = BPB_PRODUCT_SUM;
    int schedule = BPB_PARALLEL;
    double ms_scaling = 0.625;
    std::vector<uint32_t> serial_order;
    std::vector<uint32_t> serial_batches;  // levelised, padded schedule the serial kernels consume
    int serial_entries = 0;                // number of schedule entries (bits + padding)
    bool serial_program_flags = false;     // the regular-code serial program carries the "already visited" flags
    int kernel_pref = BPB_KERNEL_AUTO;
    // device state
    bool graph_dirty = true;  // blob must be (re)uploaded (prior or order changed)
    bpb::DeviceBuffer blob, order_d, smem_tab;
    // work-queue counters and bit-packed syndromes: one set per pipeline slot, so that the kernels of consecutive
    // chunks of the host pipeline may overlap (on-chip families); everything else uses slot 0
    bpb::DeviceBuffer counter_s[2], packed_s[2];
    // streaming family: message tiles, ballot words, LLR tiles, hand-off list (and the edge-parallel second stage's
    // L2 scratch) -- also per slot, so that chunk c+1 ramps up on the SMs chunk c's ramp-down leaves
    bpb::DeviceBuffer msg_s[2], dec_w_s[2], syn_w_s[2], llr_tile_s[2], handoff_s[2], edge_msg_s[2];
    int slot = 0;
    bool pipeline_dual = false;  // host_pipeline is alternating two compute streams
    cudaStream_t stream2 = nullptr;
    bpb::DeviceBuffer osd_llr, osd_fail_llr, osd_fail_idx, osd_count;  // BP+OSD batch path
    bpb::OsdDevicePlan osd_plan;
    bpb::DeviceBuffer soft_in, soft_out, soft_llr;  // bpb_soft_info_decode_batch staging
    bpb::DeviceBuffer rel_order, rel_order_out, rel_msg;  // SERIAL_RELATIVE: configured / final schedule, scratch
    bool rel_order_valid = false;  // rel_order_out holds the schedule of a finished decode
    bool order_dirty = false;      // SERIAL_RELATIVE: only the configured schedule changed (cheap re-upload)
    // bit-packed I/O (bpb_decode_batch_b8): staging per pipeline slot, observables matrix (CSR by observable)
    bpb::DeviceBuffer b8_in[2], b8_words[2], b8_out[2], b8_obs[2], obs_tab;
    int obs_k = 0;
    std::vector<uint32_t> obs_ptr, obs_col;
    bool obs_dirty = false;
    bpb::DeviceBuffer mc_thresh, mc_err, mc_syn, mc_dec, mc_conv, mc_its, mc_counts;  // bpb_mc_bsc workspaces
    int osd_location = BPB_OSD_AUTO;  // where OSD-0 runs in the BP+OSD entry points
    bool llr_last_only = false;       // BP+OSD: posterior LLRs are only needed for syndromes that ran max_iter
    int64_t osd_device_solved = 0, osd_host_solved = 0, osd_host_inconsistent = 0;
    cudaEvent_t ev_last = nullptr;    // recorded after the last enqueue of bpb_decode_batch_device
    cudaStream_t last_stream = nullptr;
    bool have_last = false;
    const uint32_t *last_packed = nullptr;  // packed syndromes of the last decode (the OSD-0 kernel reads them)
    void *pin_in[2] = {nullptr, nullptr};  // pinned staging for pageable host input (host_pipeline)
    size_t pin_in_bytes[2] = {0, 0};
    unsigned long long *host_counts = nullptr;  // pinned: failure counts of the two pipeline slots
    bpb::SmemPlan smem_plan;
    bpb::PairPlan pair_plan;
    bpb::DeviceBuffer pair_tab;
    // staging for the host API
    bpb::DeviceBuffer st_in[2], st_dec[2], st_conv[2], st_iters[2], st_llr[2], st_bp[2], osd_conv;
    std::vector<bpb_decoder *> children;  // bpb_set_devices: one full decoder per device, this handle only splits
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the host API pipeline
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    uint32_t blob_words = 0, prior_off = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, kev0 = nullptr, kev1 = nullptr;
    bool kernel_timed = false;
    // info
    int last_family = 0, last_grid = 0, last_block = 0;
    int64_t launches = 0;
    double last_kernel_ms = 0.0;
    std::string err;
};
