// ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// A thin extern "C" batch driver around the UNMODIFIED reference C++
// (ldpc::bp::BpDecoder, ldpc::osd::OsdDecoder), compiled from the headers where
// they lie under /root/reference/src_cpp by oracle/Makefile into
// oracle/_ref/libref_bp.so (git-ignored, shipped to the GPU box as a binary).
// No reference source is copied: this file only #includes the reference headers.
//
// Used (a) to pin oracle/bp_oracle.c bit-for-bit, (b) to generate tests/golden/*.npz,
// (c) as bench.py's `cpu_baseline` / `--impl reference` arm (kind = "reference").
//
// One decoder object per worker thread (the reference decoder is stateful:
// messages live inside the matrix nodes, bp.hpp:42-49), each worker decodes a
// contiguous slice of the batch by calling BpDecoder::decode once per syndrome,
// which is what the reference's Python layer does (_bp_decoder.pyx:682).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "bp.hpp"
#include "osd.hpp"

using ldpc::bp::BpDecoder;
using ldpc::bp::BpSparse;

namespace {

struct Worker {
    std::unique_ptr<BpSparse> pcm;
    std::unique_ptr<BpDecoder> bpd;
    std::unique_ptr<ldpc::osd::OsdDecoder> osd;
};

void build_worker(Worker &w, int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols,
                  const double *channel, int max_iter, int method, int schedule, double ms, const int32_t *order,
                  int order_len, int osd_method, int osd_order) {
    w.pcm = std::make_unique<BpSparse>(m, n, (int) nnz);
    for (int64_t k = 0; k < nnz; k++) w.pcm->insert_entry(rows[k], cols[k]);
    std::vector<double> ch(channel, channel + n);
    std::vector<int> ord;
    if (order) ord.assign(order, order + order_len);
    w.bpd = std::make_unique<BpDecoder>(*w.pcm, ch, max_iter, (ldpc::bp::BpMethod) method,
                                        (ldpc::bp::BpSchedule) schedule, ms, 1,
                                        order ? ord : ldpc::bp::NULL_INT_VECTOR, 0, false, ldpc::bp::SYNDROME);
    if (osd_method > 0) {
        w.osd = std::make_unique<ldpc::osd::OsdDecoder>(*w.pcm, (ldpc::osd::OsdMethod) osd_method, osd_order,
                                                        w.bpd->channel_probabilities);
    }
}

}  // namespace

extern "C" {

// Returns seconds spent in the decode loop (max over workers), <0 on error.
// osd_method: 0 = BP only; 1 = OSD_0, 2 = EXHAUSTIVE, 3 = COMBINATION_SWEEP (osd.hpp:18-23); when >0 the
// post-processor runs for non-converged syndromes only (_bposd_decoder.pyx:128-134) and `decoding` receives
// the BP+OSD output while `bp_decoding` (optional) receives the raw BP output.
double ref_decode_batch(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, const double *channel,
                        int max_iter, int method, int schedule, double ms_scaling_factor, const int32_t *serial_order,
                        int serial_order_len, int osd_method, int osd_order, const uint8_t *syndromes, int64_t batch,
                        uint8_t *out_decoding, uint8_t *out_converged, int32_t *out_iters, double *out_llr,
                        uint8_t *out_bp_decoding, int threads) {
    if (threads < 1) threads = 1;
    if ((int64_t) threads > batch) threads = (int) (batch > 0 ? batch : 1);
    std::vector<Worker> workers((size_t) threads);
    try {
        for (auto &w: workers)
            build_worker(w, m, n, nnz, rows, cols, channel, max_iter, method, schedule, ms_scaling_factor,
                         serial_order, serial_order_len, osd_method, osd_order);
    } catch (...) {
        return -1.0;
    }
    std::vector<double> secs((size_t) threads, 0.0);
    auto run = [&](int t) {
        Worker &w = workers[(size_t) t];
        int64_t lo = batch * t / threads, hi = batch * (t + 1) / threads;
        std::vector<uint8_t> syn((size_t) m);
        // SERIAL_RELATIVE (bp.hpp:469-482) re-sorts the public member serial_schedule_order in place and the object
        // carries the result into its next decode; the batch semantics pinned here are "every syndrome as by a freshly
        // constructed decoder", so the member is put back before each call (a write to a public member, like the
        // reference's own Python shim does, _bp_decoder.pyx:471-475)
        const std::vector<int> initial_order = w.bpd->serial_schedule_order;
        auto t0 = std::chrono::steady_clock::now();
        for (int64_t b = lo; b < hi; b++) {
            std::memcpy(syn.data(), syndromes + b * (int64_t) m, (size_t) m);
            if (schedule == 2) w.bpd->serial_schedule_order = initial_order;
            w.bpd->decode(syn);
            const std::vector<uint8_t> *out = &w.bpd->decoding;
            if (out_bp_decoding) std::memcpy(out_bp_decoding + b * (int64_t) n, w.bpd->decoding.data(), (size_t) n);
            if (w.osd && !w.bpd->converge) {
                w.osd->decode(syn, w.bpd->log_prob_ratios);
                out = &w.osd->osdw_decoding;
            }
            if (out_decoding) std::memcpy(out_decoding + b * (int64_t) n, out->data(), (size_t) n);
            if (out_converged) out_converged[b] = w.bpd->converge ? 1 : 0;
            if (out_iters) out_iters[b] = w.bpd->iterations;
            if (out_llr)
                std::memcpy(out_llr + b * (int64_t) n, w.bpd->log_prob_ratios.data(), sizeof(double) * (size_t) n);
        }
        secs[(size_t) t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    if (threads == 1) {
        run(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(run, t);
        for (auto &th: pool) th.join();
    }
    double mx = 0;
    for (double s: secs) mx = s > mx ? s : mx;
    return mx;
}

// Received-vector input (bp.hpp:162-180): decode(v) with bp_input_type RECEIVED_VECTOR.
double ref_decode_received(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, const double *channel,
                           int max_iter, int method, int schedule, double ms_scaling_factor, const uint8_t *vecs,
                           int64_t batch, uint8_t *out_decoding) {
    Worker w;
    try {
        build_worker(w, m, n, nnz, rows, cols, channel, max_iter, method, schedule, ms_scaling_factor, nullptr, 0, 0,
                     0);
    } catch (...) {
        return -1.0;
    }
    w.bpd->bp_input_type = ldpc::bp::RECEIVED_VECTOR;
    std::vector<uint8_t> v((size_t) n);
    for (int64_t b = 0; b < batch; b++) {
        std::memcpy(v.data(), vecs + b * (int64_t) n, (size_t) n);
        w.bpd->decode(v);
        std::memcpy(out_decoding + b * (int64_t) n, w.bpd->decoding.data(), (size_t) n);
    }
    return 0.0;
}

// ldpc::bp::BpDecoder::soft_info_decode_serial (bp.hpp:547-665), one call per soft syndrome, one decoder object.
double ref_soft_info_decode_batch(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols,
                                  const double *channel, int max_iter, double ms_scaling_factor,
                                  const int32_t *serial_order, int serial_order_len, double cutoff, double sigma,
                                  const double *soft_syndromes, int64_t batch, uint8_t *out_decoding,
                                  uint8_t *out_converged, int32_t *out_iters, double *out_llr, double *out_soft) {
    Worker w;
    try {
        build_worker(w, m, n, nnz, rows, cols, channel, max_iter, 1 /* minimum_sum */, 0 /* serial */,
                     ms_scaling_factor, serial_order, serial_order_len, 0, 0);
    } catch (...) {
        return -1.0;
    }
    std::vector<double> soft((size_t) m);
    for (int64_t b = 0; b < batch; b++) {
        std::memcpy(soft.data(), soft_syndromes + b * (int64_t) m, sizeof(double) * (size_t) m);
        w.bpd->soft_info_decode_serial(soft, cutoff, sigma);
        std::memcpy(out_decoding + b * (int64_t) n, w.bpd->decoding.data(), (size_t) n);
        if (out_converged) out_converged[b] = w.bpd->converge ? 1 : 0;
        if (out_iters) out_iters[b] = w.bpd->iterations;
        if (out_llr) std::memcpy(out_llr + b * (int64_t) n, w.bpd->log_prob_ratios.data(), sizeof(double) * (size_t) n);
        if (out_soft) std::memcpy(out_soft + b * (int64_t) m, w.bpd->soft_syndrome.data(), sizeof(double) * (size_t) m);
    }
    return 0.0;
}

int ref_hardware_threads() { return (int) std::thread::hardware_concurrency(); }

}  // extern "C"
