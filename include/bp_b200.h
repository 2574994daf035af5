/*
 * bp_b200.h -- C ABI of the B200-native batched belief-propagation decoder.
 *
 * This is the drop-in boundary for the BP / BP+OSD syndrome-decoding path of
 * quantumgizmos/ldpc (reference paths relative to /root/reference).  Today that
 * boundary is the Cython extern block src_python/ldpc/bp_decoder/_bp_decoder.pxd:9-83
 * (class ldpc::bp::BpDecoder, src_cpp/bp.hpp:51-132, driven by direct member access) and
 * src_python/ldpc/bposd_decoder/_bposd_decoder.pxd:9-29 (ldpc::osd::OsdDecoder,
 * src_cpp/osd.hpp:26-189).  A maintainer re-points those extern blocks at the functions
 * below (INTEGRATION.md shows the stub).  Plain C: opaque handle, raw pointers + sizes, no
 * exceptions, no torch / C++ types.  Every function returns BPB_OK (0) or a negative code;
 * bpb_last_error() gives the message.
 *
 * Enumerations use the reference's numeric values (bp.hpp:23-38, osd.hpp:18-23).
 *
 * Ownership: the caller owns every host/device array it passes; the library copies the
 * parity-check matrix and the channel at create/set time and owns all device workspaces.
 * Threading: one handle = one logical decoder bound to one CUDA device; calls on one handle
 * must not overlap; distinct handles are independent.
 */
#ifndef BP_B200_H
#define BP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bpb_decoder bpb_decoder;

enum { BPB_OK = 0, BPB_ERR_ARG = -1, BPB_ERR_CUDA = -2, BPB_ERR_UNSUPPORTED = -3, BPB_ERR_NOMEM = -4 };

/* ldpc::bp::BpMethod, bp.hpp:23-26 */
enum { BPB_PRODUCT_SUM = 0, BPB_MINIMUM_SUM = 1 };
/* ldpc::bp::BpSchedule, bp.hpp:28-32.  SERIAL_RELATIVE re-sorts the schedule by posterior LLR before every sweep
 * (bp.hpp:469-482); every syndrome of a batch starts from the configured serial_schedule_order. */
enum { BPB_SERIAL = 0, BPB_PARALLEL = 1, BPB_SERIAL_RELATIVE = 2 };
/* ldpc::bp::BpInputType, bp.hpp:34-38 */
enum { BPB_INPUT_SYNDROME = 0, BPB_INPUT_RECEIVED_VECTOR = 1 };
/* kernel family: AUTO picks the fastest family that supports the code (STREAM: a lane per syndrome, messages in HBM;
 * SMEM: a thread group per syndrome, messages in shared memory; PAIR: a thread group per two syndromes, double2
 * messages in shared memory; EDGE: a CTA per syndrome, a lane per edge) */
enum { BPB_KERNEL_AUTO = 0, BPB_KERNEL_STREAM = 1, BPB_KERNEL_SMEM = 2, BPB_KERNEL_EDGE = 3, BPB_KERNEL_PAIR = 4 };
/* where OSD-0 runs in the BP+OSD entry points: AUTO = on the device when the code fits (m <= 1024 and the permuted
 * bit matrix fits the shared memory of an SM), else on the host */
enum { BPB_OSD_AUTO = 0, BPB_OSD_HOST = 1, BPB_OSD_DEVICE = 2 };

/* --- life cycle ------------------------------------------------------------------------------
 * Replaces `new BpSparse(m,n,nnz)` + insert_entry per nonzero (_bp_decoder.pyx:9-49) and
 * `new BpDecoderCpp(...)` (_bp_decoder.pyx:132, bp.hpp:77-132).  (rows[k], cols[k]) are the
 * nonzeros of H in any order; they are sorted into the reference's traversal order
 * (sparse_matrix_base.hpp:423-482).  `device` is the CUDA ordinal (device < 0 creates a host-only handle on
 * which only bpb_osd0_host works; decode calls on it fail -- there is no CPU decode path).  Defaults after create are the
 * reference constructor's: product_sum, parallel, ms_scaling_factor 0.625... overwritten by the
 * shim exactly as the Cython shim does (_bp_decoder.pyx:135-155). */
int bpb_create(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, int device, bpb_decoder **out);
void bpb_destroy(bpb_decoder *h);
const char *bpb_last_error(const bpb_decoder *h); /* h may be NULL: last create error */

/* --- parameters: replace writes to public members of BpDecoder (bp.hpp:55-74) --------------- */
int bpb_set_channel(bpb_decoder *h, const double *channel_probabilities, int n); /* bp.hpp:55 */
int bpb_set_max_iter(bpb_decoder *h, int maximum_iterations);                    /* bp.hpp:58 */
int bpb_set_method(bpb_decoder *h, int bp_method);                               /* bp.hpp:59 */
int bpb_set_schedule(bpb_decoder *h, int schedule);                              /* bp.hpp:60 */
int bpb_set_ms_scaling_factor(bpb_decoder *h, double ms_scaling_factor);         /* bp.hpp:62 */
int bpb_set_serial_schedule_order(bpb_decoder *h, const int32_t *order, int len); /* bp.hpp:69; NULL = 0..n-1 */
/* SERIAL_RELATIVE: the schedule the LAST syndrome of the most recent decode call ended with -- what the reference's
 * serial_schedule_order member holds after that decode (bp.hpp:469-482 sorts it in place).  Synchronises. */
int bpb_get_last_schedule_order(bpb_decoder *h, int32_t *out, int len);
int bpb_set_kernel(bpb_decoder *h, int kernel_family);                           /* new: BPB_KERNEL_* */
int bpb_set_osd_location(bpb_decoder *h, int osd_location);                      /* new: BPB_OSD_* */
/* new: split every HOST-pointer batch call of this handle over `count` CUDA devices.  One
 * full decoder (own streams, pinned-speed staging, workspaces) is created per listed device (an ordinal may repeat);
 * the host input array is split into `count` contiguous slices, one host thread per device drives its slice through
 * the chunked H2D | kernels | D2H pipeline and the D2H copies land in disjoint ranges of the caller's output arrays:
 * that is the whole "gather", no collective is involved.  Parameter setters called afterwards are forwarded.
 * count == 0 (or ids == NULL) returns to the single device given at create time. */
int bpb_set_devices(bpb_decoder *h, const int *device_ids, int count);

/* --- decode: replaces BpDecoder::decode(vector<uint8_t>&) (bp.hpp:159-190) for a whole batch --
 * input  : [batch][m] (syndromes) or [batch][n] (received vectors), uint8 0/1, row-major
 * outputs: decoding [batch][n] uint8 (bp.hpp:63), converged [batch] uint8 (bp.hpp:72),
 *          iterations [batch] int32 (bp.hpp:70), log_prob_ratios [batch][n] double (bp.hpp:66).
 *          converged / iterations / log_prob_ratios may be NULL.
 * Every syndrome is decoded exactly as one BpDecoder::decode call would (including the per-syndrome
 * early exit, bp.hpp:300-308, 539-542); results do not depend on batch order or size.
 *
 * bpb_decode_batch takes HOST pointers (pageable or pinned) and includes the H2D/D2H copies.
 * bpb_decode_batch_device takes DEVICE pointers on the handle's device and enqueues everything on
 * `cuda_stream` (a cudaStream_t passed as void*, NULL = default stream) without synchronising; d_decoding must be
 * 4-byte aligned and d_log_prob_ratios 8-byte aligned (any cudaMalloc / framework allocation is). */
int bpb_decode_batch(bpb_decoder *h, int input_type, const uint8_t *input, int64_t batch, uint8_t *decoding,
                     uint8_t *converged, int32_t *iterations, double *log_prob_ratios);
int bpb_decode_batch_device(bpb_decoder *h, int input_type, const uint8_t *d_input, int64_t batch,
                            uint8_t *d_decoding, uint8_t *d_converged, int32_t *d_iterations,
                            double *d_log_prob_ratios, void *cuda_stream);

/* --- OSD-0 post-processing on the host (replaces OsdDecoder::decode with osd_order 0,
 * osd.hpp:110-117 -> sort.hpp:48-62 + gf2sparse_linalg.hpp:298-401).  For every b with
 * converged[b] == 0 the row decoding[b] is overwritten with the OSD-0 solution computed from
 * syndromes[b] and log_prob_ratios[b]; rows with converged[b] != 0 are left alone
 * (_bposd_decoder.pyx:128-134).  `threads` <= 0 means all hardware threads. */
int bpb_osd0_host(bpb_decoder *h, const uint8_t *syndromes, const double *log_prob_ratios,
                  const uint8_t *converged, int64_t batch, uint8_t *decoding, int threads);

/* --- BP + OSD-0 for a whole batch: replaces BpOsdDecoder.decode (_bposd_decoder.pyx:78-136: bpd.decode, then
 * osdD.decode only when !bpd.converge) for `batch` syndromes.  BP runs on the device and its posterior LLRs stay
 * there.  OSD-0 for the rows BP did not solve runs on the device too (osd_device.cu) when the code fits that kernel
 * and bpb_set_osd_location allows it (BPB_OSD_AUTO); otherwise those LLR rows are gathered, copied back and solved by
 * bpb_osd0_host's elimination on `threads` host threads.  decoding [batch][n] receives the BP output for converged rows and the OSD-0 solution otherwise;
 * bp_decoding (optional, [batch][n]) receives the raw BP output; converged / iterations are BP's (may be NULL). */
int bpb_bposd_decode_batch(bpb_decoder *h, const uint8_t *syndromes, int64_t batch, uint8_t *decoding,
                           uint8_t *converged, int32_t *iterations, uint8_t *bp_decoding, int threads);

/* The same with DEVICE pointers, enqueued on `cuda_stream` without synchronising: BP, failure list, OSD-0 (the
 * bit-packed elimination of osd_device.cu, osd.hpp:110-117) all on the device; nothing crosses PCIe.  Fails with
 * BPB_ERR_UNSUPPORTED when the code does not fit the device OSD-0 kernel (see BPB_OSD_AUTO).  d_bp_decoding may be
 * NULL. */
int bpb_bposd_decode_batch_device(bpb_decoder *h, const uint8_t *d_syndromes, int64_t batch, uint8_t *d_decoding,
                                  uint8_t *d_converged, int32_t *d_iterations, uint8_t *d_bp_decoding,
                                  void *cuda_stream);

/* --- Monte-Carlo runs on the binary symmetric channel without per-syndrome PCIe traffic: replaces the loop body of
 * MonteCarloBscSimulation.run (src_python/ldpc/monte_carlo_simulation/mcs.py:124-139: generate_bsc_error ->
 * H @ error % 2 -> Decoder.decode -> compare) for runs first_run .. first_run + runs - 1.  Run r flips bit j with
 * probability flip_prob[j] (NULL: the decoder's channel probabilities) using Philox4x32-10 keyed by `seed` with
 * counter (r, j / 4), so the drawn errors depend on (seed, r, j) only -- not on batch size, chunking or how the runs
 * are sharded over devices.  with_osd != 0 adds OSD-0 for the BP failures (needs the device OSD-0 kernel).
 * counts[0] = runs scored, [1] = runs whose decoding differs from the drawn error (the reference's fail_count),
 * [2] = runs BP converged on, [3] = sum of BP iterations, [4] = runs that converged to a wrong decoding. */
int bpb_mc_bsc(bpb_decoder *h, uint64_t seed, int64_t first_run, int64_t runs, const double *flip_prob, int with_osd,
               int64_t counts[5]);

/* --- bit-packed I/O (new) -------------------------------------------------------
 * stim's b8 layout, which is what sinter hands to SinterBpOsdDecoder.decode_via_files
 * (sinter_decoders/sinter_bposd_decoder.py:57-126): a row is ceil(bits / 8) bytes, bit k of the row is bit k % 8 of
 * byte k / 8.  The reference unpacks every shot to a uint8 vector, decodes it, multiplies the correction by the
 * observables matrix on the host and packs the result (:114-130).  Here the packed rows cross PCIe as they are
 * (ceil(m/8) bytes in, ceil(n/8) and / or ceil(k/8) bytes out per shot instead of m + n), unpacking, BP (+ OSD-0 on the
 * device when with_osd != 0), packing and the observable parities  obs = O x mod 2  all run on the device.
 * bpb_set_observables: O is k x n, given as nonzero coordinates.  decoding_b8 / observables_b8 / converged / iterations
 * may be NULL (at least one of the first two is required). */
int bpb_set_observables(bpb_decoder *h, int k, int64_t nnz, const int32_t *rows, const int32_t *cols);
int bpb_decode_batch_b8(bpb_decoder *h, int with_osd, const uint8_t *syndromes_b8, int64_t batch, uint8_t *decoding_b8,
                        uint8_t *observables_b8, uint8_t *converged, int32_t *iterations);

/* --- soft-information decoding (new) -----------------------------------------------
 * Replaces BpDecoder::soft_info_decode_serial(soft_syndrome, cutoff, sigma) (bp.hpp:547-665), which the reference's
 * SoftInfoBpDecoder.decode calls once per soft syndrome (_bp_decoder.pyx:761-785): serial-schedule min-sum (in the
 * configured serial_schedule_order, ms_scaling_factor as set, maximum_iterations as set) where checks whose soft
 * magnitude 2 s_i / sigma^2 is below `cutoff` take part as virtual variable nodes.  soft_syndromes is [B][m] doubles;
 * llr ([B][n]) and soft_out ([B][m], the reference's soft_syndrome member after the decode) may be NULL. */
int bpb_soft_info_decode_batch(bpb_decoder *h, const double *soft_syndromes, int64_t batch, double cutoff, double sigma,
                               uint8_t *decoding, uint8_t *converged, int32_t *iterations, double *llr,
                               double *soft_out);

/* --- introspection ----------------------------------------------------------------------------- */
typedef struct {
    int m, n;
    int64_t nnz;
    int max_row_degree, max_col_degree;
    int device, sm_count;
    int kernel_family;        /* family used by the last decode */
    int grid, block;          /* launch shape of the last message-update kernel */
    int64_t launches;         /* kernels launched by this handle so far */
    int64_t workspace_bytes;  /* device bytes held */
    double last_kernel_ms;    /* CUDA-event time of the last message-update kernel, 0 until it has finished */
    int smem_family_available;    /* 1 when the on-chip kernel family can serve this code (parallel schedule) */
    int smem_bank_multiplicity;   /* worst lanes-per-bank of a half-warp message access in that family (1 = none) */
    int smem_bytes_per_syndrome;  /* shared memory held per in-flight syndrome in that family */
    int64_t stream_iterations;    /* iterations executed by the last streaming-kernel launch (synchronises) */
    int64_t stream_handed_off;    /* syndromes its ramp-down handed to the second-stage kernel */
    int osd_device_available;     /* 1 when OSD-0 for this code can run on the device */
    int64_t osd_device_solved;    /* syndromes solved by the device OSD-0 kernel so far (host API calls only) */
    int64_t osd_host_solved;      /* syndromes solved by the host elimination so far */
    int64_t osd_host_inconsistent; /* of those (and of bpb_osd0_host calls): syndromes outside the image of H, for
                                      which the result is defined here but differs from the reference's (osd_host.cpp) */
    int pair_family_available;    /* 1 when the paired on-chip family (two syndromes per thread group) can serve it */
    int pair_bank_multiplicity;   /* worst lanes-per-bank-quad of a quarter-warp 16-byte message access (1 = none) */
} bpb_info;
int bpb_get_info(const bpb_decoder *h, bpb_info *out);

/* Pinned host memory helpers for callers that want full-speed H2D/D2H (optional). */
void *bpb_host_alloc(size_t bytes);
void bpb_host_free(void *p);

const char *bpb_version(void);

/* Product-sum bit-exactness depends on the host the REFERENCE runs on: the device evaluates std::tanh / std::log
 * (bp.hpp:208,216) the way glibc 2.39's x86-64 FMA variants do (ldpc_b200/csrc/ref_libm.h).  This host-side check
 * evaluates that restatement over `samples` deterministic arguments and returns how many results differ from the live
 * libm: 0 on such a host; non-zero means product-sum LLRs agree with a reference run on THIS host to ~1 ulp per
 * operation (well inside 1e-5) rather than bit for bit.  No device work. */
int bpb_libm_selfcheck(int samples);

#ifdef __cplusplus
}
#endif
#endif /* BP_B200_H */
