# cython: language_level=3, boundscheck=False, wraparound=False, initializedcheck=False, cdivision=True
"""Cython binding of the C-ABI in include/bp_b200.h.

The reference's host side is a Cython shim whose ``cdef extern`` block declares ``ldpc::bp::BpDecoder``
(reference src_python/ldpc/bp_decoder/_bp_decoder.pxd:9-83).  This is the same kind of shim re-pointed at the thin
C-ABI: the ``cdef extern`` block below declares the plain-C entry points, ``NativeHandle`` owns one ``bpb_decoder*``
(the role of ``BpDecoderBase``'s ``bpd`` member, _bp_decoder.pxd:85-94), and every call releases the GIL, which
the reference's ``decode`` never does.  The Python classes in bp_decoder.py / bposd_decoder.py keep the reference's
argument handling and call through this object; ``_capi.py`` (ctypes) binds the same library for callers without a
compiled extension.
"""
from libc.stdint cimport uint8_t, int32_t, int64_t, uint64_t, uintptr_t

cdef extern from "bp_b200.h":
    ctypedef struct bpb_decoder:
        pass
    ctypedef struct bpb_info:
        int m
        int n
        int64_t nnz
        int max_row_degree
        int max_col_degree
        int device
        int sm_count
        int kernel_family
        int grid
        int block
        int64_t launches
        int64_t workspace_bytes
        double last_kernel_ms
        int smem_family_available
        int smem_bank_multiplicity
        int smem_bytes_per_syndrome
        int64_t stream_iterations
        int64_t stream_handed_off
        int osd_device_available
        int64_t osd_device_solved
        int64_t osd_host_solved
        int64_t osd_host_inconsistent
        int pair_family_available
        int pair_bank_multiplicity
    int bpb_create(int m, int n, int64_t nnz, const int32_t *rows, const int32_t *cols, int device,
                   bpb_decoder **out) nogil
    void bpb_destroy(bpb_decoder *h) nogil
    const char *bpb_last_error(const bpb_decoder *h) nogil
    int bpb_set_channel(bpb_decoder *h, const double *p, int n) nogil
    int bpb_set_max_iter(bpb_decoder *h, int v) nogil
    int bpb_set_method(bpb_decoder *h, int v) nogil
    int bpb_set_schedule(bpb_decoder *h, int v) nogil
    int bpb_set_ms_scaling_factor(bpb_decoder *h, double v) nogil
    int bpb_set_serial_schedule_order(bpb_decoder *h, const int32_t *order, int n) nogil
    int bpb_set_kernel(bpb_decoder *h, int v) nogil
    int bpb_decode_batch(bpb_decoder *h, int input_type, const uint8_t *inp, int64_t batch, uint8_t *decoding,
                         uint8_t *converged, int32_t *iterations, double *llr) nogil
    int bpb_bposd_decode_batch(bpb_decoder *h, const uint8_t *syndromes, int64_t batch, uint8_t *decoding,
                               uint8_t *converged, int32_t *iterations, uint8_t *bp_decoding, int threads) nogil
    int bpb_osd0_host(bpb_decoder *h, const uint8_t *syndromes, const double *llr, const uint8_t *converged,
                      int64_t batch, uint8_t *decoding, int threads) nogil
    int bpb_get_info(const bpb_decoder *h, bpb_info *out) nogil
    int bpb_set_osd_location(bpb_decoder *h, int v) nogil
    int bpb_set_devices(bpb_decoder *h, const int *ids, int count) nogil
    int bpb_get_last_schedule_order(bpb_decoder *h, int32_t *out, int len) nogil
    int bpb_soft_info_decode_batch(bpb_decoder *h, const double *soft, int64_t batch, double cutoff, double sigma,
                                   uint8_t *dec, uint8_t *conv, int32_t *its, double *llr, double *soft_out) nogil
    int bpb_set_observables(bpb_decoder *h, int k, int64_t nnz, const int32_t *rows, const int32_t *cols) nogil
    int bpb_decode_batch_b8(bpb_decoder *h, int with_osd, const uint8_t *syn, int64_t batch, uint8_t *dec,
                            uint8_t *obs, uint8_t *conv, int32_t *its) nogil
    int bpb_set_serial_schedule_order(bpb_decoder *h, const int32_t *order, int len) nogil
    int bpb_mc_bsc(bpb_decoder *h, uint64_t seed, int64_t first_run, int64_t runs, const double *flip_prob,
                   int with_osd, int64_t *counts) nogil


class NativeError(RuntimeError):
    pass


cdef class NativeHandle:
    cdef bpb_decoder *h

    def __cinit__(self, int m, int n, const int32_t[::1] rows, const int32_t[::1] cols, int device):
        self.h = NULL
        cdef int rc
        cdef const int32_t *rp = &rows[0] if rows.shape[0] else NULL
        cdef const int32_t *cp = &cols[0] if cols.shape[0] else NULL
        with nogil:
            rc = bpb_create(m, n, rows.shape[0], rp, cp, device, &self.h)
        if rc != 0:
            msg = bpb_last_error(NULL)
            raise NativeError(f"bpb_create failed ({rc}): {msg.decode() if msg != NULL else ''}")

    def __dealloc__(self):
        if self.h != NULL:
            bpb_destroy(self.h)
            self.h = NULL

    cdef _check(self, int rc):
        if rc != 0:
            msg = bpb_last_error(self.h)
            raise NativeError(f"ldpc_b200 C-ABI error {rc}: {msg.decode() if msg != NULL else ''}")

    @property
    def ptr(self):
        """The raw ``bpb_decoder*`` as an integer (for callers that drive bpb_decode_batch_device themselves)."""
        return <uintptr_t> self.h

    def set_channel(self, const double[::1] p):
        self._check(bpb_set_channel(self.h, &p[0], <int> p.shape[0]))

    def set_params(self, int max_iter, int method, int schedule, double ms_scaling_factor, const int32_t[::1] order,
                   int kernel):
        self._check(bpb_set_max_iter(self.h, max_iter))
        self._check(bpb_set_method(self.h, method))
        self._check(bpb_set_schedule(self.h, schedule))
        self._check(bpb_set_ms_scaling_factor(self.h, ms_scaling_factor))
        self._check(bpb_set_serial_schedule_order(self.h, &order[0] if order.shape[0] else NULL, <int> order.shape[0]))
        self._check(bpb_set_kernel(self.h, kernel))

    def decode_batch(self, int input_type, const uint8_t[:, ::1] inp, uint8_t[:, ::1] dec, uint8_t[::1] conv,
                     int32_t[::1] its, double[:, ::1] llr=None):
        cdef int64_t B = inp.shape[0]
        cdef int rc
        cdef double *lp = &llr[0, 0] if llr is not None else NULL
        if B == 0:
            return
        with nogil:
            rc = bpb_decode_batch(self.h, input_type, &inp[0, 0], B, &dec[0, 0], &conv[0], &its[0], lp)
        self._check(rc)

    def bposd_decode_batch(self, const uint8_t[:, ::1] syn, uint8_t[:, ::1] dec, uint8_t[::1] conv, int32_t[::1] its,
                           int threads, uint8_t[:, ::1] bp_dec=None):
        cdef int64_t B = syn.shape[0]
        cdef int rc
        cdef uint8_t *bp = &bp_dec[0, 0] if bp_dec is not None else NULL
        if B == 0:
            return
        with nogil:
            rc = bpb_bposd_decode_batch(self.h, &syn[0, 0], B, &dec[0, 0], &conv[0], &its[0], bp, threads)
        self._check(rc)

    def set_osd_location(self, int v):
        self._check(bpb_set_osd_location(self.h, v))

    def set_observables(self, int k, const int32_t[::1] rows, const int32_t[::1] cols):
        cdef int rc
        cdef int64_t nnz = rows.shape[0]
        cdef const int32_t *r = &rows[0] if nnz else NULL
        cdef const int32_t *c = &cols[0] if nnz else NULL
        with nogil:
            rc = bpb_set_observables(self.h, k, nnz, r, c)
        self._check(rc)

    def decode_batch_b8(self, int with_osd, const uint8_t[:, ::1] syn, uint8_t[:, ::1] dec, uint8_t[:, ::1] obs,
                        uint8_t[::1] conv, int32_t[::1] its):
        cdef int rc
        cdef int64_t B = syn.shape[0]
        if B == 0:
            return
        cdef uint8_t *pd = &dec[0, 0] if dec is not None else NULL
        cdef uint8_t *po = &obs[0, 0] if obs is not None else NULL
        cdef uint8_t *pc = &conv[0] if conv is not None else NULL
        cdef int32_t *pi = &its[0] if its is not None else NULL
        with nogil:
            rc = bpb_decode_batch_b8(self.h, with_osd, &syn[0, 0], B, pd, po, pc, pi)
        self._check(rc)

    def soft_info_decode_batch(self, const double[:, ::1] soft, double cutoff, double sigma, uint8_t[:, ::1] dec,
                               uint8_t[::1] conv, int32_t[::1] its, double[:, ::1] llr, double[:, ::1] soft_out):
        cdef int rc
        cdef int64_t B = soft.shape[0]
        if B == 0:
            return
        cdef double *pl = &llr[0, 0] if llr is not None else NULL
        cdef double *ps = &soft_out[0, 0] if soft_out is not None else NULL
        with nogil:
            rc = bpb_soft_info_decode_batch(self.h, &soft[0, 0], B, cutoff, sigma, &dec[0, 0], &conv[0], &its[0], pl, ps)
        self._check(rc)

    def last_schedule_order(self, int32_t[::1] out):
        cdef int rc
        cdef int k = <int> out.shape[0]
        with nogil:
            rc = bpb_get_last_schedule_order(self.h, &out[0], k)
        self._check(rc)

    def set_order(self, const int32_t[::1] order):
        cdef int rc
        cdef int k = <int> order.shape[0]
        with nogil:
            rc = bpb_set_serial_schedule_order(self.h, &order[0], k)
        self._check(rc)

    def set_devices(self, const int32_t[::1] ids):
        cdef int k = <int> ids.shape[0]
        cdef int rc
        cdef const int *p = <const int *> &ids[0] if k else NULL
        with nogil:
            rc = bpb_set_devices(self.h, p, k)
        self._check(rc)

    def mc_bsc(self, uint64_t seed, int64_t first_run, int64_t runs, const double[::1] flip_prob, int with_osd):
        cdef int64_t counts[5]
        cdef int rc
        cdef const double *fp = &flip_prob[0] if flip_prob is not None else NULL
        with nogil:
            rc = bpb_mc_bsc(self.h, seed, first_run, runs, fp, with_osd, counts)
        self._check(rc)
        return [counts[i] for i in range(5)]

    def osd0_host(self, const uint8_t[:, ::1] syn, const double[:, ::1] llr, const uint8_t[::1] conv,
                  uint8_t[:, ::1] dec, int threads):
        cdef int64_t B = syn.shape[0]
        cdef int rc
        if B == 0:
            return
        with nogil:
            rc = bpb_osd0_host(self.h, &syn[0, 0], &llr[0, 0], &conv[0], B, &dec[0, 0], threads)
        self._check(rc)

    def info(self):
        cdef bpb_info inf
        self._check(bpb_get_info(self.h, &inf))
        return {"m": inf.m, "n": inf.n, "nnz": inf.nnz, "max_row_degree": inf.max_row_degree,
                "max_col_degree": inf.max_col_degree, "device": inf.device, "sm_count": inf.sm_count,
                "kernel_family": inf.kernel_family, "grid": inf.grid, "block": inf.block, "launches": inf.launches,
                "workspace_bytes": inf.workspace_bytes, "last_kernel_ms": inf.last_kernel_ms,
                "smem_family_available": inf.smem_family_available,
                "smem_bank_multiplicity": inf.smem_bank_multiplicity,
                "smem_bytes_per_syndrome": inf.smem_bytes_per_syndrome, "stream_iterations": inf.stream_iterations,
                "stream_handed_off": inf.stream_handed_off,
                "osd_device_available": inf.osd_device_available, "osd_device_solved": inf.osd_device_solved,
                "osd_host_solved": inf.osd_host_solved, "osd_host_inconsistent": inf.osd_host_inconsistent,
                "pair_family_available": inf.pair_family_available,
                "pair_bank_multiplicity": inf.pair_bank_multiplicity}
