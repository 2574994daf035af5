"""oracle -- TEST INFRASTRUCTURE ONLY (see oracle/bp_oracle.c header).

ctypes loaders for the two CPU checkers:

* ``port``  -- oracle/_build/libbp_oracle.so, the plain-C restatement (bp_oracle.c, osd_oracle.c);
* ``ref``   -- oracle/_ref/libref_bp.so, the unmodified reference C++ behind oracle/ref_wrap.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package ``ldpc_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libbp_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_bp.so")

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile the checkers (make -C oracle).  Building the checker is not using it."""
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("oracle build failed")


def _coo(H):
    Hc = sp.coo_matrix(H)
    mask = Hc.data != 0
    rows = np.ascontiguousarray(Hc.row[mask], dtype=np.int32)
    cols = np.ascontiguousarray(Hc.col[mask], dtype=np.int32)
    return int(Hc.shape[0]), int(Hc.shape[1]), rows, cols


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


_METHODS = {"ps": 0, "product_sum": 0, 0: 0, "ms": 1, "minimum_sum": 1, 1: 1}
_SCHEDULES = {"serial": 0, "s": 0, "parallel": 1, "p": 1, "serial_relative": 2, "sr": 2}


class _Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.lib = C.CDLL(path)


class PortOracle(_Lib):
    """The plain-C restatement."""

    def __init__(self):
        super().__init__(PORT_SO)
        f = self.lib.bpo_decode_batch
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, C.c_int, C.c_int, C.c_int, C.c_double, _i32p,
                      C.c_int, _u8p, C.c_int64, _u8p, _u8p, _i32p, _f64p]
        g = self.lib.bpo_osd0_batch
        g.restype = C.c_int
        g.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _u8p, _f64p, C.c_int64, _u8p]
        self.lib.bpo_soft_info_decode_batch.restype = C.c_int
        self.lib.bpo_soft_info_decode_batch.argtypes = _SOFT_ARGS

    def soft_info_decode_batch(self, H, soft, channel, max_iter, ms_scaling_factor=1.0, cutoff=np.inf, sigma=2.0,
                               serial_schedule_order=None):
        """Restatement of soft_info_decode_serial (bp.hpp:547-665) -> (decoding, converged, iters, llr, soft)."""
        return _soft_call(self.lib.bpo_soft_info_decode_batch, H, soft, channel, max_iter, ms_scaling_factor, cutoff,
                          sigma, serial_schedule_order)

    def decode_batch(self, H, syndromes, channel, max_iter, bp_method="ms", schedule="parallel",
                     ms_scaling_factor=1.0, serial_schedule_order=None, want_llr=True):
        m, n, rows, cols = _coo(H)
        syn = np.ascontiguousarray(syndromes, dtype=np.uint8).reshape(-1, m)
        B = syn.shape[0]
        ch = np.ascontiguousarray(np.broadcast_to(np.asarray(channel, dtype=np.float64), (n,)))
        dec = np.zeros((B, n), np.uint8)
        conv = np.zeros(B, np.uint8)
        its = np.zeros(B, np.int32)
        llr = np.zeros((B, n), np.float64) if want_llr else None
        order = None if serial_schedule_order is None else np.ascontiguousarray(serial_schedule_order, dtype=np.int32)
        rc = self.lib.bpo_decode_batch(m, n, rows.size, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(ch, _f64p),
                                       int(max_iter), _METHODS[bp_method], _SCHEDULES[schedule],
                                       float(ms_scaling_factor), _ptr(order, _i32p),
                                       0 if order is None else order.size, _ptr(syn, _u8p), B, _ptr(dec, _u8p),
                                       _ptr(conv, _u8p), _ptr(its, _i32p), _ptr(llr, _f64p))
        if rc:
            raise RuntimeError(f"bpo_decode_batch failed: {rc}")
        return dec, conv.astype(bool), its, llr

    def osd0_batch(self, H, syndromes, llr):
        m, n, rows, cols = _coo(H)
        syn = np.ascontiguousarray(syndromes, dtype=np.uint8).reshape(-1, m)
        l = np.ascontiguousarray(llr, dtype=np.float64).reshape(-1, n)
        out = np.zeros((syn.shape[0], n), np.uint8)
        rc = self.lib.bpo_osd0_batch(m, n, rows.size, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(syn, _u8p),
                                     _ptr(l, _f64p), syn.shape[0], _ptr(out, _u8p))
        if rc:
            raise RuntimeError("bpo_osd0_batch failed")
        return out


def _soft_call(fn, H, soft, channel, max_iter, ms_scaling_factor, cutoff, sigma, serial_schedule_order):
    m, n, rows, cols = _coo(H)
    sf = np.ascontiguousarray(soft, dtype=np.float64).reshape(-1, m)
    B = sf.shape[0]
    ch = np.ascontiguousarray(np.broadcast_to(np.asarray(channel, dtype=np.float64), (n,)))
    order = None if serial_schedule_order is None else np.ascontiguousarray(serial_schedule_order, dtype=np.int32)
    dec = np.zeros((B, n), np.uint8)
    conv = np.zeros(B, np.uint8)
    its = np.zeros(B, np.int32)
    llr = np.zeros((B, n), np.float64)
    out_soft = np.zeros((B, m), np.float64)
    rc = fn(m, n, rows.size, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(ch, _f64p), int(max_iter),
            float(ms_scaling_factor), _ptr(order, _i32p), 0 if order is None else order.size, float(cutoff),
            float(sigma), _ptr(sf, _f64p), B, _ptr(dec, _u8p), _ptr(conv, _u8p), _ptr(its, _i32p), _ptr(llr, _f64p),
            _ptr(out_soft, _f64p))
    if rc < 0:
        raise RuntimeError("soft-info decode failed")
    return dec, conv.astype(bool), its, llr, out_soft


_SOFT_ARGS = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, C.c_int, C.c_double, _i32p, C.c_int, C.c_double,
              C.c_double, _f64p, C.c_int64, _u8p, _u8p, _i32p, _f64p, _f64p]


class RefOracle(_Lib):
    """The unmodified reference C++ (ldpc::bp::BpDecoder / ldpc::osd::OsdDecoder)."""

    def __init__(self):
        super().__init__(REF_SO)
        f = self.lib.ref_decode_batch
        f.restype = C.c_double
        f.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, C.c_int, C.c_int, C.c_int, C.c_double, _i32p,
                      C.c_int, C.c_int, C.c_int, _u8p, C.c_int64, _u8p, _u8p, _i32p, _f64p, _u8p, C.c_int]
        g = self.lib.ref_decode_received
        g.restype = C.c_double
        g.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, _f64p, C.c_int, C.c_int, C.c_int, C.c_double, _u8p,
                      C.c_int64, _u8p]
        self.lib.ref_hardware_threads.restype = C.c_int
        if hasattr(self.lib, "ref_soft_info_decode_batch"):
            self.lib.ref_soft_info_decode_batch.restype = C.c_double
            self.lib.ref_soft_info_decode_batch.argtypes = _SOFT_ARGS

    def soft_info_decode_batch(self, H, soft, channel, max_iter, ms_scaling_factor=1.0, cutoff=np.inf, sigma=2.0,
                               serial_schedule_order=None):
        """ldpc::bp::BpDecoder::soft_info_decode_serial per row -> (decoding, converged, iters, llr, soft_syndrome)."""
        return _soft_call(self.lib.ref_soft_info_decode_batch, H, soft, channel, max_iter, ms_scaling_factor, cutoff,
                          sigma, serial_schedule_order)

    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())

    def decode_batch(self, H, syndromes, channel, max_iter, bp_method="ms", schedule="parallel",
                     ms_scaling_factor=1.0, serial_schedule_order=None, want_llr=True, osd_method=0, osd_order=0,
                     threads=1, return_seconds=False):
        m, n, rows, cols = _coo(H)
        syn = np.ascontiguousarray(syndromes, dtype=np.uint8).reshape(-1, m)
        B = syn.shape[0]
        ch = np.ascontiguousarray(np.broadcast_to(np.asarray(channel, dtype=np.float64), (n,)))
        dec = np.zeros((B, n), np.uint8)
        bpdec = np.zeros((B, n), np.uint8) if osd_method else None
        conv = np.zeros(B, np.uint8)
        its = np.zeros(B, np.int32)
        llr = np.zeros((B, n), np.float64) if want_llr else None
        order = None if serial_schedule_order is None else np.ascontiguousarray(serial_schedule_order, dtype=np.int32)
        secs = self.lib.ref_decode_batch(m, n, rows.size, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(ch, _f64p),
                                         int(max_iter), _METHODS[bp_method], _SCHEDULES[schedule],
                                         float(ms_scaling_factor), _ptr(order, _i32p),
                                         0 if order is None else order.size, int(osd_method), int(osd_order),
                                         _ptr(syn, _u8p), B, _ptr(dec, _u8p), _ptr(conv, _u8p), _ptr(its, _i32p),
                                         _ptr(llr, _f64p), _ptr(bpdec, _u8p), int(threads))
        if secs < 0:
            raise RuntimeError("ref_decode_batch failed")
        out = (dec, conv.astype(bool), its, llr)
        if osd_method:
            out = out + (bpdec,)
        if return_seconds:
            out = out + (secs,)
        return out

    def decode_received(self, H, vectors, channel, max_iter, bp_method="ps", schedule="parallel",
                        ms_scaling_factor=1.0):
        m, n, rows, cols = _coo(H)
        v = np.ascontiguousarray(vectors, dtype=np.uint8).reshape(-1, n)
        ch = np.ascontiguousarray(np.broadcast_to(np.asarray(channel, dtype=np.float64), (n,)))
        dec = np.zeros((v.shape[0], n), np.uint8)
        rc = self.lib.ref_decode_received(m, n, rows.size, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(ch, _f64p),
                                          int(max_iter), _METHODS[bp_method], _SCHEDULES[schedule],
                                          float(ms_scaling_factor), _ptr(v, _u8p), v.shape[0], _ptr(dec, _u8p))
        if rc < 0:
            raise RuntimeError("ref_decode_received failed")
        return dec


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_port() -> bool:
    return os.path.exists(PORT_SO)
