"""One native decoder handle (``bpb_decoder*`` of include/bp_b200.h).

Calls go through the Cython binding ``_bp_shim`` (the reference's kind of host shim, built by
``ldpc_b200/csrc/build_shim.py`` / ``__graft_entry__.build()``) and fall back to the ``ctypes`` binding of the SAME
library when the extension has not been built.  Both are bindings of the CUDA library; neither decodes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi

try:  # the compiled Cython shim
    from . import _bp_shim
except ImportError:  # pragma: no cover - depends on the build
    _bp_shim = None

BINDING = "cython" if _bp_shim is not None else "ctypes"


class Handle:
    def __init__(self, m: int, n: int, rows: np.ndarray, cols: np.ndarray, device: int):
        self.m, self.n = m, n
        self._nh = None
        self._ct = None
        if _bp_shim is not None:
            _capi.lib()  # raises a clear ImportError if libbp_b200.so itself is missing
            try:
                self._nh = _bp_shim.NativeHandle(m, n, rows, cols, device)
            except _bp_shim.NativeError as e:
                raise _capi.BpbError(str(e)) from None
            self.ptr = C.c_void_p(self._nh.ptr)
        else:
            L = _capi.lib()
            h = C.c_void_p()
            rc = L.bpb_create(m, n, rows.size, rows.ctypes.data_as(_capi._i32p), cols.ctypes.data_as(_capi._i32p),
                              device, C.byref(h))
            if rc != _capi.OK:
                msg = L.bpb_last_error(None)
                raise _capi.BpbError(f"bpb_create failed ({rc}): {msg.decode() if msg else ''}")
            self._ct = h
            self.ptr = h

    def close(self):
        if self._ct is not None:
            _capi.lib().bpb_destroy(self._ct)
            self._ct = None
        self._nh = None  # NativeHandle.__dealloc__ destroys the decoder

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _wrap(self, fn, *args):
        try:
            return fn(*args)
        except _bp_shim.NativeError as e:
            raise _capi.BpbError(str(e)) from None

    def configure(self, channel, max_iter, method, schedule, ms_scaling, order, kernel):
        ch = np.ascontiguousarray(channel, dtype=np.float64)
        od = np.ascontiguousarray(order, dtype=np.int32)
        if self._nh is not None:
            self._wrap(self._nh.set_channel, ch)
            self._wrap(self._nh.set_params, int(max_iter), int(method), int(schedule), float(ms_scaling), od, int(kernel))
            return
        L, h = _capi.lib(), self._ct
        _capi.check(h, L.bpb_set_channel(h, ch.ctypes.data_as(_capi._f64p), self.n))
        _capi.check(h, L.bpb_set_max_iter(h, int(max_iter)))
        _capi.check(h, L.bpb_set_method(h, int(method)))
        _capi.check(h, L.bpb_set_schedule(h, int(schedule)))
        _capi.check(h, L.bpb_set_ms_scaling_factor(h, float(ms_scaling)))
        _capi.check(h, L.bpb_set_serial_schedule_order(h, od.ctypes.data_as(_capi._i32p), od.size))
        _capi.check(h, L.bpb_set_kernel(h, int(kernel)))

    def decode_batch(self, input_type, inputs, dec, conv, its, llr):
        if self._nh is not None:
            return self._wrap(self._nh.decode_batch, int(input_type), inputs, dec, conv, its, llr)
        h = self._ct
        rc = _capi.lib().bpb_decode_batch(h, input_type, _capi.host_ptr(inputs), inputs.shape[0], _capi.host_ptr(dec),
                                          _capi.host_ptr(conv), _capi.host_ptr(its), _capi.host_ptr(llr))
        _capi.check(h, rc)

    def bposd_decode_batch(self, syn, dec, conv, its, threads, bp_dec=None):
        if self._nh is not None:
            return self._wrap(self._nh.bposd_decode_batch, syn, dec, conv, its, int(threads), bp_dec)
        h = self._ct
        rc = _capi.lib().bpb_bposd_decode_batch(h, _capi.host_ptr(syn), syn.shape[0], _capi.host_ptr(dec),
                                                _capi.host_ptr(conv), _capi.host_ptr(its), _capi.host_ptr(bp_dec),
                                                int(threads))
        _capi.check(h, rc)

    def set_osd_location(self, where: int):
        if self._nh is not None:
            return self._wrap(self._nh.set_osd_location, int(where))
        _capi.check(self._ct, _capi.lib().bpb_set_osd_location(self._ct, int(where)))

    def set_observables(self, k, rows, cols):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        if self._nh is not None:
            return self._wrap(self._nh.set_observables, int(k), rows, cols)
        _capi.check(self._ct, _capi.lib().bpb_set_observables(self._ct, int(k), rows.size, rows.ctypes.data_as(_capi._i32p),
                                                               cols.ctypes.data_as(_capi._i32p)))

    def decode_batch_b8(self, with_osd, syn, dec, obs, conv, its):
        if self._nh is not None:
            return self._wrap(self._nh.decode_batch_b8, int(with_osd), syn, dec, obs, conv, its)
        rc = _capi.lib().bpb_decode_batch_b8(self._ct, int(with_osd), _capi.host_ptr(syn), syn.shape[0],
                                             _capi.host_ptr(dec), _capi.host_ptr(obs), _capi.host_ptr(conv),
                                             _capi.host_ptr(its))
        _capi.check(self._ct, rc)

    def soft_info_decode_batch(self, soft, cutoff, sigma, dec, conv, its, llr, soft_out):
        if self._nh is not None:
            return self._wrap(self._nh.soft_info_decode_batch, soft, float(cutoff), float(sigma), dec, conv, its, llr,
                              soft_out)
        rc = _capi.lib().bpb_soft_info_decode_batch(self._ct, _capi.host_ptr(soft), soft.shape[0], float(cutoff),
                                                    float(sigma), _capi.host_ptr(dec), _capi.host_ptr(conv),
                                                    _capi.host_ptr(its), _capi.host_ptr(llr), _capi.host_ptr(soft_out))
        _capi.check(self._ct, rc)

    def last_schedule_order(self, n):
        """SERIAL_RELATIVE: the schedule the last syndrome of the last decode call ended with."""
        out = np.empty(n, dtype=np.int32)
        if self._nh is not None:
            self._wrap(self._nh.last_schedule_order, out)
            return out
        _capi.check(self._ct, _capi.lib().bpb_get_last_schedule_order(self._ct, out.ctypes.data_as(_capi._i32p), n))
        return out

    def set_order(self, order):
        od = np.ascontiguousarray(order, dtype=np.int32)
        if self._nh is not None:
            return self._wrap(self._nh.set_order, od)
        _capi.check(self._ct, _capi.lib().bpb_set_serial_schedule_order(self._ct, od.ctypes.data_as(_capi._i32p), od.size))

    def set_devices(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        if self._nh is not None:
            return self._wrap(self._nh.set_devices, ids)
        _capi.check(self._ct, _capi.lib().bpb_set_devices(self._ct, ids.ctypes.data_as(_capi._i32p), ids.size))

    def mc_bsc(self, seed, first_run, runs, flip_prob, with_osd):
        fp = None if flip_prob is None else np.ascontiguousarray(flip_prob, dtype=np.float64)
        if self._nh is not None:
            return self._wrap(self._nh.mc_bsc, int(seed), int(first_run), int(runs), fp, int(with_osd))
        counts = (C.c_int64 * 5)()
        rc = _capi.lib().bpb_mc_bsc(self._ct, int(seed), int(first_run), int(runs),
                                    None if fp is None else fp.ctypes.data_as(_capi._f64p), int(with_osd), counts)
        _capi.check(self._ct, rc)
        return list(counts)

    def osd0_host(self, syn, llr, conv, dec, threads):
        if self._nh is not None:
            return self._wrap(self._nh.osd0_host, syn, llr, conv, dec, int(threads))
        h = self._ct
        rc = _capi.lib().bpb_osd0_host(h, _capi.host_ptr(syn), _capi.host_ptr(llr), _capi.host_ptr(conv), syn.shape[0],
                                       _capi.host_ptr(dec), int(threads))
        _capi.check(h, rc)

    def info(self) -> dict:
        if self._nh is not None:
            return self._wrap(self._nh.info)
        inf = _capi.BpbInfo()
        _capi.check(self._ct, _capi.lib().bpb_get_info(self._ct, C.byref(inf)))
        return {k: getattr(inf, k) for k, _ in inf._fields_}
