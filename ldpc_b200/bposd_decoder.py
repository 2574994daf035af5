"""BpOsdDecoder: BP on the B200, OSD-0 for the syndromes BP did not solve -- on the device too when the code fits the
bit-packed elimination kernel (``osd_location='auto'``), else on the host.

Mirrors ``ldpc.bposd_decoder.BpOsdDecoder`` (reference src_python/ldpc/bposd_decoder/_bposd_decoder.pyx:8-299):
same constructor keywords and aliases, ``decode`` = zero shortcut -> BP -> (if not converged) OSD
(``_bposd_decoder.pyx:78-136``).  BASELINE.json keeps OSD-0 on the host; higher-order OSD (OSD_E / OSD_CS,
osd.hpp:119-187) is out of scope here and raises ``NotImplementedError``.
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Union

import numpy as np

from . import _capi
from .bp_decoder import BpDecoderBase, _BP_KWARGS

OSD_OFF, OSD_0, OSD_E, OSD_CS = 0, 1, 2, 3  # reference osd.hpp:18-23


class BpOsdDecoder(BpDecoderBase):
    def __init__(self, pcm, error_rate: Optional[float] = None, error_channel=None, max_iter: Optional[int] = 0,
                 bp_method=0, ms_scaling_factor=1.0, schedule=0, omp_thread_count: Optional[int] = 1,
                 random_schedule_seed: Optional[int] = 0, serial_schedule_order: Optional[List[int]] = None,
                 osd_method: Union[str, int, float] = 0, osd_order: int = 0, input_vector_type: str = "syndrome",
                 **kwargs):
        for key in kwargs.keys():
            if key not in _BP_KWARGS + ("osd_threads", "random_serial_schedule", "osd_location"):
                raise ValueError(f"Unknown parameter '{key}' passed to the BpDecoder constructor.")
        self._osd_threads = int(kwargs.pop("osd_threads", 0))
        osd_location = str(kwargs.pop("osd_location", "auto")).lower()
        if osd_location not in ("auto", "host", "device"):
            raise ValueError("osd_location must be 'auto', 'host' or 'device'")
        super().__init__(pcm, error_rate=error_rate, error_channel=error_channel, max_iter=max_iter,
                         bp_method=bp_method, ms_scaling_factor=ms_scaling_factor, schedule=schedule,
                         omp_thread_count=omp_thread_count, random_schedule_seed=random_schedule_seed,
                         serial_schedule_order=serial_schedule_order, **kwargs)
        self._osd_location = {"auto": _capi.OSD_AUTO, "host": _capi.OSD_HOST, "device": _capi.OSD_DEVICE}[osd_location]
        self.bp_decoding_batch = None
        self._osd_method = OSD_OFF
        self._osd_order = 0
        self.osd_method = osd_method
        self.osd_order = osd_order
        self.input_vector_type = "syndrome"  # forced, _bposd_decoder.pyx:68
        self._osd0_decoding = np.zeros(self.n, dtype=np.uint8)
        self._osdw_decoding = np.zeros(self.n, dtype=np.uint8)

    # ------------------------------------------------------------------ OSD properties (:140-233)
    @property
    def osd_method(self) -> Optional[str]:
        return {OSD_0: "OSD_0", OSD_E: "OSD_E", OSD_CS: "OSD_CS", OSD_OFF: "OSD_OFF"}.get(self._osd_method)

    @osd_method.setter
    def osd_method(self, method: Union[str, int, float]) -> None:
        key = str(method).lower()
        if key in ["osd_0", "0", "osd0"]:
            self._osd_method = OSD_0
            self._osd_order = 0
        elif key in ["osd_e", "e", "exhaustive"]:
            self._osd_method = OSD_E
        elif key in ["osd_cs", "1", "cs", "combination_sweep"]:
            self._osd_method = OSD_CS
        elif key in ["off", "osd_off", "deactivated"]:
            self._osd_method = OSD_OFF
        else:
            raise ValueError(f"ERROR: OSD method '{method}' invalid. Please choose from the following methods:\
                'OSD_0', 'OSD_E' or 'OSD_CS'.")

    @property
    def osd_order(self) -> int:
        return self._osd_order

    @osd_order.setter
    def osd_order(self, order: int) -> None:
        if order < 0:
            raise ValueError(f"ERROR: OSD order '{order}' invalid. Please choose a positive integer.")
        if self._osd_method == OSD_0 and order != 0:
            raise ValueError(f"ERROR: OSD order '{order}' invalid. The 'osd_method' is set to 'OSD_0'. The osd order must therefore be set to 0.")
        if self._osd_method == OSD_E and order > 15:
            warnings.warn("WARNING: Running the 'OSD_E' (Exhaustive method) with search depth greater than 15 is not "
                          "recommended. Use the 'osd_cs' method instead.")
        self._osd_order = order

    def _osd0(self, syndromes, llr, converged, decoding):
        if self._osd_method == OSD_OFF:
            return
        if self._osd_order != 0:
            raise NotImplementedError("only OSD-0 (osd_order == 0) is implemented; OSD_E / OSD_CS are out of scope")
        self._ensure_handle()
        conv = np.ascontiguousarray(converged, dtype=np.uint8)
        self._native.osd0_host(syndromes, np.ascontiguousarray(llr, dtype=np.float64), conv, decoding,
                               self._osd_threads)

    # ------------------------------------------------------------------ decode (:78-136)
    def decode(self, syndrome: np.ndarray) -> np.ndarray:
        syndrome = np.asarray(syndrome)
        if not len(syndrome) == self.m:
            raise ValueError(f"The syndrome must have length {self.m}. Not {len(syndrome)}.")
        dtype = syndrome.dtype
        vec = np.ascontiguousarray(syndrome.astype(np.uint8, copy=False)).reshape(1, -1)
        if not vec.any():
            self._converge = True
            return np.zeros(self.n, dtype=dtype)
        dec, conv, its, llr = self._decode_device_batch(vec, _capi.INPUT_SYNDROME, want_llr=True)
        self._decoding = dec[0].copy()
        self._converge = bool(conv[0])
        self._iterations = int(its[0])
        self._log_prob_ratios = llr[0]
        if not conv[0]:
            self._osd0(vec, llr, conv, dec)
            self._osd0_decoding = dec[0].copy()
            self._osdw_decoding = dec[0].copy()
        return dec[0].astype(dtype)

    @property
    def _b8_with_osd(self) -> bool:  # decode_batch_b8 (BpDecoderBase): BP + OSD-0 on the device
        if self._osd_method != OSD_OFF and self._osd_order != 0:
            raise NotImplementedError("only OSD-0 (osd_order == 0) is implemented; OSD_E / OSD_CS are out of scope")
        return self._osd_method != OSD_OFF

    def decode_batch(self, syndromes: np.ndarray, return_bp_decoding: bool = False) -> np.ndarray:
        """Decode ``[B, m]`` syndromes: BP for all on the GPU, then OSD-0 for the non-converged rows (on the device when
        the code fits the elimination kernel, else on the host).  ``return_bp_decoding`` also keeps the raw BP output
        in ``bp_decoding_batch``."""
        arr = np.asarray(syndromes)
        if arr.ndim != 2 or arr.shape[1] != self.m:
            raise ValueError(f"The syndromes must have shape [batch, {self.m}].")
        dtype = arr.dtype
        vec = np.ascontiguousarray(arr.astype(np.uint8, copy=False))
        if vec.shape[0] == 0:
            return np.zeros((0, self.n), dtype=dtype)
        if self._osd_method != OSD_OFF and self._osd_order != 0:
            raise NotImplementedError("only OSD-0 (osd_order == 0) is implemented; OSD_E / OSD_CS are out of scope")
        self._ensure_handle()
        B = vec.shape[0]
        big = B * self.n >= (1 << 20)
        alloc = _capi.pinned_empty if big else np.empty
        dec = alloc((B, self.n), dtype=np.uint8)
        conv = alloc((B,), dtype=np.uint8)
        its = alloc((B,), dtype=np.int32)
        self.bp_decoding_batch = None
        if self._osd_method == OSD_OFF:
            d, c, i, _ = self._decode_device_batch(vec, _capi.INPUT_SYNDROME, want_llr=False)
            dec, conv, its = d, c.astype(np.uint8), i
        else:
            bp = alloc((B, self.n), dtype=np.uint8) if return_bp_decoding else None
            self._native.bposd_decode_batch(vec, dec, conv, its, self._osd_threads, bp)
            self.bp_decoding_batch = bp
        self.converge_batch, self.iter_batch, self.log_prob_ratios_batch = conv.astype(bool), its, None
        return dec if dtype == np.uint8 else dec.astype(dtype)

    @property
    def bp_decoding(self) -> np.ndarray:
        return self._decoding.astype(int)

    @property
    def osd0_decoding(self) -> np.ndarray:
        return (self._decoding if self._converge else self._osd0_decoding).astype(int)

    @property
    def osdw_decoding(self) -> np.ndarray:
        return (self._decoding if self._converge else self._osdw_decoding).astype(int)

    @property
    def decoding(self) -> np.ndarray:
        return self.osdw_decoding
