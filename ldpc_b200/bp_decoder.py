"""BpDecoderBase / BpDecoder: the reference's Python decoder API on top of the B200 C-ABI.

Mirrors ``ldpc.bp_decoder.BpDecoderBase`` and ``ldpc.bp_decoder.BpDecoder`` (reference
``src_python/ldpc/bp_decoder/_bp_decoder.pyx:82-709``): same constructor keywords, effective defaults
(product_sum, parallel, ms_scaling_factor 1.0, max_iter 0 -> n; ``_bp_decoder.pyx:88-100``), string aliases,
exception types, properties, the all-zero shortcut and the output-dtype echo of ``decode``
(``_bp_decoder.pyx:642-695``).  New on top: ``decode_batch`` and the ``device=`` keyword.

All decoding happens in hand-written sm_100a CUDA behind ``include/bp_b200.h``; there is no CPU
fallback.  ``schedule='serial_relative'`` (the schedule re-sorted by posterior LLR before every sweep, with
libstdc++'s ``std::sort`` restated on the device) is supported; the random serial schedule (a shared RNG stream
that advances per iteration, DESIGN.md "out of scope") raises ``NotImplementedError`` at decode time rather
than silently changing behaviour.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import List, Optional, Union

import numpy as np
import scipy.sparse

from . import _capi
from ._native import Handle
from .helpers import convert_to_binary_sparse


def _coo_of(pcm):
    if not isinstance(pcm, (np.ndarray, scipy.sparse.spmatrix)):
        raise TypeError(f"The input matrix is of an invalid type. Please input\
        a np.ndarray or scipy.sparse.spmatrix object, not {type(pcm)}")
    mat = convert_to_binary_sparse(pcm)
    coo = scipy.sparse.coo_matrix(mat)
    rows = np.ascontiguousarray(coo.row, dtype=np.int32)
    cols = np.ascontiguousarray(coo.col, dtype=np.int32)
    return int(mat.shape[0]), int(mat.shape[1]), rows, cols


def io_test(pcm: Union[scipy.sparse.spmatrix, np.ndarray]):
    """Round-trip H through the flattening used by the decoder (reference ``io_test``, _bp_decoder.pyx:74-78)."""
    m, n, rows, cols = _coo_of(pcm)
    data = np.ones(rows.size, dtype=np.uint8)
    return scipy.sparse.csr_matrix((data, (rows, cols)), shape=(m, n), dtype=np.uint8)


class BpDecoderBase:
    """Bp Decoder base class (reference ``BpDecoderBase``, _bp_decoder.pyx:82-579)."""

    def __init__(self, pcm, **kwargs):
        error_rate = kwargs.get("error_rate", None)
        error_channel = kwargs.get("error_channel", None)
        max_iter = kwargs.get("max_iter", 0)
        bp_method = kwargs.get("bp_method", 0)
        ms_scaling_factor = kwargs.get("ms_scaling_factor", 1.0)
        schedule = kwargs.get("schedule", 0)
        omp_thread_count = kwargs.get("omp_thread_count", 1)
        random_serial_schedule = kwargs.get("random_serial_schedule", False)
        random_schedule_seed = kwargs.get("random_schedule_seed", 0)
        serial_schedule_order = kwargs.get("serial_schedule_order", None)
        channel_probs = kwargs.get("channel_probs", [None])
        devices = kwargs.get("devices", None)
        self._devices = None if devices is None else [int(d) for d in devices]
        if self._devices is not None and len(self._devices) < 1:
            raise ValueError("devices must name at least one CUDA device")
        self._device = int(kwargs.get("device", self._devices[0] if self._devices else 0))
        self._kernel = kwargs.get("kernel", "auto")
        self._osd_location = _capi.OSD_AUTO

        self._native = None  # the bpb_decoder handle (Cython binding, or ctypes when the extension is not built)
        self._handle = None
        self._dirty = True
        self.m, self.n, self._rows, self._cols = _coo_of(pcm)
        self._channel = np.zeros(self.n, dtype=np.float64)
        self._input_type = _capi.INPUT_SYNDROME
        self._omp_thread_count = 1
        self._random_schedule_seed = 0
        self._random_serial_schedule = False
        self._serial_schedule_order = np.arange(self.n, dtype=np.int64)  # bp.hpp:122-127
        # per-decode outputs (reference members bp.hpp:63-72)
        self._converge = False
        self._iterations = 0
        self._log_prob_ratios = np.zeros(self.n, dtype=np.float64)
        self._decoding = np.zeros(self.n, dtype=np.uint8)
        self.converge_batch = None
        self.iter_batch = None
        self.log_prob_ratios_batch = None

        self.bp_method = bp_method
        self.max_iter = max_iter
        self.ms_scaling_factor = ms_scaling_factor
        self.schedule = schedule
        self.serial_schedule_order = serial_schedule_order
        self.random_schedule_seed = random_schedule_seed
        self.omp_thread_count = omp_thread_count
        self.random_serial_schedule = random_serial_schedule

        # the ldpc_v1 backwards compatibility (_bp_decoder.pyx:144-147)
        if isinstance(channel_probs, (list, np.ndarray)):
            if len(channel_probs) > 0 and channel_probs[0] is not None:
                error_channel = channel_probs

        if error_channel is not None:
            self.error_channel = error_channel
        elif error_rate is not None:
            self.error_rate = error_rate
        else:
            raise ValueError("Please specify the error channel. Either: 1) error_rate: float or 2) error_channel:\
            list of floats of length equal to the block length of the code {self.n}.")

    # ------------------------------------------------------------------ C-ABI plumbing
    def _ensure_handle(self):
        """Create / refresh the native decoder; returns the raw ``bpb_decoder*`` (ctypes.c_void_p)."""
        if self._native is None:
            self._native = Handle(self.m, self.n, self._rows, self._cols, self._device)
            self._handle = self._native.ptr
            self._dirty = True
            if self._devices is not None:
                # one handle, one full decoder per device behind it (bpb_set_devices): host batches are split into
                # contiguous slices, the D2H copies land in disjoint ranges of the output arrays
                self._native.set_devices(self._devices)
        if self._dirty:
            if self._random_serial_schedule:
                raise NotImplementedError("random_serial_schedule is not implemented on the GPU path")
            kern = {"auto": _capi.KERNEL_AUTO, "stream": _capi.KERNEL_STREAM, "smem": _capi.KERNEL_SMEM,
                    "edge": _capi.KERNEL_EDGE, "pair": _capi.KERNEL_PAIR}[str(self._kernel).lower()]
            self._native.configure(self._channel, self._max_iter, self._bp_method, self._schedule,
                                   self._ms_scaling_factor, self._serial_schedule_order, kern)
            self._native.set_osd_location(self._osd_location)
            self._dirty = False
        return self._handle

    def _decode_device_batch(self, inputs: np.ndarray, input_type: int, want_llr: bool, out=None):
        """inputs: contiguous uint8 [B, m|n].  Returns (decoding u8 [B,n], converged bool[B], iters i32[B], llr|None)."""
        self._ensure_handle()
        B = inputs.shape[0]
        big = B * self.n >= (1 << 20)  # large results land in pinned memory (asynchronous D2H at full PCIe speed)
        alloc = _capi.pinned_empty if big else np.empty
        if out is not None:
            if not (isinstance(out, np.ndarray) and out.dtype == np.uint8 and out.shape == (B, self.n)
                    and out.flags.c_contiguous and out.flags.writeable):
                raise ValueError(f"out must be a writable C-contiguous uint8 array of shape ({B}, {self.n})")
            dec = out
        else:
            dec = alloc((B, self.n), dtype=np.uint8)
        conv = alloc((B,), dtype=np.uint8)
        its = alloc((B,), dtype=np.int32)
        llr = alloc((B, self.n), dtype=np.float64) if want_llr else None
        self._native.decode_batch(input_type, inputs, dec, conv, its, llr)
        return dec, conv.astype(bool), its, llr

    # ------------------------------------------------------------------ bit-packed I/O (stim's b8 layout)
    def set_observables(self, observables_matrix) -> None:
        """``k x n`` binary matrix O: ``decode_batch_b8(..., observables=True)`` then returns ``O x mod 2`` for every
        decoded row, computed on the device (the reference's sinter driver does this product per shot on the host,
        sinter_bposd_decoder.py:128-130)."""
        import scipy.sparse as sp
        O = sp.coo_matrix(observables_matrix)
        if O.shape[1] != self.n:
            raise ValueError(f"observables_matrix must have {self.n} columns")
        keep = O.data != 0
        self._ensure_handle()
        self._native.set_observables(O.shape[0], O.row[keep].astype(np.int32), O.col[keep].astype(np.int32))
        self._obs_rows = int(O.shape[0])

    _b8_with_osd = False

    def decode_batch_b8(self, syndromes_b8: np.ndarray, decoding: bool = True, observables: bool = False):
        """Decode bit-packed syndromes: ``[B, ceil(m/8)]`` uint8 rows in stim's b8 layout (bit k of a row = bit k % 8
        of byte k // 8).  Returns the packed decisions ``[B, ceil(n/8)]`` and / or the packed observable parities
        ``[B, ceil(k/8)]`` (after ``set_observables``); ``converge_batch`` / ``iter_batch`` are filled as by
        ``decode_batch``.  Only the packed rows cross PCIe: an eighth of ``decode_batch``'s traffic or less."""
        arr = np.ascontiguousarray(syndromes_b8, dtype=np.uint8)
        mb = (self.m + 7) // 8
        if arr.ndim != 2 or arr.shape[1] != mb:
            raise ValueError(f"syndromes_b8 must have shape [batch, {mb}]")
        if not decoding and not observables:
            raise ValueError("nothing to return")
        if observables and not getattr(self, "_obs_rows", 0):
            raise ValueError("call set_observables first")
        self._ensure_handle()
        B = arr.shape[0]
        alloc = _capi.pinned_empty if B * mb >= (1 << 20) else np.empty
        dec = alloc((B, (self.n + 7) // 8), dtype=np.uint8) if decoding else None
        obs = alloc((B, (self._obs_rows + 7) // 8), dtype=np.uint8) if observables else None
        conv = alloc((B,), dtype=np.uint8)
        its = alloc((B,), dtype=np.int32)
        self._native.decode_batch_b8(1 if self._b8_with_osd else 0, arr, dec, obs, conv, its)
        self.converge_batch, self.iter_batch = conv.astype(bool), its
        if decoding and observables:
            return dec, obs
        return dec if decoding else obs

    def monte_carlo_bsc(self, runs: int, seed: int = 0, error_rate=None, first_run: int = 0,
                        with_osd: bool = False) -> dict:
        """``runs`` Monte-Carlo runs on the binary symmetric channel entirely on the device: errors drawn with
        Philox4x32-10 (run r, bit j -> counter (r, j // 4), word j % 4, key = seed), syndromes ``H e``, decode,
        compare ``decoding != error`` (the loop body of the reference's ``MonteCarloBscSimulation.run``,
        mcs.py:124-139).  Only counters cross PCIe.  ``error_rate``: flip probability (scalar or per bit); default
        is the decoder's own channel."""
        self._ensure_handle()
        fp = None
        if error_rate is not None:
            fp = np.ascontiguousarray(np.broadcast_to(np.asarray(error_rate, dtype=np.float64), (self.n,)))
        c = self._native.mc_bsc(int(seed), int(first_run), int(runs), fp, 1 if with_osd else 0)
        return {"run_count": int(c[0]), "fail_count": int(c[1]), "bp_converged": int(c[2]), "iter_sum": int(c[3]),
                "converged_wrong": int(c[4])}

    def info(self) -> dict:
        """Introspection of the native handle (kernel family, launch shape, launches so far)."""
        self._ensure_handle()
        return self._native.info()

    # ------------------------------------------------------------------ properties (reference :167-579)
    @property
    def error_rate(self) -> np.ndarray:
        return self._channel.astype(float).copy()

    @error_rate.setter
    def error_rate(self, value: Optional[float]) -> None:
        if value is not None:
            if not isinstance(value, float):
                raise ValueError("The `error_rate` parameter must be specified as a single float value.")
            self._channel[:] = value
            self._dirty = True

    @property
    def error_channel(self) -> np.ndarray:
        return self._channel.astype(float).copy()

    @error_channel.setter
    def error_channel(self, value) -> None:
        if value is not None:
            if len(value) != self.n:
                raise ValueError(f"The error channel vector must have length {self.n}, not {len(value)}.")
            for i in range(self.n):
                self._channel[i] = value[i]
            self._dirty = True

    def update_channel_probs(self, value) -> None:
        self.error_channel = value

    @property
    def channel_probs(self) -> np.ndarray:
        return self._channel.astype(float).copy()

    @property
    def input_vector_type(self) -> str:
        return {_capi.INPUT_SYNDROME: "syndrome", _capi.INPUT_RECEIVED_VECTOR: "received_vector",
                _capi.INPUT_AUTO: "auto"}[self._input_type]

    @input_vector_type.setter
    def input_vector_type(self, input_type: str):
        if input_type.lower() in ["auto", "a", "2"]:
            if self.m == self.n:
                raise ValueError("Please specify the input vector type. Either: 1) input_vector_type: 'syndrome' or 2) input_vector_type:\
                'received_vector'.")
            self._input_type = _capi.INPUT_AUTO
        elif input_type.lower() in ["syndrome", "s", "0"]:
            self._input_type = _capi.INPUT_SYNDROME
        elif input_type.lower() in ["received_vector", "r", "1"]:
            self._input_type = _capi.INPUT_RECEIVED_VECTOR
        else:
            raise ValueError(f"The input vector type '{input_type}' is invalid. \
                    Please choose from the following methods: \
                    'input_vector_type=syndrome', 'input_vector_type=received_vector'")

    @property
    def log_prob_ratios(self) -> np.ndarray:
        return self._log_prob_ratios.copy()

    @property
    def converge(self) -> bool:
        return bool(self._converge)

    @property
    def iter(self) -> int:
        return int(self._iterations)

    @property
    def check_count(self) -> int:
        return self.m

    @property
    def bit_count(self) -> int:
        return self.n

    @property
    def max_iter(self) -> int:
        return self._max_iter

    @max_iter.setter
    def max_iter(self, value: int) -> None:
        if not isinstance(value, int):
            raise ValueError("max_iter input parameter is invalid. This must be specified as a positive int.")
        if value < 0:
            raise ValueError(f"max_iter input parameter must be a positive int. Not {value}.")
        self._max_iter = value if value != 0 else self.n
        self._dirty = True

    @property
    def bp_method(self) -> str:
        return "product_sum" if self._bp_method == _capi.PRODUCT_SUM else "minimum_sum"

    @bp_method.setter
    def bp_method(self, value: Union[str, int]) -> None:
        if str(value).lower() in ["prod_sum", "product_sum", "ps", "0", "prod sum"]:
            self._bp_method = _capi.PRODUCT_SUM
        elif str(value).lower() in ["min_sum", "minimum_sum", "ms", "1", "minimum sum", "min sum"]:
            self._bp_method = _capi.MINIMUM_SUM
        else:
            raise ValueError(f"BP method '{value}' is invalid. \
                    Please choose from the following methods: \
                    'product_sum', 'minimum_sum'")
        self._dirty = True

    @property
    def schedule(self) -> str:
        return {_capi.PARALLEL: "parallel", _capi.SERIAL: "serial", _capi.SERIAL_RELATIVE: "serial_relative"}[
            self._schedule]

    @schedule.setter
    def schedule(self, value: Union[str, int]) -> None:
        if str(value).lower() in ["parallel", "p", "0"]:
            self._schedule = _capi.PARALLEL
        elif str(value).lower() in ["serial", "s", "1"]:
            self._schedule = _capi.SERIAL
        elif str(value).lower() in ["serial_relative", "sr", "2"]:
            self._schedule = _capi.SERIAL_RELATIVE
        else:
            raise ValueError(f"The BP schedule method '{value}' is invalid. \
                    Please choose from the following methods: \
                    'schedule=parallel', 'schedule=serial', 'schedule=serial_relative'")
        self._dirty = True

    @property
    def serial_schedule_order(self) -> Union[None, np.ndarray]:
        if self._serial_schedule_order is None or len(self._serial_schedule_order) == 0:
            return None
        return np.asarray(self._serial_schedule_order).astype(int).copy()

    @serial_schedule_order.setter
    def serial_schedule_order(self, value) -> None:
        if value is None:
            return
        if not len(value) == self.n:
            raise Exception("Input error. The `serial_schedule_order` input parameter must have length equal to the length of the code.")
        for i in range(self.n):
            if not isinstance(value[i], (int, np.int64, np.int32)) or value[i] < 0 or value[i] >= self.n:
                raise ValueError(f"serial_schedule_order[{i}] is invalid. It must be a non-negative integer less than {self.n}.")
        self._serial_schedule_order = np.asarray(value, dtype=np.int64).copy()
        self.random_serial_schedule = False
        self._dirty = True

    @property
    def ms_scaling_factor(self) -> float:
        return self._ms_scaling_factor

    @ms_scaling_factor.setter
    def ms_scaling_factor(self, value: float) -> None:
        if not isinstance(value, (float, int)):
            raise TypeError("The ms_scaling factor must be specified as a float")
        self._ms_scaling_factor = float(value)
        self._dirty = True

    @property
    def omp_thread_count(self) -> int:
        if self._omp_thread_count != 1:
            warnings.warn("The OpenMP functionality is not yet implemented")
        return self._omp_thread_count

    @omp_thread_count.setter
    def omp_thread_count(self, value: int) -> None:
        if not isinstance(value, int) or value < 1:
            raise TypeError("The omp_thread_count must be specified as a\
            positive integer.")
        self._omp_thread_count = value
        if self._omp_thread_count != 1:
            warnings.warn("The OpenMP functionality is not yet implemented")

    @property
    def random_schedule_seed(self) -> int:
        return self._random_schedule_seed

    @random_schedule_seed.setter
    def random_schedule_seed(self, value: int) -> None:
        if not isinstance(value, int) or value < -2:
            raise ValueError("The value of random_schedule_seed must\
            be a positive integer. Set as -1 to disable to the random\
            schedule. Set as 0 to use the system clock.")
        self._random_serial_schedule = True  # reference quirk, _bp_decoder.pyx:553
        self._random_schedule_seed = value
        self._dirty = True

    @property
    def random_serial_schedule(self) -> bool:
        return self._random_serial_schedule

    @random_serial_schedule.setter
    def random_serial_schedule(self, value: bool) -> None:
        self._random_serial_schedule = bool(value)
        self._dirty = True


_BP_KWARGS = ("channel_probs", "device", "devices", "kernel")


class BpDecoder(BpDecoderBase):
    """Belief propagation decoder for binary linear codes (reference ``BpDecoder``, _bp_decoder.pyx:581-709).

    Parameters are the reference's; ``device`` (CUDA ordinal), ``devices`` (list of ordinals: every batch call is split over
    them inside the library) and ``kernel`` ('auto' | 'stream' | 'smem' | 'pair' | 'edge') are additions.  ``decode`` takes one syndrome (or received vector); ``decode_batch`` takes ``[B, m]``.
    """

    def __init__(self, pcm, error_rate: Optional[float] = None, error_channel=None, max_iter: Optional[int] = 0,
                 bp_method=0, ms_scaling_factor=1.0, schedule=0, omp_thread_count: Optional[int] = 1,
                 random_schedule_seed: Optional[int] = 0, serial_schedule_order: Optional[List[int]] = None,
                 input_vector_type: str = "auto", random_serial_schedule: bool = False, **kwargs):
        for key in kwargs.keys():
            if key not in _BP_KWARGS:
                raise ValueError(f"Unknown parameter '{key}' passed to the BpDecoder constructor.")
        super().__init__(pcm, error_rate=error_rate, error_channel=error_channel, max_iter=max_iter,
                         bp_method=bp_method, ms_scaling_factor=ms_scaling_factor, schedule=schedule,
                         omp_thread_count=omp_thread_count, random_schedule_seed=random_schedule_seed,
                         serial_schedule_order=serial_schedule_order, random_serial_schedule=random_serial_schedule,
                         **kwargs)
        self.input_vector_type = input_vector_type

    def _resolve_input(self, length: int) -> int:
        t = self._input_type
        if t == _capi.INPUT_SYNDROME and not length == self.m:
            raise ValueError(f"The input_vector must have length {self.m} (for syndrome decoding). Not length {length}.")
        elif t == _capi.INPUT_RECEIVED_VECTOR and not length == self.n:
            raise ValueError(f"The input_vector must have length {self.n} (for received vector decoding). Not length {length}.")
        elif t == _capi.INPUT_AUTO and not (length == self.m or length == self.n):
            raise ValueError(f"The input_vector must have length {self.m} (for syndrome decoding) or length {self.n} (for received vector decoding). Not length {length}.")
        if t == _capi.INPUT_SYNDROME or (t == _capi.INPUT_AUTO and length == self.m):
            return _capi.INPUT_SYNDROME
        return _capi.INPUT_RECEIVED_VECTOR

    def decode(self, input_vector: np.ndarray) -> np.ndarray:
        """Decode one syndrome / received vector (reference ``BpDecoder.decode``, _bp_decoder.pyx:642-695)."""
        input_vector = np.asarray(input_vector)
        kind = self._resolve_input(len(input_vector))
        dtype = input_vector.dtype
        vec = np.ascontiguousarray(input_vector.astype(np.uint8, copy=False)).reshape(1, -1)
        if not vec.any():
            # zero shortcut: converge=True, iter / log_prob_ratios keep their previous values (:679-681)
            self._converge = True
            return np.zeros(self.n, dtype=dtype)
        dec, conv, its, llr = self._decode_device_batch(vec, kind, want_llr=True)
        if self._schedule == _capi.SERIAL_RELATIVE and self._devices is None:
            # the reference sorts its serial_schedule_order member in place in every iteration (bp.hpp:469-482) and
            # the object carries it into its next decode: mirror that state (decode_batch does not: every row of a
            # batch starts from the configured order)
            order = self._native.last_schedule_order(len(self._serial_schedule_order))
            self._serial_schedule_order = order.astype(np.int64)
            self._native.set_order(order)
        self._decoding = dec[0]
        self._converge = bool(conv[0])
        self._iterations = int(its[0])
        self._log_prob_ratios = llr[0]
        return dec[0].astype(dtype)

    def decode_batch(self, input_vectors: np.ndarray, return_llr: bool = False, out=None) -> np.ndarray:
        """Decode ``[B, m]`` syndromes (or ``[B, n]`` received vectors) in one GPU call.

        ``out``: optional preallocated uint8 ``[B, n]`` result array (numpy style; pinned memory from
        ``ldpc_b200._capi.PinnedArray`` makes the device-to-host copies asynchronous).

        Every row is decoded exactly as ``BpDecoder::decode`` would decode it on its own (reference
        src_cpp/bp.hpp:159-190), i.e. with the C++ semantics: an all-zero syndrome runs one iteration and
        converges with ``iter == 1``.  Returns ``[B, n]`` in the input dtype and fills ``converge_batch``,
        ``iter_batch`` and (if requested) ``log_prob_ratios_batch``.
        """
        arr = np.asarray(input_vectors)
        if arr.ndim != 2:
            raise ValueError("decode_batch expects a 2-D array [batch, length]")
        kind = self._resolve_input(arr.shape[1])
        dtype = arr.dtype
        vec = np.ascontiguousarray(arr.astype(np.uint8, copy=False))
        if vec.shape[0] == 0:
            self.converge_batch = np.zeros(0, bool)
            self.iter_batch = np.zeros(0, np.int32)
            self.log_prob_ratios_batch = np.zeros((0, self.n)) if return_llr else None
            return np.zeros((0, self.n), dtype=dtype)
        dec, conv, its, llr = self._decode_device_batch(vec, kind, want_llr=return_llr, out=out)
        self.converge_batch, self.iter_batch, self.log_prob_ratios_batch = conv, its, llr
        return dec if dtype == np.uint8 else dec.astype(dtype)

    @property
    def decoding(self) -> np.ndarray:
        return self._decoding.astype(int)


class SoftInfoBpDecoder(BpDecoderBase):
    """Soft-information belief propagation (reference ``SoftInfoBpDecoder``, _bp_decoder.pyx:712-812, on top of
    ``BpDecoder::soft_info_decode_serial``, src_cpp/bp.hpp:547-665): serial-schedule min-sum on a real-valued
    syndrome; checks whose soft magnitude ``2 s_i / sigma^2`` is below ``cutoff`` take part as virtual variable nodes.
    Same constructor as the reference (``cutoff``, ``sigma`` on top of the base keywords; schedule, method and input
    type are forced to serial / minimum_sum / syndrome, :751-753).  ``decode(soft_syndrome)`` as in the reference;
    ``decode_batch(soft_syndromes [B, m])`` decodes a batch in one GPU call."""

    def __init__(self, pcm, error_rate: Optional[float] = None, error_channel: Optional[List[float]] = None,
                 max_iter: Optional[int] = 0, bp_method: Optional[str] = "minimum_sum",
                 ms_scaling_factor: Optional[float] = 1.0, cutoff: Optional[float] = np.inf, sigma: float = 2.0,
                 **kwargs):
        super().__init__(pcm, error_rate=error_rate, error_channel=error_channel, max_iter=max_iter,
                         bp_method=bp_method, ms_scaling_factor=ms_scaling_factor, **kwargs)
        self.cutoff = cutoff
        if not isinstance(sigma, float) or sigma <= 0:
            raise ValueError("The sigma value must be a float greater than 0.")
        self.sigma = sigma
        self.schedule = "serial"
        self.bp_method = "minimum_sum"
        self.input_vector_type = "syndrome"
        self._soft_syndrome = np.zeros(self.m)
        self.soft_syndrome_batch = None

    def _soft_batch(self, soft: np.ndarray, want_llr: bool):
        self._ensure_handle()
        B = soft.shape[0]
        dec = np.empty((B, self.n), dtype=np.uint8)
        conv = np.empty((B,), dtype=np.uint8)
        its = np.empty((B,), dtype=np.int32)
        llr = np.empty((B, self.n), dtype=np.float64) if want_llr else None
        out = np.empty((B, self.m), dtype=np.float64)
        self._native.soft_info_decode_batch(soft, self.cutoff, self.sigma, dec, conv, its, llr, out)
        return dec, conv.astype(bool), its, llr, out

    def decode(self, soft_info_syndrome: np.ndarray) -> np.ndarray:
        soft = np.ascontiguousarray(np.asarray(soft_info_syndrome, dtype=np.float64).reshape(1, -1))
        if soft.shape[1] != self.m:
            raise ValueError(f"The soft syndrome must have length {self.m}. Not {soft.shape[1]}.")
        dec, conv, its, llr, out = self._soft_batch(soft, True)
        self._decoding = dec[0]
        self._converge = bool(conv[0])
        self._iterations = int(its[0])
        self._log_prob_ratios = llr[0]
        self._soft_syndrome = out[0]
        return dec[0].copy()

    def decode_batch(self, soft_info_syndromes: np.ndarray, return_llr: bool = False) -> np.ndarray:
        soft = np.ascontiguousarray(np.asarray(soft_info_syndromes, dtype=np.float64))
        if soft.ndim != 2 or soft.shape[1] != self.m:
            raise ValueError(f"The soft syndromes must have shape [batch, {self.m}].")
        if soft.shape[0] == 0:
            self.converge_batch, self.iter_batch = np.zeros(0, bool), np.zeros(0, np.int32)
            return np.zeros((0, self.n), np.uint8)
        dec, conv, its, llr, out = self._soft_batch(soft, return_llr)
        self.converge_batch, self.iter_batch, self.log_prob_ratios_batch = conv, its, llr
        self.soft_syndrome_batch = out
        return dec

    @property
    def soft_syndrome(self) -> np.ndarray:
        return np.asarray(self._soft_syndrome, dtype=float).copy()

    @property
    def decoding(self) -> np.ndarray:
        return np.asarray(self._decoding).astype(int)

