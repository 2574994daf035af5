"""ctypes binding of the C-ABI in include/bp_b200.h (ldpc_b200/libbp_b200.so).

This is the Python side of the drop-in boundary: what the reference reaches through its Cython
``cdef extern`` block (src_python/ldpc/bp_decoder/_bp_decoder.pxd:9-83) is reached here through plain C
functions.  There is NO CPU fallback: if the CUDA library is missing or no GPU is visible, loading or
creating a decoder raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(_HERE, "libbp_b200.so")  # env var: experimental builds

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)

OK = 0
PRODUCT_SUM, MINIMUM_SUM = 0, 1          # reference bp.hpp:23-26
SERIAL, PARALLEL, SERIAL_RELATIVE = 0, 1, 2  # reference bp.hpp:28-32
INPUT_SYNDROME, INPUT_RECEIVED_VECTOR, INPUT_AUTO = 0, 1, 2  # reference bp.hpp:34-38
KERNEL_AUTO, KERNEL_STREAM, KERNEL_SMEM, KERNEL_EDGE, KERNEL_PAIR = 0, 1, 2, 3, 4
OSD_AUTO, OSD_HOST, OSD_DEVICE = 0, 1, 2


class BpbInfo(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("nnz", C.c_int64), ("max_row_degree", C.c_int),
                ("max_col_degree", C.c_int), ("device", C.c_int), ("sm_count", C.c_int), ("kernel_family", C.c_int),
                ("grid", C.c_int), ("block", C.c_int), ("launches", C.c_int64), ("workspace_bytes", C.c_int64),
                ("last_kernel_ms", C.c_double), ("smem_family_available", C.c_int),
                ("smem_bank_multiplicity", C.c_int), ("smem_bytes_per_syndrome", C.c_int),
                ("stream_iterations", C.c_int64), ("stream_handed_off", C.c_int64),
                ("osd_device_available", C.c_int), ("osd_device_solved", C.c_int64), ("osd_host_solved", C.c_int64), ("osd_host_inconsistent", C.c_int64),
                ("pair_family_available", C.c_int), ("pair_bank_multiplicity", C.c_int)]


EXPORTS = [
    "bpb_create", "bpb_destroy", "bpb_last_error", "bpb_set_channel", "bpb_set_max_iter", "bpb_set_method",
    "bpb_set_schedule", "bpb_set_ms_scaling_factor", "bpb_set_serial_schedule_order", "bpb_set_kernel",
    "bpb_decode_batch", "bpb_decode_batch_device", "bpb_osd0_host", "bpb_bposd_decode_batch", "bpb_get_info", "bpb_host_alloc",
    "bpb_host_free", "bpb_version", "bpb_set_osd_location", "bpb_set_devices", "bpb_bposd_decode_batch_device",
    "bpb_mc_bsc", "bpb_get_last_schedule_order", "bpb_libm_selfcheck",
    "bpb_set_observables", "bpb_decode_batch_b8", "bpb_soft_info_decode_batch",
]

_lib = None


def lib():
    """Load libbp_b200.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `make -C ldpc_b200/csrc -j8` "
                          "(or __graft_entry__.build()).  ldpc_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.bpb_create.restype = C.c_int
    L.bpb_create.argtypes = [C.c_int, C.c_int, C.c_int64, _i32p, _i32p, C.c_int, C.POINTER(vp)]
    L.bpb_destroy.restype = None
    L.bpb_destroy.argtypes = [vp]
    L.bpb_last_error.restype = C.c_char_p
    L.bpb_last_error.argtypes = [vp]
    L.bpb_set_channel.restype = C.c_int
    L.bpb_set_channel.argtypes = [vp, _f64p, C.c_int]
    for name in ("bpb_set_max_iter", "bpb_set_method", "bpb_set_schedule", "bpb_set_kernel"):
        f = getattr(L, name)
        f.restype = C.c_int
        f.argtypes = [vp, C.c_int]
    L.bpb_set_ms_scaling_factor.restype = C.c_int
    L.bpb_set_ms_scaling_factor.argtypes = [vp, C.c_double]
    L.bpb_set_serial_schedule_order.restype = C.c_int
    L.bpb_set_serial_schedule_order.argtypes = [vp, _i32p, C.c_int]
    L.bpb_decode_batch.restype = C.c_int
    L.bpb_decode_batch.argtypes = [vp, C.c_int, vp, C.c_int64, vp, vp, vp, vp]
    L.bpb_decode_batch_device.restype = C.c_int
    L.bpb_decode_batch_device.argtypes = [vp, C.c_int, vp, C.c_int64, vp, vp, vp, vp, vp]
    L.bpb_osd0_host.restype = C.c_int
    L.bpb_osd0_host.argtypes = [vp, vp, vp, vp, C.c_int64, vp, C.c_int]
    L.bpb_bposd_decode_batch.restype = C.c_int
    L.bpb_bposd_decode_batch.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, C.c_int]
    L.bpb_get_info.restype = C.c_int
    L.bpb_get_info.argtypes = [vp, C.POINTER(BpbInfo)]
    L.bpb_host_alloc.restype = vp
    L.bpb_host_alloc.argtypes = [C.c_size_t]
    L.bpb_host_free.restype = None
    L.bpb_host_free.argtypes = [vp]
    L.bpb_version.restype = C.c_char_p
    L.bpb_set_osd_location.restype = C.c_int
    L.bpb_set_osd_location.argtypes = [vp, C.c_int]
    L.bpb_set_devices.restype = C.c_int
    L.bpb_set_devices.argtypes = [vp, _i32p, C.c_int]
    L.bpb_bposd_decode_batch_device.restype = C.c_int
    L.bpb_bposd_decode_batch_device.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp]
    L.bpb_set_observables.restype = C.c_int
    L.bpb_set_observables.argtypes = [vp, C.c_int, C.c_int64, _i32p, _i32p]
    L.bpb_decode_batch_b8.restype = C.c_int
    L.bpb_decode_batch_b8.argtypes = [vp, C.c_int, vp, C.c_int64, vp, vp, vp, vp]
    L.bpb_soft_info_decode_batch.restype = C.c_int
    L.bpb_soft_info_decode_batch.argtypes = [vp, vp, C.c_int64, C.c_double, C.c_double, vp, vp, vp, vp, vp]
    L.bpb_libm_selfcheck.restype = C.c_int
    L.bpb_libm_selfcheck.argtypes = [C.c_int]
    L.bpb_get_last_schedule_order.restype = C.c_int
    L.bpb_get_last_schedule_order.argtypes = [vp, _i32p, C.c_int]
    L.bpb_mc_bsc.restype = C.c_int
    L.bpb_mc_bsc.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, _f64p, C.c_int, C.POINTER(C.c_int64)]
    _lib = L
    return L


class BpbError(RuntimeError):
    pass


def check(handle, rc: int) -> None:
    if rc != OK:
        msg = lib().bpb_last_error(handle)
        raise BpbError(f"ldpc_b200 C-ABI error {rc}: {msg.decode() if msg else ''}")


def host_ptr(a: np.ndarray | None):
    return None if a is None else C.c_void_p(a.ctypes.data)


class _PinnedBlock:
    """Owner of one pinned allocation; goes back to the pool when the numpy arrays built on it are gone."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes

    def __del__(self):
        try:
            _pool_give(self.ptr, self.nbytes)
        except Exception:
            pass


_POOL = {}            # bucket bytes (power of two) -> [ptr, ...] free pinned blocks
_POOL_BYTES = [0]     # bytes currently allocated through the pool (free + in use)
_POOL_FREE = [0]      # bytes sitting free in the pool
_POOL_LIMIT = int(os.environ.get("LDPC_B200_PINNED_LIMIT", str(8 << 30)))   # allocated through the pool at most
_POOL_CACHE = int(os.environ.get("LDPC_B200_PINNED_CACHE", str(8 << 30)))   # kept free for reuse at most (pinning
# 5 GB takes about a second: a result buffer of the n = 10^4 workload must survive between calls)
_POOL_LOCK = threading.Lock()  # MultiGpuBpDecoder / user threads call decode_batch concurrently


def _pool_give(ptr, nbytes):
    with _POOL_LOCK:
        if _POOL_FREE[0] + nbytes <= _POOL_CACHE:
            _POOL.setdefault(nbytes, []).append(ptr)
            _POOL_FREE[0] += nbytes
            return
        _POOL_BYTES[0] -= nbytes
    lib().bpb_host_free(ptr)  # above the cache cap: really unpin


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (full-speed, asynchronous PCIe copies), recycled through a small pool of
    power-of-two sized blocks.  Falls back to ordinary memory when the pool limit is reached or pinning fails."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    need = max(4096, count * dtype.itemsize)
    # buckets: powers of two up to 64 MiB, multiples of 64 MiB above (a power of two would pin up to 2x the request)
    nbytes = 1 << (need - 1).bit_length() if need <= (64 << 20) else -(-need // (64 << 20)) * (64 << 20)
    ptr = None
    with _POOL_LOCK:
        free = _POOL.get(nbytes)
        if free:
            ptr = free.pop()
            _POOL_FREE[0] -= nbytes
        elif _POOL_BYTES[0] + nbytes > _POOL_LIMIT:
            return np.empty(shape, dtype=dtype)
        else:
            _POOL_BYTES[0] += nbytes
    if ptr is None:
        ptr = lib().bpb_host_alloc(nbytes)
        if not ptr:
            with _POOL_LOCK:
                _POOL_BYTES[0] -= nbytes
            return np.empty(shape, dtype=dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    buf._owner = _PinnedBlock(ptr, nbytes)  # keeps the block out of the pool while any view of `buf` lives
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


class PinnedArray:
    """A numpy array backed by page-locked host memory from bpb_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        self._ptr = lib().bpb_host_alloc(max(nbytes, 1))
        if not self._ptr:
            raise MemoryError("bpb_host_alloc failed")
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self.array = None
                lib().bpb_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass
