"""Batch driver for sinter's file protocol: b8 detection events in, b8 observable predictions out.

Mirrors ``ldpc.sinter_decoders.SinterBpOsdDecoder`` (reference
``src_python/ldpc/sinter_decoders/sinter_bposd_decoder.py:9-126``): same constructor, same ``decode_via_files`` keyword
signature.  The reference reads all shots, then calls ``BpOsdDecoder.decode`` once per shot and multiplies each
correction by the observables matrix (``:110-126``); here the shots go through ``BpOsdDecoder.decode_batch`` in one
call.

``stim`` / ``sinter`` are optional: when ``stim`` is importable it parses the detector error model, else the small
parser below reads the DEM text (``error``, ``detector``, ``logical_observable``, ``shift_detectors``, ``repeat``)
and builds the same matrices as the reference's ``detector_error_model_to_check_matrices(dem,
allow_undecomposed_hyperedges=True)`` (``src_python/ldpc/ckt_noise/dem_matrices.py:60-170``): one column per distinct
detector set, probabilities of repeated mechanisms combined as ``p <- p (1 - q) + q (1 - p)``, the last seen
observable set kept.  The b8 format (stim's ``doc/result_formats.md``): each shot is ``ceil(bits / 8)`` bytes, bit
``k`` of the shot is bit ``k % 8`` (little endian) of byte ``k // 8``.
"""
from __future__ import annotations

import pathlib
import re
from dataclasses import dataclass
from typing import Dict, FrozenSet, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp

from .bposd_decoder import BpOsdDecoder


# ------------------------------------------------------------------------------------------ b8 files
def read_b8(path, num_bits: int) -> np.ndarray:
    """``[shots, num_bits]`` uint8 (0/1) from a b8 file."""
    nbytes = (num_bits + 7) // 8
    raw = np.fromfile(str(path), dtype=np.uint8)
    if nbytes == 0:
        return np.zeros((0, 0), np.uint8)
    if raw.size % nbytes:
        raise ValueError(f"b8 file size {raw.size} is not a multiple of {nbytes} bytes per shot")
    bits = np.unpackbits(raw.reshape(-1, nbytes), axis=1, bitorder="little")
    return np.ascontiguousarray(bits[:, :num_bits])


def write_b8(path, bits: np.ndarray) -> None:
    """Write ``[shots, num_bits]`` 0/1 values as a b8 file."""
    b = np.asarray(bits).astype(np.uint8, copy=False)
    if b.ndim != 2:
        raise ValueError("bits must be [shots, num_bits]")
    packed = np.packbits(b, axis=1, bitorder="little") if b.shape[1] else np.zeros((b.shape[0], 0), np.uint8)
    packed.tofile(str(path))


# ------------------------------------------------------------------------------------------ detector error models
@dataclass
class DemMatrices:
    check_matrix: sp.csc_matrix        # [num_detectors, num_mechanisms]
    observables_matrix: sp.csc_matrix  # [num_observables, num_mechanisms]
    priors: np.ndarray                 # [num_mechanisms]


_TOKEN = re.compile(r"\s+")


def _flatten_dem_text(text: str) -> List[Tuple[float, List[List[int]], List[List[int]]]]:
    """``[(p, detector groups, observable groups)]`` of every ``error`` instruction with absolute detector indices
    (``repeat`` blocks unrolled, ``shift_detectors`` applied), plus the largest indices seen through declarations."""
    lines = [ln.split("#", 1)[0].strip() for ln in text.splitlines()]
    lines = [ln for ln in lines if ln]
    out: List[Tuple[float, List[List[int]], List[List[int]]]] = []
    state = {"shift": 0, "max_det": -1, "max_obs": -1}

    def run(block: List[str]) -> None:
        i = 0
        while i < len(block):
            ln = block[i]
            if ln.startswith("repeat"):
                count = int(ln.split()[1])
                depth, j = 1, i + 1
                while j < len(block) and depth:
                    if block[j].startswith("repeat"):
                        depth += 1
                    elif block[j] == "}":
                        depth -= 1
                    j += 1
                body = block[i + 1:j - 1]
                for _ in range(count):
                    run(body)
                i = j
                continue
            head, _, rest = ln.partition(" ")
            name, _, args = head.partition("(")
            args = args.rstrip(")")
            toks = [t for t in _TOKEN.split(rest.strip()) if t]
            if name == "error":
                dets: List[List[int]] = [[]]
                obs: List[List[int]] = [[]]
                for t in toks:
                    if t == "^":
                        dets.append([])
                        obs.append([])
                    elif t[0] == "D":
                        d = int(t[1:]) + state["shift"]
                        dets[-1].append(d)
                        state["max_det"] = max(state["max_det"], d)
                    elif t[0] == "L":
                        o = int(t[1:])
                        obs[-1].append(o)
                        state["max_obs"] = max(state["max_obs"], o)
                    else:
                        raise ValueError(f"unexpected target {t!r} in {ln!r}")
                out.append((float(args.split(",")[0]), dets, obs))
            elif name == "shift_detectors":
                state["shift"] += int(toks[0]) if toks else 0
            elif name == "detector":
                for t in toks:
                    if t[0] == "D":
                        state["max_det"] = max(state["max_det"], int(t[1:]) + state["shift"])
            elif name == "logical_observable":
                for t in toks:
                    if t[0] == "L":
                        state["max_obs"] = max(state["max_obs"], int(t[1:]))
            elif name == "}":
                pass
            else:
                raise NotImplementedError(f"DEM instruction {name!r}")
            i += 1

    run(lines)
    out.append((float("nan"), [[state["max_det"]]], [[state["max_obs"]]]))  # sentinel carrying the index ranges
    return out


def _set_xor(groups: List[List[int]]) -> FrozenSet[int]:
    acc: set = set()
    for g in groups:
        acc ^= set(g)
    return frozenset(acc)


def dem_text_to_matrices(text: str, num_detectors: Optional[int] = None,
                         num_observables: Optional[int] = None) -> DemMatrices:
    """The check / observables / priors triple of ``detector_error_model_to_check_matrices(dem,
    allow_undecomposed_hyperedges=True)`` (reference dem_matrices.py:85-170) from DEM text."""
    items = _flatten_dem_text(text)
    (_, [[max_det]], [[max_obs]]) = items.pop()
    ids: Dict[FrozenSet[int], int] = {}
    obs_of: Dict[int, FrozenSet[int]] = {}
    priors: Dict[int, float] = {}
    for p, dets, obs in items:
        key = _set_xor(dets)
        if key not in ids:
            ids[key] = len(ids)
            priors[ids[key]] = 0.0
        hid = ids[key]
        obs_of[hid] = _set_xor(obs)
        priors[hid] = priors[hid] * (1 - p) + p * (1 - priors[hid])
    nd = max_det + 1 if num_detectors is None else int(num_detectors)
    no = max_obs + 1 if num_observables is None else int(num_observables)
    k = len(ids)

    def build(cols: Dict[int, FrozenSet[int]], rows: int) -> sp.csc_matrix:
        r = [x for c in range(k) for x in sorted(cols.get(c, ()))]
        c = [c for c in range(k) for _ in cols.get(c, ())]
        return sp.csc_matrix((np.ones(len(r), np.uint8), (r, c)), shape=(rows, k))

    return DemMatrices(check_matrix=build({v: key for key, v in ids.items()}, nd),
                       observables_matrix=build(obs_of, no),
                       priors=np.array([priors[i] for i in range(k)], dtype=np.float64))


def _matrices_from_dem_file(dem_path, num_dets: int, num_obs: int) -> DemMatrices:
    try:  # the reference's own route when stim is installed
        import stim  # noqa: F401
        dem = stim.DetectorErrorModel.from_file(str(dem_path))
        return dem_text_to_matrices(str(dem.flattened()), dem.num_detectors, dem.num_observables)
    except ImportError:
        return dem_text_to_matrices(pathlib.Path(dem_path).read_text(), num_dets, num_obs)


# ------------------------------------------------------------------------------------------ the decoder
class SinterBpOsdDecoder:
    """BP+OSD decoder for sinter's ``decode_via_files`` protocol (reference sinter_bposd_decoder.py:9-126).
    Subclasses ``sinter.Decoder`` when sinter is importable (it is only a protocol marker there)."""

    def __init__(self, max_iter=0, bp_method="ms", ms_scaling_factor=0.625, schedule="parallel", omp_thread_count=1,
                 serial_schedule_order=None, osd_method="osd0", osd_order=0, device: int = 0):
        self.max_iter = max_iter
        self.bp_method = bp_method
        self.ms_scaling_factor = ms_scaling_factor
        self.schedule = schedule
        self.omp_thread_count = omp_thread_count
        self.serial_schedule_order = serial_schedule_order
        self.osd_method = osd_method
        self.osd_order = osd_order
        self.device = device
        self.matrices: Optional[DemMatrices] = None
        self.bposd: Optional[BpOsdDecoder] = None

    def load_matrices(self, matrices: DemMatrices) -> None:
        """Configure from ready-made matrices (what ``decode_via_files`` derives from the DEM file)."""
        self.matrices = matrices
        self.bposd = BpOsdDecoder(sp.csr_matrix(matrices.check_matrix), error_channel=list(matrices.priors),
                                  max_iter=self.max_iter, bp_method=self.bp_method,
                                  ms_scaling_factor=self.ms_scaling_factor, schedule=self.schedule,
                                  omp_thread_count=self.omp_thread_count,
                                  serial_schedule_order=self.serial_schedule_order, osd_method=self.osd_method,
                                  osd_order=self.osd_order, device=self.device)
        self._obs = sp.csr_matrix(matrices.observables_matrix, dtype=np.int32)
        self.bposd.set_observables(self._obs)

    def decode_via_files(self, *, num_shots: int, num_dets: int, num_obs: int, dem_path, dets_b8_in_path,
                         obs_predictions_b8_out_path, tmp_dir=None) -> None:
        self.load_matrices(_matrices_from_dem_file(dem_path, num_dets, num_obs))
        # the b8 rows go to the device as they are: unpacking, BP + OSD-0, the observable parities and the packing of
        # the predictions all happen there (bpb_decode_batch_b8); an all-zero shot decodes to the zero correction
        nbytes = (num_dets + 7) // 8
        raw = np.fromfile(str(dets_b8_in_path), dtype=np.uint8)
        if nbytes == 0 or raw.size != num_shots * nbytes:
            raise ValueError(f"expected {num_shots} shots of {nbytes} bytes, the file holds {raw.size} bytes")
        if self.bposd.m != num_dets or self._obs.shape[0] != num_obs:
            raise ValueError("the detector error model does not match num_dets / num_obs")
        from ._capi import BpbError
        try:
            pred = self.bposd.decode_batch_b8(raw.reshape(num_shots, nbytes), decoding=False, observables=True)
        except BpbError:
            # codes the device OSD-0 kernel cannot hold (m > 1024 ...): unpacked route with the host elimination
            pred = np.packbits(self.decode_shots(read_b8(dets_b8_in_path, num_dets)), axis=1, bitorder="little")
        np.ascontiguousarray(pred).tofile(str(obs_predictions_b8_out_path))

    def decode_shots(self, shots: np.ndarray) -> np.ndarray:
        """``[shots, num_dets]`` detection events -> ``[shots, num_obs]`` predicted observable flips."""
        shots = np.asarray(shots)
        out = np.zeros((shots.shape[0], self._obs.shape[0]), dtype=np.uint8)
        nz = shots.any(axis=1)  # the reference's per-shot decode returns zeros for an all-zero syndrome (:78-81 of
        if nz.any():            # _bposd_decoder.pyx): same here, and those shots need no GPU work
            corr = self.bposd.decode_batch(np.ascontiguousarray(shots[nz].astype(np.uint8, copy=False)))
            out[nz] = np.asarray((self._obs @ corr.T.astype(np.int32)).T % 2).astype(np.uint8)
        return out

    def decode(self, syndrome: np.ndarray) -> np.ndarray:
        corr = self.bposd.decode(syndrome)
        return (self.matrices.observables_matrix @ corr) % 2


try:  # pragma: no cover - sinter is not in this image
    import sinter as _sinter

    class SinterBpOsdDecoder(SinterBpOsdDecoder, _sinter.Decoder):  # type: ignore[no-redef]
        pass
except ImportError:
    pass
