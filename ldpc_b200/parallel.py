"""Sharding a syndrome batch over the GPUs of one box.

The decode of one syndrome never touches another (reference src_cpp/bp.hpp:192-325 re-initialises all messages
per call), so the batch splits into contiguous shards with no collective on the data path (BASELINE.json
north_star: "host-side split and final gather, NCCL not on the hot path").  Two ways to drive it:

* ``MultiGpuBpDecoder``  -- one process, ONE C-ABI handle split over the devices by ``bpb_set_devices``: a host
  thread and a chunked H2D | kernels | D2H pipeline per device, results written into disjoint slices of one host
  output array (that *is* the gather).
* ``decode_sharded``     -- one process per GPU (torchrun): every rank decodes its own shard; rank 0 can collect
  the pieces with ``torch.distributed`` (gloo or nccl) after the timed region.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``total`` items for ``rank`` of ``world`` (sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(int(total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class MultiGpuBpDecoder:
    """Host-side split of one batch over several devices of one box, done INSIDE the library: one C-ABI handle with
    ``bpb_set_devices`` (one full decoder, pinned-speed staging and host thread per device; contiguous input slices,
    outputs copied into disjoint ranges of one host array).  Equivalent to ``BpDecoder(pcm, devices=[...])``."""

    def __init__(self, pcm, devices: Sequence[int], **kwargs):
        from .bp_decoder import BpDecoder
        if len(devices) < 1:
            raise ValueError("need at least one device")
        self.devices = [int(d) for d in devices]
        self.decoder = BpDecoder(pcm, devices=self.devices, **kwargs)
        self.n = self.decoder.n
        self.converge_batch = None
        self.iter_batch = None

    def decode_batch(self, syndromes: np.ndarray) -> np.ndarray:
        out = self.decoder.decode_batch(syndromes)
        self.converge_batch, self.iter_batch = self.decoder.converge_batch, self.decoder.iter_batch
        return out

    def monte_carlo_bsc(self, runs: int, **kw) -> dict:
        return self.decoder.monte_carlo_bsc(runs, **kw)


def decode_sharded(decode_fn: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray, np.ndarray]],
                   syndromes: np.ndarray, gather_to: Optional[int] = 0, group=None):
    """One-process-per-GPU driver: ``decode_fn(shard) -> (decoding, converged, iters)`` runs on this rank's shard of
    ``syndromes`` (every rank passes the same full array or an array of the same length); if ``gather_to`` is a
    rank, that rank returns the full results in batch order and the others return their shard only."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = int(syndromes.shape[0])
    lo, hi = shard_bounds(B, world, rank)
    dec, conv, its = decode_fn(syndromes[lo:hi])
    if world == 1 or gather_to is None:
        return dec, conv, its
    n = dec.shape[1] if dec.ndim == 2 else 0
    n_t = torch.tensor([n], dtype=torch.int64)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n_t = n_t.to(dev)
    dist.all_reduce(n_t, op=dist.ReduceOp.MAX, group=group)
    n = int(n_t.item())
    cap = max(shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world))
    pad_dec = torch.zeros((cap, n), dtype=torch.uint8)
    pad_meta = torch.zeros((cap, 2), dtype=torch.int32)
    if hi > lo:
        pad_dec[: hi - lo] = torch.from_numpy(np.ascontiguousarray(dec).reshape(hi - lo, n))
        pad_meta[: hi - lo, 0] = torch.from_numpy(np.asarray(conv).astype(np.int32))
        pad_meta[: hi - lo, 1] = torch.from_numpy(np.asarray(its).astype(np.int32))
    pad_dec, pad_meta = pad_dec.to(dev), pad_meta.to(dev)
    all_dec = [torch.empty_like(pad_dec) for _ in range(world)]
    all_meta = [torch.empty_like(pad_meta) for _ in range(world)]
    dist.all_gather(all_dec, pad_dec, group=group)
    dist.all_gather(all_meta, pad_meta, group=group)
    if rank != gather_to:
        return dec, conv, its
    full_dec = np.empty((B, n), np.uint8)
    full_conv = np.empty(B, bool)
    full_its = np.empty(B, np.int32)
    for r in range(world):
        a, b = shard_bounds(B, world, r)
        full_dec[a:b] = all_dec[r][: b - a].cpu().numpy()
        meta = all_meta[r][: b - a].cpu().numpy()
        full_conv[a:b] = meta[:, 0] != 0
        full_its[a:b] = meta[:, 1]
    return full_dec, full_conv, full_its
