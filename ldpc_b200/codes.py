"""Parity-check matrix constructors and syndrome generators for the BASELINE configs.

These mirror the reference's fixtures (``ldpc.codes.rep_code / ring_code / hamming_code``,
reference ``src_python/ldpc/codes/rep_code.py``, ``hamming_code.py``; BSC noise model
``src_python/ldpc/noise_models/bsc.py:4-23``) and add the codes ``BASELINE.json`` names that the
reference does not ship: the (3,6)-regular random LDPC, the rotated surface code X checks and the
[[144,12,12]] bivariate-bicycle code.  Host-only numpy/scipy; nothing here is
on the GPU hot path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def rep_code(distance: int) -> sp.csr_matrix:
    """(d-1) x d repetition-code checks: row i touches bits i and i+1."""
    if distance < 2:
        raise ValueError("Distance should be greater than or equal to 2.")
    i = np.arange(distance - 1)
    rows = np.repeat(i, 2)
    cols = np.stack([i, i + 1], axis=1).ravel()
    return sp.csr_matrix((np.ones(rows.size, np.uint8), (rows, cols)), shape=(distance - 1, distance), dtype=np.uint8)


def ring_code(distance: int) -> sp.csr_matrix:
    """d x d closed-loop repetition code."""
    if distance < 2:
        raise ValueError("Distance should be greater than or equal to 2.")
    i = np.arange(distance)
    rows = np.repeat(i, 2)
    cols = np.stack([i, (i + 1) % distance], axis=1).ravel()
    return sp.csr_matrix((np.ones(rows.size, np.uint8), (rows, cols)), shape=(distance, distance), dtype=np.uint8)


def hamming_code(rank: int) -> sp.csr_matrix:
    """rank x (2^rank - 1) Hamming checks: column i is the binary expansion of i+1, MSB in row 0."""
    if not isinstance(rank, int):
        raise TypeError("The input variable 'rank' must be of type 'int'.")
    n = (1 << rank) - 1
    vals = np.arange(1, n + 1)
    dense = ((vals[None, :] >> (rank - 1 - np.arange(rank))[:, None]) & 1).astype(np.uint8)
    return sp.csr_matrix(dense)


def regular_ldpc(n: int, dv: int = 3, dc: int = 6, seed: int = 1, max_tries: int = 10000) -> sp.csr_matrix:
    """(dv,dc)-regular LDPC from the configuration model.

    ``repeat(arange(n), dv)`` is shuffled into ``n*dv/dc`` rows of ``dc`` sockets; rows that received
    the same bit twice are repaired by swapping sockets with random other rows until the graph is simple.
    """
    if (n * dv) % dc:
        raise ValueError("n*dv must be divisible by dc")
    m = n * dv // dc
    rng = np.random.default_rng(seed)
    sockets = np.repeat(np.arange(n), dv)
    rng.shuffle(sockets)
    table = sockets.reshape(m, dc)
    for _ in range(max_tries):
        srt = np.sort(table, axis=1)
        bad = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(axis=1))[0]
        if bad.size == 0:
            break
        for r in bad:
            row = table[r]
            _, first = np.unique(row, return_index=True)
            dup = np.setdiff1d(np.arange(dc), first)
            for k in dup:
                r2 = int(rng.integers(m))
                k2 = int(rng.integers(dc))
                table[r, k], table[r2, k2] = table[r2, k2], table[r, k]
    else:  # pragma: no cover
        raise RuntimeError("could not build a simple regular graph")
    rows = np.repeat(np.arange(m), dc)
    H = sp.csr_matrix((np.ones(m * dc, np.uint8), (rows, table.ravel())), shape=(m, n), dtype=np.uint8)
    H.sum_duplicates()
    assert H.nnz == m * dc and H.data.max() == 1
    H.sort_indices()
    return H


def rotated_surface_code_x(d: int) -> sp.csr_matrix:
    """X-type check matrix of the distance-d rotated surface code: (d^2-1)/2 x d^2.

    Data qubits on a d x d grid (index r*d + c).  Bulk plaquettes (r,c), 0<=r,c<d-1 touch qubits
    (r,c),(r,c+1),(r+1,c),(r+1,c+1); X-type ones are those with (r+c) even.  Weight-2 X checks sit on
    the top and bottom boundaries where the bulk colouring leaves a gap.
    """
    if d < 3 or d % 2 == 0:
        raise ValueError("d must be odd and >= 3")
    checks = []
    for r in range(d - 1):
        for c in range(d - 1):
            if (r + c) % 2 == 0:
                checks.append([r * d + c, r * d + c + 1, (r + 1) * d + c, (r + 1) * d + c + 1])
    for c in range(d - 1):
        # top boundary: the virtual plaquette (-1, c) is X-type when (-1+c) is even
        if (c - 1) % 2 == 0:
            checks.append([c, c + 1])
        # bottom boundary: virtual plaquette (d-1, c)
        if (d - 1 + c) % 2 == 0:
            checks.append([(d - 1) * d + c, (d - 1) * d + c + 1])
    rows = np.concatenate([np.full(len(q), i) for i, q in enumerate(checks)])
    cols = np.concatenate([np.asarray(q) for q in checks])
    m = len(checks)
    H = sp.csr_matrix((np.ones(rows.size, np.uint8), (rows, cols)), shape=(m, d * d), dtype=np.uint8)
    H.sort_indices()
    assert m == (d * d - 1) // 2
    return H


def _cyclic_shift(size: int, power: int) -> np.ndarray:
    return np.roll(np.eye(size, dtype=np.int64), power % size, axis=1)


def bivariate_bicycle_144() -> sp.csr_matrix:
    """H_X = [A | B] of the [[144,12,12]] bivariate-bicycle code (l=12, m=6,
    A = x^3 + y + y^2, B = y^3 + x + x^2; x = S_12 (x) I_6, y = I_12 (x) S_6): 72 x 144, row weight 6."""
    l, m = 12, 6
    x = {k: np.kron(_cyclic_shift(l, k), np.eye(m, dtype=np.int64)) for k in range(1, 4)}
    y = {k: np.kron(np.eye(l, dtype=np.int64), _cyclic_shift(m, k)) for k in range(1, 4)}
    A = (x[3] + y[1] + y[2]) % 2
    B = (y[3] + x[1] + x[2]) % 2
    H = sp.csr_matrix(np.hstack([A, B]).astype(np.uint8))
    H.sort_indices()
    return H


def bsc_errors(n: int, error_rate: float, batch: int, seed: int = 7) -> np.ndarray:
    """[batch, n] uint8 i.i.d. Bernoulli(error_rate) errors (the reference's BSC, noise_models/bsc.py:23)."""
    rng = np.random.default_rng(seed)
    return (rng.random((batch, n)) < error_rate).astype(np.uint8)


def syndromes_of(H: sp.spmatrix, errors: np.ndarray) -> np.ndarray:
    """[batch, m] uint8 syndromes s = H e mod 2 (reference GF2Sparse::mulvec, gf2sparse.hpp:177-214)."""
    Hc = sp.csr_matrix(H, dtype=np.int32)
    return np.ascontiguousarray((Hc @ errors.T.astype(np.int32)).T % 2).astype(np.uint8)


def bsc_syndromes(H: sp.spmatrix, error_rate: float, batch: int, seed: int = 7, chunk: int = 1 << 16):
    """Seeded BSC syndromes for a batch, generated chunk-wise; returns [batch, m] uint8."""
    m, n = H.shape
    out = np.empty((batch, m), np.uint8)
    rng = np.random.default_rng(seed)
    Hc = sp.csr_matrix(H, dtype=np.float32)
    for lo in range(0, batch, chunk):
        hi = min(batch, lo + chunk)
        e = (rng.random((hi - lo, n), dtype=np.float32) < error_rate).astype(np.float32)
        out[lo:hi] = (np.asarray((Hc @ e.T).T) % 2).astype(np.uint8)
    return out
