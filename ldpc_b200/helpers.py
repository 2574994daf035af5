"""Input validation for parity-check matrices.

Same contract as the reference's ``ldpc.helpers.scipy_helpers.convert_to_binary_sparse``
(reference src_python/ldpc/helpers/scipy_helpers.py:6-68, pinned by python_test/test_scipy_helpers.py):
accepts numpy arrays and scipy sparse matrices of dtype uint8 / int8 / int / float whose entries are all
0 or 1, and returns a scipy sparse matrix without explicit zeros.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import scipy.sparse

_ALLOWED_DTYPES = (np.dtype(np.uint8), np.dtype(np.int8), np.dtype(int), np.dtype(float))


def convert_to_binary_sparse(matrix: Union[np.ndarray, scipy.sparse.spmatrix]) -> scipy.sparse.spmatrix:
    if not isinstance(matrix, (np.ndarray, scipy.sparse.spmatrix)):
        raise TypeError(f"Input must be a binary numpy array or scipy sparse matrix, not {type(matrix)}")
    if np.dtype(matrix.dtype) not in _ALLOWED_DTYPES:
        raise TypeError(f"Input matrix must have dtype uint8, int8, or int, not {matrix.dtype}")
    values = matrix if isinstance(matrix, np.ndarray) else matrix.data
    if values.size and not np.all((values == 0) | (values == 1)):
        raise ValueError("Input matrix must be a binary matrix.")
    if isinstance(matrix, np.ndarray):
        out = scipy.sparse.csr_matrix(matrix, dtype=np.uint8)
    else:
        out = matrix.astype(np.uint8) if matrix.dtype == float else matrix
    out.eliminate_zeros()
    return out
