"""Batched Monte-Carlo driver for the binary symmetric channel.

Same constructor, attributes and result dictionary as the reference's ``MonteCarloBscSimulation``
(reference src_python/ldpc/monte_carlo_simulation/mcs.py:10-159), which loops
``error -> syndrome -> Decoder.decode(syndrome) -> compare`` one run at a time (mcs.py:124-149).  Here the runs
are generated and decoded ``batch_size`` at a time through ``Decoder.decode_batch`` (the
natural producer of large syndrome batches).  With the same ``seed`` the errors are the very ones the reference
draws (``np.random.binomial`` fills a ``[batch, n]`` array in the order consecutive length-``n`` calls would), so
``fail_count`` equals the reference's whenever the decoder reproduces the reference's decodings.
"""
from __future__ import annotations

import datetime
import time
from typing import Dict, Union

import numpy as np
import scipy.sparse as sp


class MonteCarloBscSimulation:
    def __init__(self, parity_check_matrix: Union[np.ndarray, sp.csr_matrix] = None, error_rate: float = None,
                 Decoder=None, target_run_count=1000, tqdm_disable=False, save_interval=60, seed=None, run=False,
                 batch_size: int = 1 << 16, device_side: bool = False) -> None:
        if parity_check_matrix is None or not isinstance(parity_check_matrix, (np.ndarray, sp.csr_matrix)):
            raise ValueError(
                f"parity_check_matrix should be of type np.ndarray or scipy.sparse.csr_matrix. Not {type(parity_check_matrix)}")
        self.parity_check_matrix = parity_check_matrix
        if error_rate is None or not isinstance(error_rate, float) or error_rate < 0 or error_rate > 1:
            raise ValueError("Invalid error rate provided. The error rate should be a float with value between 0 and 1.")
        self.error_rate = error_rate
        if Decoder is None:
            raise ValueError("Invalid Decoder object provided.")
        self.Decoder = Decoder
        if not isinstance(target_run_count, int) or target_run_count <= 0:
            raise ValueError("Invalid target run count provided.")
        self.target_run_count = target_run_count
        if not isinstance(tqdm_disable, bool):
            raise ValueError("Invalid value for tqdm_disable flag.")
        self.tqdm_disable = tqdm_disable
        if not isinstance(save_interval, int) or save_interval <= 0:
            raise ValueError("Invalid save interval provided.")
        self.save_interval = save_interval
        if not isinstance(batch_size, int) or batch_size <= 0:
            raise ValueError("Invalid batch size provided.")
        self.batch_size = batch_size
        # device_side: errors are drawn, turned into syndromes, decoded and scored on the GPU (Decoder.monte_carlo_bsc,
        # Philox4x32-10 keyed by `seed`); only counters come back.  The drawn errors then differ from numpy's.
        self.device_side = bool(device_side)
        if seed is None:
            self.seed = None
        else:
            if not isinstance(seed, int):
                raise ValueError("Invalid seed provided. Please provide a postive integer")
            self.seed = seed
            np.random.seed(self.seed)
        self.run_count = 0
        self.fail_count = 0
        self.logical_error_rate = 0.0
        self.logical_error_rate_eb = 0.0
        if run:
            self.run()

    def run(self) -> Dict:
        self.start_date = datetime.datetime.fromtimestamp(time.time()).strftime("%A, %B %d, %Y %H:%M:%S")
        H = sp.csr_matrix(self.parity_check_matrix, dtype=np.int32)
        n = H.shape[1]
        self.fail_count = 0
        done = self.run_count
        if self.device_side:
            r = self.Decoder.monte_carlo_bsc(self.target_run_count - done, seed=0 if self.seed is None else self.seed,
                                             error_rate=self.error_rate, first_run=done,
                                             with_osd=hasattr(self.Decoder, "osd_method"))
            self.fail_count = r["fail_count"]
            self.run_count = done + r["run_count"]
            self.logical_error_rate = self.fail_count / self.run_count
            self.logical_error_rate_eb = np.sqrt(
                self.logical_error_rate * (1 - self.logical_error_rate) / self.run_count)
            return self.save()
        while done < self.target_run_count:
            nb = min(self.batch_size, self.target_run_count - done)
            # the reference's generate_bsc_error (noise_models/bsc.py:23), nb runs at once
            errors = np.random.binomial(1, self.error_rate, (nb, n)).astype(np.uint8)
            syndromes = np.ascontiguousarray((H @ errors.T.astype(np.int32)).T % 2).astype(np.uint8)
            decodings = self.Decoder.decode_batch(syndromes)
            self.fail_count += int(np.count_nonzero((np.asarray(decodings) != errors).any(axis=1)))
            done += nb
            self.run_count = done
            self.logical_error_rate = self.fail_count / self.run_count
            self.logical_error_rate_eb = np.sqrt(
                self.logical_error_rate * (1 - self.logical_error_rate) / self.run_count)
        return self.save()

    def save(self):
        return {"logical_error_rate": self.logical_error_rate, "logical_error_rate_eb": self.logical_error_rate_eb,
                "error_rate": self.error_rate, "run_count": self.run_count, "fail_count": self.fail_count}
