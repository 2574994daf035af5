"""ldpc v1 syntax: ``bp_decoder`` and ``bposd_decoder``.

Thin subclasses that translate the old constructor arguments (``channel_probs``, v1 method names) to
``BpDecoder`` / ``BpOsdDecoder``, like the reference's ``_legacy_ldpc_v1`` wrappers (reference
src_python/ldpc/_legacy_ldpc_v1/_legacy_bp_decoder.py:6-133, _legacy_bposd_decoder.py:6-140; pinned by
python_test/test_bp_decoder.py:238-263: same logical error rate as the v2 class on 1000 seeded runs).
"""
from __future__ import annotations

import warnings

import numpy as np

from .bp_decoder import BpDecoder
from .bposd_decoder import BpOsdDecoder

_PS = ["prod_sum", "product_sum", "ps", "0", "prod sum"]
_MS = ["min_sum", "minimum_sum", "ms", "1", "minimum sum", "min sum"]


def _v1_method(bp_method):
    if str(bp_method).lower() in _PS:
        return "ps"
    if str(bp_method).lower() in _MS:
        return "ms"
    raise ValueError(f"BP method '{bp_method}' is invalid.\
                            Please choose from the following methods:'product_sum',\
                            'minimum_sum'")


def _v1_channel(parity_check_matrix, error_rate, channel_probs):
    n = parity_check_matrix.shape[1]
    if channel_probs[0] is not None:
        if len(channel_probs) != n:
            raise ValueError(f"The length of the channel probability vector must be eqaul to the block length n={n}.")
        return np.asarray([channel_probs[j] for j in range(n)], dtype=float)
    if error_rate != 0:
        return None
    raise ValueError("Either the error_rate or channel_probs must be specified.")


class bp_decoder(BpDecoder):
    """Legacy ldpc_v1 class: a belief propagation decoder for LDPC codes (v1 argument names)."""

    def __init__(self, parity_check_matrix, error_rate=None, max_iter=0, bp_method="ps", ms_scaling_factor=1.0,
                 channel_probs=[None], input_vector_type="auto", error_channel=None, **kwargs):
        warnings.warn("This is the old syntax for the `bp_decoder` from `ldpc v1`. Use the `BpDecoder` class from "
                      "`ldpc v2` for additional features.")
        error_channel = _v1_channel(parity_check_matrix, error_rate, channel_probs)
        if type(input_vector_type) is int and input_vector_type == -1:
            input_vector_type = "auto"
        elif type(input_vector_type) is str and input_vector_type in ("auto", "syndrome", "received_vector"):
            pass
        else:
            raise Exception(f"TypeError: input_vector type must be either 'syndrome', 'received_vector' or 'auto'. "
                            f"Not {input_vector_type}")
        super().__init__(parity_check_matrix, error_rate=error_rate, error_channel=error_channel,
                         max_iter=int(max_iter), bp_method=_v1_method(bp_method),
                         ms_scaling_factor=float(ms_scaling_factor), input_vector_type=input_vector_type, **kwargs)

    @property
    def channel_probs(self):
        return self.error_channel

    def update_channel_probs(self, channel):
        self.error_channel = channel


class bposd_decoder(BpOsdDecoder):
    """Legacy ldpc_v1 class: belief propagation plus ordered statistics decoding (v1 argument names)."""

    def __init__(self, parity_check_matrix, error_rate=None, max_iter=0, bp_method="ps", ms_scaling_factor=1.0,
                 channel_probs=[None], osd_method="osd_0", osd_order=0, **kwargs):
        warnings.warn("This is the old syntax for the `bposd_decoder` from `ldpc v1`. Use the `BpOsdDecoder` class "
                      "from `ldpc v2` for additional features.")
        method = _v1_method(bp_method)
        key = str(osd_method).lower()
        if key in ["osd_0", "0", "osd0"]:
            osd_method, osd_order = "osd_0", 0
        elif key in ["osd_e", "1", "osde", "exhaustive", "e"]:
            osd_method = "osd_e"
            if osd_order > 15:
                print("WARNING: Running the 'OSD_E' (Exhaustive method) with search depth greater than 15 is not "
                      "recommended. Use the 'osd_cs' method instead.")
        elif key in ["osd_cs", "2", "osdcs", "combination_sweep", "cs"]:
            osd_method = "osd_cs"
        else:
            raise ValueError(f"ERROR: OSD method '{osd_method}' invalid. Please choose from the following methods: "
                             "'OSD_0', 'OSD_E' or 'OSD_CS'.")
        error_channel = _v1_channel(parity_check_matrix, error_rate, channel_probs)
        super().__init__(parity_check_matrix, error_rate=error_rate, error_channel=error_channel,
                         max_iter=int(max_iter), bp_method=method, ms_scaling_factor=float(ms_scaling_factor),
                         osd_method=osd_method, osd_order=osd_order, **kwargs)

    @property
    def channel_probs(self):
        return self.error_channel

    def update_channel_probs(self, channel):
        self.error_channel = channel
