"""Known-answer vectors restated from the reference's own tests (paths relative to /root/reference).

Each entry: (name, code, decoder kwargs, inputs, expected decodings, input kind).
"""
import numpy as np

from ldpc_b200 import codes

_REP3_SYN = [[0, 0], [0, 1], [1, 0], [1, 1]]
_REP3_DEC = [[0, 0, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0]]
_REP5_SYN = [[0, 0, 0, 0], [0, 0, 0, 1], [0, 1, 0, 1], [1, 0, 1, 0], [1, 1, 1, 1]]
_REP5_DEC = [[0, 0, 0, 0, 0], [0, 0, 0, 0, 1], [0, 0, 1, 1, 0], [0, 1, 1, 0, 0], [0, 1, 0, 1, 0]]

KATS = [
    # cpp_test/TestBPDecoder.cpp:122-164 (product_sum_parallel, rep code n=3, p=0.1, max_iter=n)
    ("cpp_ps_parallel_rep3", codes.rep_code(3), dict(channel=0.1, max_iter=3, bp_method="ps", schedule="parallel"),
     _REP3_SYN, _REP3_DEC, "syndrome"),
    # cpp_test/TestBPDecoder.cpp:301-344 (min_sum_parallel, ms_scaling 0.625)
    ("cpp_ms_parallel_rep3", codes.rep_code(3),
     dict(channel=0.1, max_iter=3, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625),
     _REP3_SYN, _REP3_DEC, "syndrome"),
    # cpp_test/TestBPDecoder.cpp:166-197
    ("cpp_ps_parallel_rep5", codes.rep_code(5), dict(channel=0.1, max_iter=5, bp_method="ps", schedule="parallel"),
     _REP5_SYN, _REP5_DEC, "syndrome"),
    # cpp_test/TestBPDecoder.cpp:200-231 (ms_scaling 1)
    ("cpp_ms_parallel_rep5", codes.rep_code(5),
     dict(channel=0.1, max_iter=5, bp_method="ms", schedule="parallel", ms_scaling_factor=1.0),
     _REP5_SYN, _REP5_DEC, "syndrome"),
    # cpp_test/TestBPDecoder.cpp:234-265
    ("cpp_ps_serial_rep5", codes.rep_code(5), dict(channel=0.1, max_iter=5, bp_method="ps", schedule="serial"),
     _REP5_SYN, _REP5_DEC, "syndrome"),
    # cpp_test/TestBPDecoder.cpp:268-299
    ("cpp_ms_serial_rep5", codes.rep_code(5),
     dict(channel=0.1, max_iter=5, bp_method="ms", schedule="serial", ms_scaling_factor=1.0),
     _REP5_SYN, _REP5_DEC, "syndrome"),
    # cpp_test/TestBPDecoder.cpp:511-534 (received-vector input, product-sum serial)
    ("cpp_received_vector_rep5", codes.rep_code(5), dict(channel=0.1, max_iter=5, bp_method="ps", schedule="serial"),
     [[0, 0, 0, 0, 1], [0, 1, 1, 0, 0], [1, 0, 0, 1, 1]],
     [[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1]], "received_vector"),
    # python_test/test_bp_decoder.py:175-186 (rep_code(3), error_rate 0.1, syndrome [1,1] -> [0,1,0])
    ("py_ps_rep3", codes.rep_code(3), dict(channel=0.1, max_iter=3, bp_method="ps", schedule="parallel"),
     [[1, 1]], [[0, 1, 0]], "syndrome"),
    # python_test/test_bp_decoder.py:188-192 (error_channel [0.1, 0, 0.1]: p = 0 gives an infinite prior)
    ("py_ps_rep3_certain_bit", codes.rep_code(3),
     dict(channel=[0.1, 0.0, 0.1], max_iter=3, bp_method="ps", schedule="parallel"), [[1, 1]], [[1, 0, 1]], "syndrome"),
    ("py_ms_rep3_certain_bit", codes.rep_code(3),
     dict(channel=[0.1, 0.0, 0.1], max_iter=3, bp_method="ms", schedule="parallel", ms_scaling_factor=1.0),
     [[1, 1]], [[1, 0, 1]], "syndrome"),
    # python_test/test_bp_decoder.py:214-235 (serial schedule with a custom order)
    ("py_ps_serial_custom_order", codes.rep_code(3),
     dict(channel=0.1, max_iter=3, bp_method="ps", schedule="serial", serial_schedule_order=[1, 2, 0]),
     [[1, 1]], [[0, 1, 0]], "syndrome"),
    # docs/source/bp_decoding_example.ipynb cell 7: rep_code(3) received vector [1,0,1] -> [1,1,1]
    ("nb_received_vector_rep3", codes.rep_code(3), dict(channel=0.1, max_iter=3, bp_method="ps", schedule="parallel"),
     [[1, 0, 1]], [[1, 1, 1]], "received_vector"),
]


def kat_arrays(entry):
    name, H, kw, inputs, expected, kind = entry
    return name, H, dict(kw), np.asarray(inputs, dtype=np.uint8), np.asarray(expected, dtype=np.uint8), kind
