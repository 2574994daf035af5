"""Drive one decode of the n=10^4 serial min-sum configuration (for ncu).  Usage: python scripts/prof_serial.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldpc_b200 import BpDecoder, codes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
H = codes.regular_ldpc(10000, 3, 6, seed=1)
syn = codes.bsc_syndromes(H, 0.05, B, seed=7)
d = BpDecoder(H, error_rate=0.05, max_iter=100, bp_method="ms", schedule="serial", ms_scaling_factor=0.625,
              input_vector_type="syndrome")
for _ in range(2):
    d.decode_batch(syn)
print("mean iterations", d.iter_batch.mean(), "kernel ms", d.info()["last_kernel_ms"])
