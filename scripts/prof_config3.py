"""Drive one decode of BASELINE config 3 (d=13 rotated surface code X checks, product-sum, 30 iterations) for ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldpc_b200 import BpDecoder, codes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
H = codes.rotated_surface_code_x(13)
syn = codes.bsc_syndromes(H, 0.05, B, seed=7)
d = BpDecoder(H, error_rate=0.05, max_iter=30, bp_method="ps", schedule="parallel", input_vector_type="syndrome")
for _ in range(2):
    d.decode_batch(syn)
print("mean iterations", d.iter_batch.mean(), "kernel ms", d.info()["last_kernel_ms"], d.info()["grid"], d.info()["block"])
