mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "soft_info or relative" 2>&1 | tail -12 > gpurun_out/r2soft_pytest.log
tail -12 gpurun_out/r2soft_pytest.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python - > gpurun_out/r2soft_san.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from ldpc_b200 import SoftInfoBpDecoder, codes
H = codes.regular_ldpc(240, 3, 6, seed=3)
rng = np.random.default_rng(1)
syn = codes.bsc_syndromes(H, 0.05, 64, seed=1)
soft = (1 - 2.0 * syn) + rng.normal(0, 0.7, size=syn.shape)
d = SoftInfoBpDecoder(H, error_rate=0.05, max_iter=15, ms_scaling_factor=0.625, cutoff=3.0, sigma=0.7)
d.decode_batch(soft, return_llr=True)
print("done")
PY
echo "sanitizer rc=$?"; grep "ERROR SUMMARY" gpurun_out/r2soft_san.log
