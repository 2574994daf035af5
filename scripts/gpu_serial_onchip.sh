mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -4
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, ".")
from ldpc_b200 import BpDecoder, codes
H = codes.regular_ldpc(1000, 3, 6, seed=1)
syn = codes.bsc_syndromes(H, 0.05, 1 << 19, seed=7)
for kern in ("smem", "stream"):
    for meth in ("ms", "ps"):
        d = BpDecoder(H, error_rate=0.05, max_iter=50, bp_method=meth, schedule="serial", ms_scaling_factor=0.625,
                      input_vector_type="syndrome", kernel=kern)
        d.decode_batch(syn[:4096]); t = time.perf_counter(); d.decode_batch(syn); dt = time.perf_counter() - t
        print("n=1000 serial", meth, kern, "e2e(pageable) %.3e dec/s" % (syn.shape[0] / dt), "mean it %.2f" % d.iter_batch.mean(),
              "kernel ms (last chunk)", round(d.info()["last_kernel_ms"], 2), "grid/block", d.info()["grid"], d.info()["block"])
PY
