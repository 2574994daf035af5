# round 2, run O: SERIAL_RELATIVE kernel -- parity tests, sanitizer on a small case, throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "relative or golden" 2>&1 | tail -15 > gpurun_out/r2o_pytest.log
timeout 600 compute-sanitizer --tool memcheck python - > gpurun_out/r2o_sanitizer.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from ldpc_b200 import BpDecoder, codes
H = codes.regular_ldpc(240, 3, 6, seed=3)
syn = codes.bsc_syndromes(H, 0.07, 64, seed=1)
for meth in ("ms", "ps"):
    d = BpDecoder(H, error_rate=0.07, input_vector_type="syndrome", max_iter=10, bp_method=meth, schedule="serial_relative", ms_scaling_factor=0.625)
    d.decode_batch(syn)
    d.decode(syn[0])
print("sanitizer run done")
PY
timeout 300 python - > gpurun_out/r2o_speed.log 2>&1 <<'PY'
import numpy as np, sys, time
sys.path.insert(0, '.')
from ldpc_b200 import BpDecoder, codes
for n, B in ((1000, 1 << 15), (10000, 2048)):
    H = codes.regular_ldpc(n, 3, 6, seed=1)
    syn = codes.bsc_syndromes(H, 0.05, B, seed=7)
    d = BpDecoder(H, error_rate=0.05, input_vector_type="syndrome", max_iter=50, bp_method="ms", schedule="serial_relative", ms_scaling_factor=0.625)
    d.decode_batch(syn)
    t0 = time.perf_counter(); d.decode_batch(syn); dt = time.perf_counter() - t0
    print(n, B, "decodes/s", B / dt, "mean it", d.iter_batch.mean(), "conv", d.converge_batch.mean(), "kernel ms", d.info()["last_kernel_ms"], d.info()["grid"])
PY
tail -5 gpurun_out/r2o_pytest.log; tail -5 gpurun_out/r2o_sanitizer.log; cat gpurun_out/r2o_speed.log
