# round 2, run B: new feature tests first (fast fail), then the whole GPU suite, bench config 3/4 with device OSD-0, Python e2e profile
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_features.py -q -x 2>&1 | tail -30 > gpurun_out/r2b_features.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -30 > gpurun_out/r2b_pytest.log
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2b_bench_c4.json 2> gpurun_out/r2b_bench_c4.err
timeout 600 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
timeout 300 python - > gpurun_out/r2b_pyprof.log 2>&1 <<'PY'
import cProfile, pstats, time, numpy as np, sys
sys.path.insert(0, '.')
from ldpc_b200 import BpDecoder, codes
H = codes.hamming_code(5)
syn = np.random.default_rng(0).integers(0, 2, size=(1 << 20, 5)).astype(np.uint8)
d = BpDecoder(H, error_rate=0.1, max_iter=2, bp_method="ps", input_vector_type="syndrome")
d.decode_batch(syn); d.decode_batch(syn)
t0 = time.perf_counter(); d.decode_batch(syn); print("hamming decode_batch s", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); d.decode_batch(syn); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(12)
PY
tail -4 gpurun_out/r2b_features.log; tail -4 gpurun_out/r2b_pytest.log
for f in gpurun_out/r2b_bench_c*.json; do echo $f; python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:d.get(k) for k in ("value","ms_per_step","parity_ok","parity_checked")}); print(d["e2e"]); print(d.get("e2e_python")); print(d.get("e2e_bposd")); print(d["roofline"]["frac"], d["roofline"]["bound"])
PY
done
head -30 gpurun_out/r2b_pyprof.log
