mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -k "not stream" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json")); r = d["roofline"]
print("value %.3e" % d["value"], "ms/step %.2f" % d["ms_per_step"], "kernel_ms %.2f" % r["kernel_ms"], "e2e %.3e" % d["e2e"]["value"])
PY
tail -2 gpurun_out/bench_quick.err
