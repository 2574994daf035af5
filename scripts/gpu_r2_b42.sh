mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "surface or config3 or bit_packed or irregular" 2>&1 | tail -4 > gpurun_out/r2b42_pytest.log; cat gpurun_out/r2b42_pytest.log
timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family --no-python-e2e > gpurun_out/r2b42_bench_c3.json 2> gpurun_out/r2b42_bench_c3.err
python -c "
import json
d=json.load(open('gpurun_out/r2b42_bench_c3.json')); print(d['value'], d['ms_per_step'], d['config']['kernel_family'], d['config']['block'], d['e2e']['value'])"
