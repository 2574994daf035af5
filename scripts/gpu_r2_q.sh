# round 2, run Q: dual-stream host pipeline -- whole GPU suite, e2e with and without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2q_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench_c2.err
BPB_NO_DUAL_STREAM=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2q_bench_c2_single.json 2> gpurun_out/r2q_bench_c2_single.err
timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2q_bench_c3.json 2> gpurun_out/r2q_bench_c3.err
tail -3 gpurun_out/r2q_pytest.log
for f in gpurun_out/r2q_bench_*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config']['kernel_family'], d.get('parity_ok'), d['e2e'], (d.get('e2e_python') or {}).get('value'))"; done
