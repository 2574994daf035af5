# round 2, run J: full GPU suite with the paired family as AUTO for min-sum; e2e chunk size; config 4/3 with AUTO
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25 > gpurun_out/r2j_pytest.log
for c in 32768 65536 131072; do
BPB_CHUNK_ROWS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2j_bench_c2_chunk$c.json 2> gpurun_out/r2j_bench_c2_chunk$c.err
done
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --kernel pair > gpurun_out/r2j_bench_c3_pair.json 2> gpurun_out/r2j_bench_c3_pair.err
tail -6 gpurun_out/r2j_pytest.log
for f in gpurun_out/r2j_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config']['kernel_family'], d['config']['block'], d.get('parity_ok'), d['e2e']['value'], (d.get('e2e_python') or {}).get('value'))"; done
