mkdir -p gpurun_out
timeout 200 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err
python -c "
import json
d=json.load(open('gpurun_out/r2h_bench_c3.json')); print(d['value'], d['e2e']['value'], d.get('parity_ok'), (d.get('e2e_bposd') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))"
