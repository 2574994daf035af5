# round 2, run Z: dual-stream host pipeline for the streaming family -- parity with many small chunks, config 5 e2e
mkdir -p gpurun_out
BPB_CHUNK_ROWS=1024 timeout 1200 python -m pytest tests -m gpu -q -x -k "regular_n1000 or config2 or config5 or ragged or pageable" 2>&1 | tail -6 > gpurun_out/r2z_pytest_chunks.log
timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_c5.json 2> gpurun_out/r2z_bench_c5.err
BPB_NO_DUAL_STREAM=1 timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline --no-python-e2e > gpurun_out/r2z_bench_c5_single.json 2> gpurun_out/r2z_bench_c5_single.err
tail -3 gpurun_out/r2z_pytest_chunks.log
for f in gpurun_out/r2z_bench_c5*.json; do python -c "
import json
d=json.load(open('$f')); print('$f', d['value'], d['e2e']['value'], (d.get('e2e_python') or {}).get('value'), d['roofline']['frac'])"; done
