# round 2, run D: paired on-chip family -- parity, then bench config 2 pair vs smem
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pair" 2>&1 | tail -15 > gpurun_out/r2d_pytest_pair.log
for k in pair smem; do
  timeout 300 python bench.py --kernel $k --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2d_bench_$k.json 2> gpurun_out/r2d_bench_$k.err
done
for t in 64 96 128 160 256; do
  BPB_PAIR_GROUP_THREADS=$t timeout 300 python bench.py --kernel pair --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2d_bench_pair_T$t.json 2> gpurun_out/r2d_bench_pair_T$t.err
done
tail -5 gpurun_out/r2d_pytest_pair.log
for f in gpurun_out/r2d_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config']['grid'], d['config']['block'], d.get('parity_ok'), d['e2e']['value'])"; done
