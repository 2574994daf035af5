# round 2, run I: paired on-chip family v2 -- parity, CTA-size sweep on config 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pair" 2>&1 | tail -15 > gpurun_out/r2i_pytest_pair.log
for c in 512; do
  BPB_PAIR_CTA_THREADS=$c timeout 300 python bench.py --kernel pair --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2i_bench_pair_$c.json 2> gpurun_out/r2i_bench_pair_$c.err
done
tail -5 gpurun_out/r2i_pytest_pair.log
for f in gpurun_out/r2i_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config']['grid'], d['config']['block'], d.get('parity_ok'), d['e2e']['value'])"; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_pair -s 1 -c 1 -f \
    -o gpurun_out/prof_pair_r2i python bench.py --kernel pair --steps 1 --warmup 1 --no-cpu-baseline --no-stream-family \
    > gpurun_out/prof_pair_r2i.log 2>&1
