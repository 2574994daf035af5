# round 2, run K: ncu of the product-sum kernels (config 3), config 5 at 2^19, config 3/4 AUTO, launch list of config 2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_pair -s 1 -c 1 -f \
    -o gpurun_out/prof_ps_pair_r2k python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-stream-family \
    > gpurun_out/prof_ps_pair_r2k.log 2>&1
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err
timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_c5.json 2> gpurun_out/r2k_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bp_|pack_|xor_|compact_|osd|mc_|list_' -c 40 --csv \
    --log-file gpurun_out/r2k_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_launches_c2.log 2>&1
for f in gpurun_out/r2k_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config']['kernel_family'], d['config']['block'], d.get('parity_ok'), d['e2e']['value'], (d.get('e2e_python') or {}).get('value'), d['roofline']['frac'], (d.get('roofline_hbm') or {}).get('frac'))"; done
