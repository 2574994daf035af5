# round 2, run L: product-sum after the libm restructure; two-at-a-time variant; parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "surface or ps or product or hamming or golden or bposd or config3" 2>&1 | tail -8 > gpurun_out/r2l_pytest.log
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2l_bench_c3.json 2> gpurun_out/r2l_bench_c3.err
BPB_PS2=1 timeout 400 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2l_bench_c3_ps2.json 2> gpurun_out/r2l_bench_c3_ps2.err
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family --kernel smem > gpurun_out/r2l_bench_c3_smem.json 2> gpurun_out/r2l_bench_c3_smem.err
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_c4.json 2> gpurun_out/r2l_bench_c4.err
tail -4 gpurun_out/r2l_pytest.log
for f in gpurun_out/r2l_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d['config']['kernel_family'], d['config']['block'], d.get('parity_ok'), d['e2e']['value'], (d.get('e2e_python') or {}).get('value'))"; done
