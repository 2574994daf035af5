# round 2, run T: bit-packed I/O -- tests (b8, sinter), bench config 4 and 2 with the packed leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "packed or sinter or bposd or device_osd" 2>&1 | tail -8 > gpurun_out/r2t_pytest.log
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_c4.json 2> gpurun_out/r2t_bench_c4.err
timeout 400 python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2t_bench_c2.json 2> gpurun_out/r2t_bench_c2.err
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family > gpurun_out/r2t_bench_c3.json 2> gpurun_out/r2t_bench_c3.err
tail -4 gpurun_out/r2t_pytest.log; tail -3 gpurun_out/r2t_bench_c4.err
for f in gpurun_out/r2t_bench_c*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d['value'], d['e2e']['value'], d.get('e2e_packed'), (d.get('e2e_python') or {}).get('value'))"; done
