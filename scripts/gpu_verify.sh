# final check: smoke(), the whole GPU suite, the default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/verify_smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/verify_pytest.log; cat gpurun_out/verify_pytest.log
timeout 400 python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; python -c "
import json
d=json.load(open('gpurun_out/verify_bench.json')); print(d['value'], d['e2e']['value'], d['parity_ok'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/verify_bench_c5.json 2> gpurun_out/verify_bench_c5.err; python -c "
import json
d=json.load(open('gpurun_out/verify_bench_c5.json')); print('c5', d['value'], d['e2e']['value'], (d.get('e2e_python') or {}).get('value'), d['roofline']['frac'])"
