# round 2, run P: paired kernel with claim-ahead -- parity subset + bench config 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pair or ragged or second or bposd" 2>&1 | tail -6 > gpurun_out/r2p_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stream-family --no-python-e2e > gpurun_out/r2p_bench_c2.json 2> gpurun_out/r2p_bench_c2.err
tail -3 gpurun_out/r2p_pytest.log
python -c "
import json
d=json.load(open('gpurun_out/r2p_bench_c2.json')); print(d['value'], d['ms_per_step'], d['config']['kernel_family'], d.get('parity_ok'), d['e2e']['value'])"
