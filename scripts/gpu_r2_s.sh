# round 2, run S: serial streaming kernel without message initialisation -- parity, HBM fractions, bench config 5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "stream or config5 or serial or generic or large or ragged or golden" 2>&1 | tail -8 > gpurun_out/r2s_pytest.log
rm -f gpurun_out/r2s_stream_frac.jsonl
run() { timeout 300 python scripts/stream_frac.py "$@" >> gpurun_out/r2s_stream_frac.jsonl 2>> gpurun_out/r2s_stream_frac.err; }
run 10000 serial 262144 0.05
BPB_SERIAL_INIT=1 run 10000 serial 262144 0.05
run 10000 serial 524288 0.05
run 10000 serial 75776 0.05
run 10000 serial 262144 0.08
run 1000 serial 1048576 0.05
timeout 600 python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench_c5.json 2> gpurun_out/r2s_bench_c5.err
tail -4 gpurun_out/r2s_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2s_stream_frac.jsonl'):
    d=json.loads(l); print(d['n'],d['schedule'],d['batch'],d['p'],'step_ms %.1f kern_ms %.1f mean_it %.2f handed %d frac %.3f grid %d'%(d['step_ms'],d['kernel_ms'],d['mean_it'],d['handed_off'],d['frac_of_measured_hbm'],d['grid']))
d=json.load(open('gpurun_out/r2s_bench_c5.json')); print(d['value'], d['ms_per_step'], d.get('parity_ok'), d['e2e']['value'], d['roofline']['frac'])
PY
