# round 2, run R: ramp-down lane compaction in the streaming family -- parity, then HBM fractions with / without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "stream or config5 or serial or generic or large or ragged" 2>&1 | tail -8 > gpurun_out/r2r_pytest.log
rm -f gpurun_out/r2r_stream_frac.jsonl
run() { timeout 300 python scripts/stream_frac.py "$@" >> gpurun_out/r2r_stream_frac.jsonl 2>> gpurun_out/r2r_stream_frac.err; }
run 10000 serial 262144 0.05
BPB_NO_COMPACTION=1 run 10000 serial 262144 0.05
run 10000 serial 524288 0.05
run 10000 serial 75776 0.05
run 10000 parallel 262144 0.05
BPB_NO_COMPACTION=1 run 10000 parallel 262144 0.05
run 1000 serial 1048576 0.05
run 1000 parallel 1048576 0.05
BPB_NO_COMPACTION=1 run 1000 parallel 1048576 0.05
tail -4 gpurun_out/r2r_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r2r_stream_frac.jsonl'):
    d=json.loads(l); print(d['n'],d['schedule'],d['batch'],d['p'],'step_ms %.1f kern_ms %.1f mean_it %.2f handed %d frac %.3f'%(d['step_ms'],d['kernel_ms'],d['mean_it'],d['handed_off'],d['frac_of_measured_hbm']))
PY
