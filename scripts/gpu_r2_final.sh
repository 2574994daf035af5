# round 2, evidence run: GPU suite, bench of every BASELINE config, launch list, ncu full captures of the dominant kernels,
# latency table.  Outputs are copied into profiles/ by hand afterwards.
mkdir -p gpurun_out
T=r2g
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv > gpurun_out/${T}_gpu.txt 2>&1; nproc >> gpurun_out/${T}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -25 > gpurun_out/${T}_pytest.log
for c in 2 1 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_c2_reference.json 2> gpurun_out/${T}_bench_c2_reference.err
# launch list of the default bench command (our kernels only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bp_|pack_|xor_|compact_|osd|mc_|list_' -c 60 --csv \
    --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-python-e2e > gpurun_out/${T}_launches_c2.log 2>&1
# ncu --set full, one launch each
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f -o gpurun_out/${T}_prof_stream_c2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel stream --no-python-e2e > gpurun_out/${T}_prof_stream_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f -o gpurun_out/${T}_prof_serial_n10000 \
    python scripts/stream_frac.py 10000 serial 524288 > gpurun_out/${T}_prof_serial_n10000.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f -o gpurun_out/${T}_prof_parallel_n10000 \
    python scripts/stream_frac.py 10000 parallel 262144 > gpurun_out/${T}_prof_parallel_n10000.log 2>&1
timeout 600 python scripts/latency_table.py 2 > gpurun_out/${T}_latency_c2.jsonl 2> gpurun_out/${T}_latency_c2.err
tail -4 gpurun_out/${T}_pytest.log
for f in gpurun_out/${T}_bench_c*.json; do echo $f; python -c "
import json
d=json.load(open('$f')); print(d.get('value'), d.get('ms_per_step'), (d.get('config') or {}).get('kernel_family'), d.get('parity_ok'), (d.get('e2e') or {}).get('value'), (d.get('e2e_python') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_hbm') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))"; done
ls -la gpurun_out | grep ${T}_prof
