# round 2, run A: full GPU suite incl. the full-size parity tests, bench --config 2/3/5/1, streaming roofline at n=10^4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2a_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
timeout 900 python bench.py --config 5 --steps 2 --warmup 3 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err
timeout 600 python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
for args in "10000 serial 32768" "10000 serial 262144" "10000 parallel 262144" "10000 serial 262144 0.02" "10000 serial 262144 0.08" "1000 serial 1048576"; do
  timeout 600 python scripts/stream_frac.py $args >> gpurun_out/r2a_stream_frac.jsonl 2>> gpurun_out/r2a_stream_frac.err
done
tail -5 gpurun_out/r2a_pytest.log
for f in gpurun_out/r2a_bench_c*.json; do echo $f; cut -c1-600 $f; done
cat gpurun_out/r2a_stream_frac.jsonl
