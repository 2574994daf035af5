# round 2, run C (re-entry): whole GPU suite, then bench config 2 (default), 5, 3, 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c_gpu.txt 2>&1
nproc >> gpurun_out/r2c_gpu.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2c_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 > gpurun_out/r2c_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err
timeout 400 python bench.py --config 5 --steps 2 --warmup 3 > gpurun_out/r2c_bench_c5.json 2> gpurun_out/r2c_bench_c5.err
timeout 400 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err
timeout 400 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
tail -8 gpurun_out/r2c_pytest.log
for f in gpurun_out/r2c_bench_c*.json; do echo $f; cut -c1-1500 $f; done
