# quick loop: smem parity subset + bench (auto) + one ncu capture of the smem kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "smem or golden" 2>&1 | tail -5 > gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_smem -s 1 -c 1 -f \
    -o gpurun_out/prof_smem_quick python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 262144 > gpurun_out/prof_smem_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log; cat gpurun_out/bench_quick.json | cut -c1-900; tail -2 gpurun_out/bench_quick.err
