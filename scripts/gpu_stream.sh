mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stream" 2>&1 | tail -5 > gpurun_out/pytest_stream.log
timeout 600 python bench.py --steps 3 --warmup 3 --kernel stream --no-cpu-baseline > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
timeout 1500 python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
tail -3 gpurun_out/pytest_stream.log; cat gpurun_out/bench_stream.json | cut -c1-1700; tail -2 gpurun_out/bench_stream.err; cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
