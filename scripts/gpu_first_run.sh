mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
