# pytest -m gpu + smoke + bench (both kernel families)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
timeout 600 python bench.py --steps 3 --warmup 3 --kernel stream --no-cpu-baseline > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
tail -8 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_auto.json; tail -3 gpurun_out/bench_auto.err; cat gpurun_out/bench_stream.json; tail -3 gpurun_out/bench_stream.err
