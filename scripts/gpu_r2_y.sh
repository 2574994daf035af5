mkdir -p gpurun_out
timeout 600 python scripts/strong_scaling.py --total 2097152 --gpus 1 --b8 > gpurun_out/r2y_strong_b8_1gpu.jsonl 2> gpurun_out/r2y_strong_b8_1gpu.err
cat gpurun_out/r2y_strong_b8_1gpu.jsonl | cut -c1-800; tail -3 gpurun_out/r2y_strong_b8_1gpu.err
