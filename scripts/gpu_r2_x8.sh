# round 2, 8-GPU run 2: strong scaling with bit-packed I/O (config 2 and 4), one process, one handle
mkdir -p gpurun_out
timeout 900 python scripts/strong_scaling.py --config 2 --total 8388608 --gpus 1,2,4,8 --b8 > gpurun_out/r2x8_strong_c2_b8.jsonl 2> gpurun_out/r2x8_strong_c2_b8.err
timeout 600 python scripts/strong_scaling.py --config 4 --total 8000000 --gpus 1,8 --reps 2 --b8 > gpurun_out/r2x8_strong_c4_b8.jsonl 2> gpurun_out/r2x8_strong_c4_b8.err
cat gpurun_out/r2x8_strong_c2_b8.jsonl gpurun_out/r2x8_strong_c4_b8.jsonl | cut -c1-60,150-700; tail -3 gpurun_out/r2x8_strong_c2_b8.err
