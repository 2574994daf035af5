# round 2, run M: smoke test of the strong-scaling script on one GPU
mkdir -p gpurun_out
timeout 600 python scripts/strong_scaling.py --total 2097152 --gpus 1,2 > gpurun_out/r2m_strong_1gpu.jsonl 2> gpurun_out/r2m_strong_1gpu.err
cat gpurun_out/r2m_strong_1gpu.jsonl; tail -5 gpurun_out/r2m_strong_1gpu.err
