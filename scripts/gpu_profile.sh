# ncu launch list (our kernels only) + one full capture of the message-update kernel per family
# (B200_PROFILING.md recipe).  Usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bp_|pack_|xor_' -c 12 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/launches_bench_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bp_smem -s 1 -c 1 -f \
    -o gpurun_out/prof_smem_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 262144 \
    > gpurun_out/prof_smem_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f \
    -o gpurun_out/prof_stream_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 262144 --kernel stream \
    > gpurun_out/prof_stream_${TAG}.log 2>&1
ls -la gpurun_out | tail -8
