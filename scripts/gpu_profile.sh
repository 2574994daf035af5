# Evidence run (B200_PROFILING.md recipe): launch list of the bench command (our kernels only) and one full ncu
# capture per message-update kernel at the bench batch.  Usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bp_|pack_|xor_|compact_' -c 16 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > gpurun_out/launches_bench_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bp_smem -s 1 -c 1 -f \
    -o gpurun_out/prof_smem_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-stream-family \
    > gpurun_out/prof_smem_${TAG}.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f \
    -o gpurun_out/prof_stream_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --kernel stream \
    > gpurun_out/prof_stream_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f \
    -o gpurun_out/prof_serial_${TAG} python scripts/prof_serial.py 32768 > gpurun_out/prof_serial_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
