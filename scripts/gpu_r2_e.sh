# round 2, run E: ncu full capture of the paired on-chip kernel and of the one-syndrome on-chip kernel (config 2)
mkdir -p gpurun_out
for k in pair smem; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_$k -s 1 -c 1 -f \
    -o gpurun_out/prof_${k}_r2e python bench.py --kernel $k --steps 1 --warmup 1 --no-cpu-baseline --no-stream-family \
    > gpurun_out/prof_${k}_r2e.log 2>&1
done
ls -la gpurun_out | grep r2e
