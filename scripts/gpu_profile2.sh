mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f \
    -o gpurun_out/prof_serial_n10000 python scripts/prof_serial.py 32768 > gpurun_out/prof_serial.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
tail -2 gpurun_out/prof_serial.log; cut -c1-300 gpurun_out/bench_auto.json; tail -2 gpurun_out/bench_auto.err
