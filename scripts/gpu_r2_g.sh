# round 2, run G: streaming family at n = 10^4 -- where does the HBM fraction go (steady state vs tail)?
mkdir -p gpurun_out
rm -f gpurun_out/r2g_stream_frac.jsonl
run() { timeout 300 python scripts/stream_frac.py "$@" >> gpurun_out/r2g_stream_frac.jsonl 2>> gpurun_out/r2g_stream_frac.err; }
run 10000 serial 262144 0.05
run 10000 serial 262144 0.05 4
run 10000 serial 524288 0.05
run 10000 serial 75776 0.05
run 10000 serial 262144 0.08
run 10000 parallel 262144 0.05
BPB_NO_SECOND_STAGE=1 run 10000 parallel 262144 0.05
run 10000 parallel 262144 0.05 6
run 10000 parallel 262144 0.08
run 1000 serial 1048576 0.05
run 1000 parallel 1048576 0.05
cat gpurun_out/r2g_stream_frac.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f \
    -o gpurun_out/prof_serial_r2g python scripts/prof_serial.py 75776 > gpurun_out/prof_serial_r2g.log 2>&1
