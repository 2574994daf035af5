mkdir -p gpurun_out
timeout 600 python scripts/host_bw_probe.py 512 > gpurun_out/r2w8_host_bw_8gpu.log 2>&1
grep -v "^$" gpurun_out/r2w8_host_bw_8gpu.log | cut -c1-260 | tail -60
