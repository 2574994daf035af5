# ncu launch list + one full capture of the message-update kernel (B200_PROFILING.md recipe).
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/cpuinfo.txt 2>&1
import os
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
os.system("grep -m1 'model name' /proc/cpuinfo; grep -c processor /proc/cpuinfo")
import sys; sys.path.insert(0, ".")
import numpy as np, time, oracle
from ldpc_b200 import codes
H = codes.regular_ldpc(1000, 3, 6, seed=1); R = oracle.RefOracle()
syn = codes.bsc_syndromes(H, 0.05, 1 << 16, seed=7)
kw = dict(max_iter=50, bp_method="ms", schedule="parallel", ms_scaling_factor=0.625, want_llr=False)
for t in (1, 8, 16, 32, 64, 128):
    n = min(1 << 16, 4096 * t)
    out = R.decode_batch(H, syn[:n], 0.05, threads=t, return_seconds=True, **kw)
    print("threads", t, "n", n, "decodes/s (decode loop only)", n / out[-1])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:bp_stream -s 1 -c 1 -f -o gpurun_out/prof_stream_r1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 262144 > gpurun_out/prof_bench.log 2>&1
cat gpurun_out/cpuinfo.txt; tail -3 gpurun_out/launches_bench.log; ls -la gpurun_out
