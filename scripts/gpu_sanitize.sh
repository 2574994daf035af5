# compute-sanitizer over small decodes of every kernel family: memcheck everywhere, racecheck on the shared-memory kernels
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
import numpy as np
sys.path.insert(0, '.')
from ldpc_b200 import BpDecoder, BpOsdDecoder, codes
which = sys.argv[1]
H = codes.regular_ldpc(240, 3, 6, seed=3)
syn = codes.bsc_syndromes(H, 0.07, 96, seed=1)
kw = dict(max_iter=12, ms_scaling_factor=0.625, input_vector_type="syndrome")
if which in ("pair", "smem", "edge", "stream"):
    for meth in ("ms", "ps"):
        d = BpDecoder(H, error_rate=0.07, bp_method=meth, schedule="parallel", kernel=which, **kw)
        d.decode_batch(syn, return_llr=True)
        d.decode(syn[0])
elif which == "stream_serial":
    for meth in ("ms", "ps"):
        d = BpDecoder(H, error_rate=0.07, bp_method=meth, schedule="serial", kernel="stream", **kw)
        d.decode_batch(syn, return_llr=True)
elif which == "smem_serial":
    d = BpDecoder(H, error_rate=0.07, bp_method="ms", schedule="serial", kernel="smem", **kw)
    d.decode_batch(syn)
elif which == "relative":
    d = BpDecoder(H, error_rate=0.07, bp_method="ms", schedule="serial_relative", **kw)
    d.decode_batch(syn[:32])
elif which == "b8_osd":
    Hb = codes.bivariate_bicycle_144()
    s2 = codes.bsc_syndromes(Hb, 0.03, 512, seed=2)
    d = BpOsdDecoder(Hb, error_rate=0.03, bp_method="ms", ms_scaling_factor=0.625, max_iter=10, osd_method="osd0")
    d.set_observables(np.eye(12, 144, dtype=np.uint8))
    d.decode_batch_b8(np.packbits(s2, axis=1, bitorder="little"), decoding=True, observables=True)
    d.decode_batch(s2)
elif which == "mc":
    d = BpDecoder(H, error_rate=0.05, bp_method="ms", **kw)
    d.monte_carlo_bsc(2048, seed=3)
print("case", which, "done")
PY
for c in pair smem edge stream stream_serial smem_serial relative b8_osd mc; do
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_case.py $c > gpurun_out/san_mem_$c.log 2>&1; echo "memcheck $c rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/san_mem_$c.log)"
done
for c in pair smem edge smem_serial relative; do
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san_case.py $c > gpurun_out/san_race_$c.log 2>&1; echo "racecheck $c rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/san_race_$c.log | head -1)"
done
