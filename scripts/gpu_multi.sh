# 2-GPU check of the bench contract (torchrun, NCCL only for the timing barrier)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
timeout 600 python - <<'PY' > gpurun_out/multi_api.log 2>&1
import sys, time, numpy as np
sys.path.insert(0, ".")
from ldpc_b200 import codes
from ldpc_b200.parallel import MultiGpuBpDecoder
from ldpc_b200 import BpDecoder
H = codes.regular_ldpc(1000, 3, 6, seed=1)
syn = codes.bsc_syndromes(H, 0.05, 1 << 18, seed=7)
kw = dict(error_rate=0.05, max_iter=50, bp_method="ms", ms_scaling_factor=0.625, input_vector_type="syndrome")
one = BpDecoder(H, device=0, **kw); ref = one.decode_batch(syn)
multi = MultiGpuBpDecoder(H, devices=[0, 1], **kw)
out = multi.decode_batch(syn)
t = time.perf_counter(); out = multi.decode_batch(syn); dt = time.perf_counter() - t
print("multi == single:", np.array_equal(out, ref), np.array_equal(multi.iter_batch, one.iter_batch), "decodes/s (pageable host arrays)", syn.shape[0] / dt)
PY
cat gpurun_out/bench_n2.json | cut -c1-400; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_ref_n2.json | cut -c1-300; tail -2 gpurun_out/bench_ref_n2.err; cat gpurun_out/multi_api.log | tail -3
