# N-GPU check of the bench contract (torchrun, NCCL only for the timing barrier).  Usage: bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("N=%d value %.4e ms/step %.2f e2e %.4e stream_family %.4e frac %.3f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["stream_family"]["value"], d["stream_family"]["roofline"]["frac"]))
r = json.load(open("gpurun_out/bench_ref_n$N.json")); print("reference arm %.4e" % r["value"], r["cpu_baseline"]["cores"])
PY
tail -2 gpurun_out/bench_n$N.err
