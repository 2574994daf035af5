# round 2, 8-GPU run: strong scaling of one host batch in one process (bpb_set_devices), then the weak-scaling bench
# (one process per GPU, torchrun) for configs 2, 3, 4 at N = 8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n8_gpus.txt; nproc >> gpurun_out/r2n8_gpus.txt; free -g >> gpurun_out/r2n8_gpus.txt
timeout 900 python scripts/strong_scaling.py --config 2 --total 8388608 --gpus 1,2,4,8 > gpurun_out/r2n8_strong_c2.jsonl 2> gpurun_out/r2n8_strong_c2.err
timeout 600 python scripts/strong_scaling.py --config 4 --total 8000000 --gpus 1,8 --reps 2 > gpurun_out/r2n8_strong_c4.jsonl 2> gpurun_out/r2n8_strong_c4.err
for c in 2 3 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-stream-family --no-python-e2e > gpurun_out/r2n8_bench_c${c}_n8.json 2> gpurun_out/r2n8_bench_c${c}_n8.err
done
cat gpurun_out/r2n8_strong_c2.jsonl gpurun_out/r2n8_strong_c4.jsonl | cut -c1-60,150-700
for f in gpurun_out/r2n8_bench_c*_n8.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['n_gpus'], d['ms_per_step'], d['config']['kernel_family'], d['e2e']['value'])"; done
