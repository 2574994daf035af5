mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "serial or golden or custom or irregular or stream" 2>&1 | tail -3
timeout 300 python scripts/prof_serial.py 65536 2>&1 | tail -1
timeout 300 python bench.py --steps 2 --warmup 2 --kernel stream --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('stream bench: %.3e dec/s frac %.3f' % (d['value'], d['roofline']['frac']))"
