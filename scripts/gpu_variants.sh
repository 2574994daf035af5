# compare experimental builds (LDPC_B200_LIB) on the n=10^4 serial configuration and the streaming bench
mkdir -p gpurun_out
for v in "" _sb2 _mb2 _sb2mb3; do
  lib=ldpc_b200/libbp_b200$v.so
  [ -f $lib ] || continue
  echo "== variant '$v'"
  LDPC_B200_LIB=$PWD/$lib timeout 300 python scripts/prof_serial.py 65536 2>&1 | tail -1
  LDPC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 2 --warmup 2 --kernel stream --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  stream bench: %.3e dec/s frac %.3f' % (d['value'], d['roofline']['frac']))"
done 2>&1 | tee gpurun_out/variants.log
LDPC_B200_LIB=$PWD/ldpc_b200/libbp_b200_sb2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "serial or golden or custom" 2>&1 | tail -2
