# A/B the streaming-kernel configurations: correctness (streaming tests) then stencil-only bench
for cfg in "$@"; do
  echo "== STREAM_CFG=$cfg"
  CHEFSI_B200_STREAM_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or full_size or device_resident" 2>&1 | tail -2
  CHEFSI_B200_STREAM_CFG=$cfg timeout 300 python bench.py --ncol 512 --steps 2 --warmup 1 --skip-cpu-baseline --no-nloc --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.3e  stencil avg ms %.3f  achieved %.0f GB/s frac %.3f clocks %s'%(d['value'], r['avg_launch_ms'], r['achieved'], r['frac'], d['clocks']))
    else: print(l.rstrip())
"
done
