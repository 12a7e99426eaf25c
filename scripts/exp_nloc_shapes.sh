# A/B the pipeline shapes of the fused projector kernel (nloc.cu launch_mode): parity tests, then a 512-column bench
for shp in "$@"; do
  echo "== NLOC_SHAPE=$shp"
  CHEFSI_B200_NLOC_SHAPE=$shp timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or full_size or device_resident" 2>&1 | tail -1
  CHEFSI_B200_NLOC_SHAPE=$shp timeout 300 python bench.py --ncol 512 --steps 2 --warmup 1 --skip-cpu-baseline --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.3e  stencil ms %.3f  nloc ms/degree %.3f  clocks %s'%(d['value'], r['avg_launch_ms'], r['nloc_ms_per_degree'], d['clocks']))
    else: print(l.rstrip())
"
done
