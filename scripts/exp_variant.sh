# A/B of the dense streaming kernel's thread mappings, stencil-only (512 columns)
for rep in 1 2; do
for cfg in ${VARIANTS:-0 1 2}; do
  echo "== STREAM_VARIANT=$cfg"
  env CHEFSI_B200_DENSE=1 CHEFSI_B200_STREAM_VARIANT=$cfg timeout 300 python bench.py --ncol 512 --steps 3 --warmup 1 --skip-cpu-baseline --e2e-cols 16 --no-nloc 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.3e  stencil ms %.3f frac %.3f  clocks %s'%(d['value'], r['avg_launch_ms'], r['frac'], d['clocks']))
    else: print(l.rstrip())
"
done
done
