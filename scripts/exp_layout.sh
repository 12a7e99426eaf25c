# A/B padded vs dense column layout: 512-column bench, stencil-only and with projectors
for dense in 0 1 0 1; do
  for extra in "--no-nloc" ""; do
  echo "== DENSE=$dense $extra"
  CHEFSI_B200_DENSE=$dense timeout 300 python bench.py --ncol 512 --steps 2 --warmup 1 --skip-cpu-baseline --e2e-cols 64 $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.3e  stencil ms %.3f frac %.3f nloc ms/degree %.3f  e2e %.3e clocks %s'%(d['value'], r['avg_launch_ms'], r['frac'], r['nloc_ms_per_degree'], d['e2e']['value'], d['clocks']))
    else: print(l.rstrip())
"
  done
done
