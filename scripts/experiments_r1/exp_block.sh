for rep in 1 2; do
for blk in 128 148 296; do
  timeout 600 python bench.py --ncol 1184 --block $blk --steps 2 --warmup 1 --skip-cpu-baseline --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('block $blk: value %.3e  stencil ms %.3f frac %.3f nloc %.3f clocks %s'%(d['value'], r['avg_launch_ms'], r['frac'], r['nloc_ms_per_degree'], d['clocks']))
    elif 'rror' in l: print(l.rstrip())
"
done
done
