run() { label=$1; shift; env "$@" timeout 300 python bench.py --ncol 512 --steps 3 --warmup 1 --skip-cpu-baseline --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$label: value %.3e  stencil ms %.3f frac %.3f nloc %.3f clocks %s'%(d['value'], r['avg_launch_ms'], r['frac'], r['nloc_ms_per_degree'], d['clocks']))
    elif 'rror' in l: print(l.rstrip())
"; }
for rep in 1 2; do
run gridsync1 CHEFSI_B200_GRIDSYNC=1
run gridsync0 CHEFSI_B200_GRIDSYNC=0
done
