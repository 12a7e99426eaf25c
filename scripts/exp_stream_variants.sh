for gs in 0 1; do for pr in 0 2 3; do
  echo "== gridsync=$gs promo=$pr"
  CHEFSI_B200_GRIDSYNC=$gs CHEFSI_B200_TMA_L2PROMO=$pr timeout 300 python bench.py --ncol 512 --steps 2 --warmup 1 --skip-cpu-baseline --no-nloc --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.3e  stencil avg ms %.3f  achieved %.0f GB/s frac %.3f clocks %s'%(d['value'], r['avg_launch_ms'], r['achieved'], r['frac'], d['clocks']))
    else: print(l.rstrip())
"
done; done
