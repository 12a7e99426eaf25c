# knobs re-checked on the leaner K1: TMA L2 promotion, round barrier
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 150 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s40_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run default X=1
run promo2 CHEFSI_B200_TMA_L2PROMO=2
run promo1 CHEFSI_B200_TMA_L2PROMO=1
run promo0 CHEFSI_B200_TMA_L2PROMO=0
run nosync CHEFSI_B200_GRIDSYNC=0
run default2 X=1
