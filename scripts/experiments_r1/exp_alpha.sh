run() { label=$1; shift; env "$@" timeout 300 python bench.py --ncol 512 --steps 3 --warmup 1 --skip-cpu-baseline --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$label: value %.3e  stencil ms %.3f nloc %.3f launches %d'%(d['value'], r['avg_launch_ms'], r['nloc_ms_per_degree'], d['gpu_launches']))
    elif 'rror' in l: print(l.rstrip())
"; }
for rep in 1 2; do
run reduce_min_2 CHEFSI_B200_ALPHA_REDUCE_MIN=2
run reduce_min_8 CHEFSI_B200_ALPHA_REDUCE_MIN=8
run reduce_min_100 CHEFSI_B200_ALPHA_REDUCE_MIN=100
done
