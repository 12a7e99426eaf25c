run() { label=$1; shift; env "$@" timeout 600 python bench.py --ncol 64 --block 64 --steps 2 --warmup 1 --skip-cpu-baseline --e2e-cols 8 --kpt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$label: value %.3e  %s avg ms %.3f frac %.3f nloc ms/degree %.3f'%(d['value'], r['kernel'][:20], r['avg_launch_ms'], r['frac'], r['nloc_ms_per_degree']))
    elif 'rror' in l or 'Trace' in l: print(l.rstrip())
"; }
for rep in 1 2; do
run barrier_on CHEFSI_B200_GRIDSYNC=2
run barrier_off CHEFSI_B200_GRIDSYNC=1
done
