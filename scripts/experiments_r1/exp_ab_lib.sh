# same-box A/B of two library builds: scripts/ab/libchefsi_old.so (HEAD) vs the working tree
run() { label=$1; shift; env "$@" timeout 300 python bench.py --ncol 512 --steps 3 --warmup 1 --skip-cpu-baseline --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$label: value %.3e  stencil ms %.3f nloc %.3f launches %d'%(d['value'], r['avg_launch_ms'], r['nloc_ms_per_degree'], d['gpu_launches']))
    elif 'rror' in l: print(l.rstrip())
"; }
for rep in 1 2 3; do
run old CHEFSI_B200_LIB=$PWD/scripts/ab/libchefsi_old.so
run new A=1
done
