run() { label=$1; shift; env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label: value %.3e e2e %.3e  %.4f s'%(d['value'], d['e2e']['value'], d['e2e']['seconds']))
    elif 'rror' in l or 'Trace' in l: print(l.rstrip())
"; }
for rep in 1 2; do
run flat64 CHEFSI_B200_FLAT_CHUNKS=1 CHEFSI_B200_HOST_CHUNK=64
run ramp32 A=1
run ramp64 CHEFSI_B200_HOST_CHUNK=64
done
