run() { label=$1; shift; env "$@" timeout 600 python bench.py --ncol 256 --steps 1 --warmup 1 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$label: e2e %.3e  %.4f s'%(d['e2e']['value'], d['e2e']['seconds']))
    elif 'rror' in l or 'Trace' in l: print(l.rstrip())
"; }
for rep in 1 2; do
run flat64 CHEFSI_B200_FLAT_CHUNKS=1
run ramp64 A=1
run ramp32 CHEFSI_B200_HOST_CHUNK=32
run ramp128 CHEFSI_B200_HOST_CHUNK=128
done
