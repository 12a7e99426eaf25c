# compare library builds / configs on the same box, interleaved, stencil-only bench (512 columns)
# usage: bash scripts/exp_ab.sh "label ENV=.. ENV=.." "label2 ENV=.." ...
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --ncol 512 --steps 6 --warmup 1 --skip-cpu-baseline --no-nloc --e2e-cols 16 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$label: stencil avg ms %.3f  frac %.3f  value %.3e  sm_mhz %s power %s W'%(r['avg_launch_ms'], r['frac'], d['value'], d['clocks']['sm_mhz'], d['clocks'].get('power_w')))
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"
}
for rep in 1 2; do
  for spec in "$@"; do run $spec; done
done
