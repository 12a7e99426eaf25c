# SURVEY.md 8d secondary synthetic configurations: k-point (complex) and cell_typ 17 on the 160^3 workload (64 columns)
run() { label=$1; shift; timeout 600 python bench.py --ncol 64 --block 64 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$label: value %.3e  %s avg ms %.3f frac %.3f nloc ms/degree %.3f e2e %.3e'%(d['value'], r['kernel'][:28], r['avg_launch_ms'], r['frac'], r['nloc_ms_per_degree'], d['e2e']['value']))
    elif 'rror' in l or 'Trace' in l: print(l.rstrip())
"; }
run real_orth
run kpt_orth --kpt
run real_typ17 --cell-typ 17
run kpt_typ17 --kpt --cell-typ 17
