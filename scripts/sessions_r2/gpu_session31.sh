# compact sphere-value exchange (nlc): parity of the new path, A/B bench against the own-pass projector kernels, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact_exchange or own_pass or dense_stream or golden or real_sparc or bench_problem or device_resident" > gpurun_out/r2_s31_tests.log 2>&1; tail -5 gpurun_out/r2_s31_tests.log
run() { tag=$1; nlc=$2; CHEFSI_B200_NLC=$nlc timeout 600 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s31_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f parity %s'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1, d.get('parity_rel_fro')), d['clocks']['sm_mhz'])
"; }
run nlc1 1
run nlc0 0
run nlc1b 1
CHEFSI_B200_NLC=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_s31_launches_nlc.csv python bench.py --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 > gpurun_out/r2_s31_ncu_run.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_s31_launches_nlc.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4][:70]].append(float(r[-1]))
for k,v in agg.items(): print(k, len(v), 'avg us %.1f'%(sum(v)/len(v)/1e3 if max(v)>1e4 else sum(v)/len(v)))
PY
