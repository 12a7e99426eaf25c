# launch list of the bench command (gpu__time_duration per launch): the first 300 launches = warm-up, timed and profiled step at 128 columns per launch
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_bench_ncol256_final.csv python bench.py --ncol 256 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 > gpurun_out/r2_s52_launches.log 2>&1; tail -1 gpurun_out/r2_s52_launches.log | cut -c1-200
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_bench_ncol256_final.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4][:64]].append(float(r[-1]))
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items(): print(k, len(v), 'avg %.1f us'%(sum(v)/len(v)/1e3), 'share %.3f'%(sum(v)/tot))
PY
