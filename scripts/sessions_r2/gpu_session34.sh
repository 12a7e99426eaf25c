# exchange epilogue pipelined across planes: parity + A/B; ncu of the compact projector kernels and of the exchange kernel
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact_exchange or own_pass or golden or bench_problem" > gpurun_out/r2_s34_tests.log 2>&1; tail -3 gpurun_out/r2_s34_tests.log
run() { tag=$1; shift; env "$@" timeout 150 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s34_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run nlc1 CHEFSI_B200_NLC=1
run nlc0 CHEFSI_B200_NLC=0
CHEFSI_B200_NLC=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:nloc_kernel -s 2 -c 2 -o gpurun_out/prof_s34_nloc_compact python bench.py --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 > gpurun_out/prof_s34_nloc_compact.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_s34_nloc_compact.ncu-rep > gpurun_out/prof_s34_nloc_compact.txt 2>&1
bash scripts/ncu_kernel.sh s34_dense_nlc "stream_dense_kernel" 6 CHEFSI_B200_NLC=1
cat gpurun_out/prof_s34_nloc_compact.txt gpurun_out/prof_s34_dense_nlc.txt
