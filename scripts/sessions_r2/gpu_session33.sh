# branch-free exchange epilogue: parity + A/B; L2 fetch granularity hint for the own-pass projector kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact_exchange or own_pass or golden or bench_problem" > gpurun_out/r2_s33_tests.log 2>&1; tail -3 gpurun_out/r2_s33_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s33_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run nlc1 CHEFSI_B200_NLC=1
run nlc0 CHEFSI_B200_NLC=0
run nlc0_l2f32 CHEFSI_B200_NLC=0 CHEFSI_B200_L2_FETCH=32
run nlc0_l2f128 CHEFSI_B200_NLC=0 CHEFSI_B200_L2_FETCH=128
run nlc1_l2f32 CHEFSI_B200_NLC=1 CHEFSI_B200_L2_FETCH=32
