# projector CTAs in order of decreasing segment size: parity + A/B
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_projector or golden or overlapping or bench_problem or real_sparc or empty_block" > gpurun_out/r2_s57_tests.log 2>&1; tail -2 gpurun_out/r2_s57_tests.log
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --ncol 512 --steps 2 --warmup 2 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s57_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run sorted CHEFSI_B200_NLOC_SORT=1
run unsorted CHEFSI_B200_NLOC_SORT=0
run sorted2 CHEFSI_B200_NLOC_SORT=1
