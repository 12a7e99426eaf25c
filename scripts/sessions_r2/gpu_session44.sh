# K1 with the xprev operand read straight from global memory (L2 prefetch by the producer) instead of through shared memory
mkdir -p gpurun_out
for lib in xldg xldg6; do CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_$lib.so timeout 250 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_stream or golden or fused_projector or lap_vec or empty_block or bench_problem or degree_one" > gpurun_out/r2_s44_tests_$lib.log 2>&1; tail -2 gpurun_out/r2_s44_tests_$lib.log; done
run() { tag=$1; lib=$2; CHEFSI_B200_LIB=$lib timeout 150 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s44_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run base sparc_b200/libchefsi_b200.so
run xldg sparc_b200/libchefsi_b200_xldg.so
run xldg6 sparc_b200/libchefsi_b200_xldg6.so
run base2 sparc_b200/libchefsi_b200.so
run xldg_2 sparc_b200/libchefsi_b200_xldg.so
run xldg6_2 sparc_b200/libchefsi_b200_xldg6.so
