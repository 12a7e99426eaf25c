# K1: steady path as one block (xprev loads early, no zero init): parity + A/B against the previous commit
# template flags): parity + A/B against the previous build
mkdir -p gpurun_out
timeout 250 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_stream or golden or fused_projector or lap_vec or empty_block or no_projectors or bench_problem or leading_dimension or degree_one" > gpurun_out/r2_s39_tests.log 2>&1; tail -3 gpurun_out/r2_s39_tests.log
run() { tag=$1; lib=$2; CHEFSI_B200_LIB=$lib timeout 150 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s39_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run base sparc_b200/libchefsi_b200_base.so
run lean sparc_b200/libchefsi_b200.so
run base2 sparc_b200/libchefsi_b200_base.so
run lean2 sparc_b200/libchefsi_b200.so
