# projector kernel pipeline shapes re-measured on the round-2 kernel (fused mode, NP = 20)
mkdir -p gpurun_out
run() { tag=$1; shape=$2; CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_nlshape.so CHEFSI_B200_NLOC_SHAPE=$shape timeout 150 python bench.py --ncol 512 --steps 2 --warmup 2 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s46_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_nlshape.so CHEFSI_B200_NLOC_SHAPE=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bench_problem" 2>&1 | tail -1
run s0_256x2x64 0
run s1_256x3x64 1
run s2_256x2x128 2
run s3_256x3x32 3
run s4_128x3x64 4
run s5_256x4x32 5
run s0_again 0
