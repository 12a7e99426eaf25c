# K1c (k-point) and K2m (non-orthogonal) streaming kernels with the leaner consumer loop: parity + A/B against HEAD
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kpt_stream or mixed_stream or golden or real_sparc or dense_stream or hamiltonian_all" > gpurun_out/r2_s42_tests.log 2>&1; tail -3 gpurun_out/r2_s42_tests.log
run() { tag=$1; lib=$2; shift 2; CHEFSI_B200_LIB=$lib timeout 200 python bench.py --ncol 256 --steps 2 --warmup 2 --skip-cpu-baseline --e2e-cols 8 "$@" 2>&1 | tee gpurun_out/r2_s42_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run kpt_base sparc_b200/libchefsi_b200_base.so --kpt
run kpt_lean sparc_b200/libchefsi_b200.so --kpt
run t17_base sparc_b200/libchefsi_b200_base.so --cell-typ 17
run t17_lean sparc_b200/libchefsi_b200.so --cell-typ 17
run kpt_base2 sparc_b200/libchefsi_b200_base.so --kpt
run kpt_lean2 sparc_b200/libchefsi_b200.so --kpt
run t17_base2 sparc_b200/libchefsi_b200_base.so --cell-typ 17
run t17_lean2 sparc_b200/libchefsi_b200.so --cell-typ 17
