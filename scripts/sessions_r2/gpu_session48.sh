# projector kernels with the sphere point lists rewritten as aligned x pairs (16-byte gather / scatter): parity + A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unpaired or fused_projector or golden or real_sparc or dense_stream or bench_problem or empty_block or overlapping or hamiltonian_all or mixed_stream_many or device_resident" > gpurun_out/r2_s48_tests.log 2>&1; tail -4 gpurun_out/r2_s48_tests.log | cut -c1-300
run() { tag=$1; shift; env "$@" timeout 150 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s48_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run pairs CHEFSI_B200_NLOC_PAIRS=1
run nopairs CHEFSI_B200_NLOC_PAIRS=0
run pairs2 CHEFSI_B200_NLOC_PAIRS=1
run nopairs2 CHEFSI_B200_NLOC_PAIRS=0
