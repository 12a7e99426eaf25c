# L2 prefetch-size qualifier on the projector kernel's 8-byte sphere gather: A/B
mkdir -p gpurun_out
run() { tag=$1; lib=$2; CHEFSI_B200_LIB=$lib timeout 150 python bench.py --ncol 512 --steps 2 --warmup 3 --skip-cpu-baseline --e2e-cols 8 2>&1 | tee gpurun_out/r2_s35_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1), d['clocks']['sm_mhz'])
"; }
run base sparc_b200/libchefsi_b200.so
run nl64 sparc_b200/libchefsi_b200_nl64.so
run nl128 sparc_b200/libchefsi_b200_nl128.so
run nl256 sparc_b200/libchefsi_b200_nl256.so
run base2 sparc_b200/libchefsi_b200.so
