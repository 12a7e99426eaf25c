# launch-group size and e2e block size A/B
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 280 python bench.py --ncol 1024 --steps 2 --warmup 2 --skip-cpu-baseline "$@" 2>&1 | tee gpurun_out/r2_s36_bench_$tag.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; e=d['e2e']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f nloc ms/deg %.3f | e2e %.3e cols %d link %.1f GB/s'%(d['value'], r['avg_launch_ms'], r['frac'], r.get('nloc_ms_per_degree') or -1, e['value'], e['columns_per_rank'], e['host_link_gbs_per_direction']), d['clocks']['sm_mhz'])
"; }
run b128_e256 --block 128 --e2e-cols 256
run b256_e512 --block 256 --e2e-cols 512
run b64_e1024 --block 64 --e2e-cols 1024
