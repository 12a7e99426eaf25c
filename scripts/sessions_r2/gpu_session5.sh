mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --ncol 256 --steps 1 --warmup 1 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$tag', 'e2e %.3e  %.4f s  link %.1f GB/s  value %.3e'%(e['value'], e['seconds'], e['host_link_gbs_per_direction'], d['value']))
"; }
run default A=1
run chunk16 CHEFSI_B200_HOST_CHUNK=16
run chunk64 CHEFSI_B200_HOST_CHUNK=64
run chunk128 CHEFSI_B200_HOST_CHUNK=128
run flat32 CHEFSI_B200_FLAT_CHUNKS=1
run flat64 CHEFSI_B200_FLAT_CHUNKS=1 CHEFSI_B200_HOST_CHUNK=64
run flat16 CHEFSI_B200_FLAT_CHUNKS=1 CHEFSI_B200_HOST_CHUNK=16
