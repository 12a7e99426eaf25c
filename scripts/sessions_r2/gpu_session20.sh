mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --ncol 256 --steps 1 --warmup 1 --skip-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$tag', 'e2e %.3e  %.4f s  link %.1f GB/s  value %.3e'%(e['value'], e['seconds'], e['host_link_gbs_per_direction'], d['value']))
"; }
run default A=1
run s1 CHEFSI_B200_HOST_CHUNK=32 CHEFSI_B200_CHUNK_SCHED=11,21,32,32,32,32,32,32,21,11
run s2 CHEFSI_B200_HOST_CHUNK=32 CHEFSI_B200_CHUNK_SCHED=11,29,29,29,29,29,29,29,29,13
run s3 CHEFSI_B200_HOST_CHUNK=64 CHEFSI_B200_CHUNK_SCHED=11,23,29,59,59,29,23,12,11
run s4 CHEFSI_B200_HOST_CHUNK=32 CHEFSI_B200_CHUNK_SCHED=5,11,23,29,29,29,29,29,29,23,11,9
run s5 CHEFSI_B200_HOST_CHUNK=32 CHEFSI_B200_CHUNK_SCHED=11,17,29,29,29,29,29,29,29,14,11
run default2 A=1
