mkdir -p gpurun_out
run() { tag=$1; lib=$2; CHEFSI_B200_LIB=$lib timeout 600 python bench.py --cell-typ 17 --ncol 256 --steps 2 --warmup 2 --skip-cpu-baseline --no-nloc --e2e-cols 8 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f'%(d['value'], r['avg_launch_ms'], r['frac']), d['clocks']['sm_mhz'], d['clocks']['power_w'])
"; }
for rep in 1 2; do
run base sparc_b200/libchefsi_b200.so
run s5 sparc_b200/libchefsi_b200_mix_s5.so
run roll sparc_b200/libchefsi_b200_mix_roll.so
run s5roll sparc_b200/libchefsi_b200_mix_s5roll.so
done
# does racecheck flag the same mbarrier hand-over (merge warps -> consumers) in the round-1 K1 kernel?  (tool limitation check)
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_stream_kernel_vs_oracle and N0" > gpurun_out/r2_s9_racecheck_k1.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_s9_racecheck_k1.log; grep -E "Error: Race" gpurun_out/r2_s9_racecheck_k1.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | head -5 | cut -c1-250
