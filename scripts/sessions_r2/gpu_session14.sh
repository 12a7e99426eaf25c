mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s14_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s14_tests.log; tail -8 gpurun_out/r2_s14_tests.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_s14_smoke.log 2>&1; tail -2 gpurun_out/r2_s14_smoke.log
CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_mix_arrive.so timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(mixed_stream_kernel and 17 and N0)" > gpurun_out/r2_s14_racecheck_arrive.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_s14_racecheck_arrive.log
CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_mix_arrive.so timeout 600 python bench.py --cell-typ 17 --ncol 256 --steps 2 --warmup 2 --skip-cpu-baseline --no-nloc --e2e-cols 8 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('arrive-all', 'value %.3e  stencil ms %.3f'%(d['value'], r['avg_launch_ms']))
"
