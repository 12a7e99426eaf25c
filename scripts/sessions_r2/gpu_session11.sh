mkdir -p gpurun_out
free -g | head -2 > gpurun_out/r2_s11_hostmem.txt; cat gpurun_out/r2_s11_hostmem.txt
timeout 900 python -m pytest tests/test_subspace_gpu.py -m gpu -x -q > gpurun_out/r2_s11_subspace_tests.log 2>&1; tail -12 gpurun_out/r2_s11_subspace_tests.log
timeout 900 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s -k "Si8_kpt or Si8" > gpurun_out/r2_s11_scf_tests.log 2>&1; tail -6 gpurun_out/r2_s11_scf_tests.log | cut -c1-250
bash scripts/run_sparc_case.sh Si8_kpt 2>&1 | sed "s/^/[Si8_kpt gpu] /" | grep -E "wall|walltime|Lap_vec|Lanczos|ChebyshevFiltering calls|DP_Project|context creation|Free energy" | cut -c1-260
# racecheck: does it accept the hand-over when every merge thread arrives on the mbarrier itself?
CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_mix_arrive.so timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(mixed_stream_kernel and 17 and N0)" > gpurun_out/r2_s11_racecheck_arrive.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_s11_racecheck_arrive.log
CHEFSI_B200_LIB=sparc_b200/libchefsi_b200_mix_arrive.so timeout 600 python bench.py --cell-typ 17 --ncol 256 --steps 2 --warmup 2 --skip-cpu-baseline --no-nloc --e2e-cols 8 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('arrive-all', 'value %.3e  stencil ms %.3f'%(d['value'], r['avg_launch_ms']))
"
