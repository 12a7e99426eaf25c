mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace_gpu.py -m gpu -x -q -k lanczos > gpurun_out/r2_s10_lanczos_tests.log 2>&1; tail -12 gpurun_out/r2_s10_lanczos_tests.log
timeout 1500 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s > gpurun_out/r2_s10_scf_tests.log 2>&1; tail -10 gpurun_out/r2_s10_scf_tests.log | cut -c1-250
for c in Si8 Au_fcc211 O2_spin_coarse; do bash scripts/run_sparc_case.sh $c 2>&1 | sed "s/^/[$c gpu] /" | grep -E "wall|walltime|Lap_vec|Lanczos|ChebyshevFiltering calls|DP_Project|context creation|Free energy"; done > gpurun_out/r2_s10_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s10_scf.log
run() { tag=$1; lib=$2; CHEFSI_B200_LIB=$lib timeout 600 python bench.py --cell-typ 17 --ncol 256 --steps 2 --warmup 2 --skip-cpu-baseline --no-nloc --e2e-cols 8 2>&1 | python -c "
import sys,json
ok=False
for l in sys.stdin:
    if l.startswith('{'):
        ok=True; d=json.loads(l); r=d['roofline']; print('$tag', 'value %.3e  stencil ms %.3f frac %.3f'%(d['value'], r['avg_launch_ms'], r['frac']), d['clocks']['sm_mhz'], d['clocks']['power_w'])
    elif 'rror' in l: print('$tag', l.rstrip()[:200])
"; }
for rep in 1 2; do
run base sparc_b200/libchefsi_b200.so
run s5 sparc_b200/libchefsi_b200_mix_s5.so
run roll sparc_b200/libchefsi_b200_mix_roll.so
run s5roll sparc_b200/libchefsi_b200_mix_s5roll.so
done
