mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace_gpu.py -m gpu -x -q > gpurun_out/r2_s13_tests.log 2>&1; tail -25 gpurun_out/r2_s13_tests.log | cut -c1-220
timeout 1500 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s > gpurun_out/r2_s13_scf_tests.log 2>&1; tail -10 gpurun_out/r2_s13_scf_tests.log | cut -c1-250
for c in Si8 BaTiO3 Au_fcc211 Si8_kpt O2_spin_coarse; do bash scripts/run_sparc_case.sh $c 2>&1 | sed "s/^/[$c gpu] /" | grep -E "wall|walltime|Lap_vec|Lanczos|AAR|ChebyshevFiltering calls|DP_Project|context creation|Free energy"; done > gpurun_out/r2_s13_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s13_scf.log
