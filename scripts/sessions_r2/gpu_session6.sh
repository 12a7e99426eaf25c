mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace_gpu.py -m gpu -x -q > gpurun_out/r2_s6_subspace_tests.log 2>&1; tail -15 gpurun_out/r2_s6_subspace_tests.log
timeout 900 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s > gpurun_out/r2_s6_scf_tests.log 2>&1; tail -8 gpurun_out/r2_s6_scf_tests.log | cut -c1-250
for c in Si8 BaTiO3 Au_fcc211; do bash scripts/run_sparc_case.sh $c 2>&1 | sed "s/^/[$c gpu] /" | grep -E "wall|walltime|Lap_vec|ChebyshevFiltering calls|DP_Project|context creation"; done > gpurun_out/r2_s6_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s6_scf.log
timeout 600 python scripts/subspace_bench.py > gpurun_out/r2_s6_subspace_bench.log 2>&1; cat gpurun_out/r2_s6_subspace_bench.log
