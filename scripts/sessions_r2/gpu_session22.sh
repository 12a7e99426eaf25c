mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rayleigh_ritz_gpu.py tests/test_subspace_gpu.py -m gpu -x -q > gpurun_out/r2_s22_tests.log 2>&1; tail -25 gpurun_out/r2_s22_tests.log | cut -c1-250
OMP_NUM_THREADS=1 timeout 600 python scripts/eig_latency.py > gpurun_out/r2_s22_eig_latency.log 2>&1; cat gpurun_out/r2_s22_eig_latency.log | cut -c1-200
timeout 1200 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s > gpurun_out/r2_s22_scf_tests.log 2>&1; tail -16 gpurun_out/r2_s22_scf_tests.log | cut -c1-300
