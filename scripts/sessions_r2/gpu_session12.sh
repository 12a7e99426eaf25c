mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace_gpu.py -m gpu -x -q -k "resident" > gpurun_out/r2_s12_tests.log 2>&1; tail -25 gpurun_out/r2_s12_tests.log | cut -c1-200
bash scripts/run_sparc_case.sh Si8_kpt CHEFSI_B200_NO_SUBSPACE=1 2>&1 | sed "s/^/[Si8_kpt no-subspace] /" | grep -E "wall|walltime|DP_Project|Free energy|NaN|ERROR" | cut -c1-260
bash scripts/run_sparc_case.sh Si8_kpt CHEFSI_B200_SHIM_VERBOSE=2 2>&1 | sed "s/^/[Si8_kpt subspace] /" | tail -30 | cut -c1-260
