mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace_gpu.py -m gpu -x -q > gpurun_out/r2_s21_tests.log 2>&1; tail -6 gpurun_out/r2_s21_tests.log | cut -c1-220
timeout 1200 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s > gpurun_out/r2_s21_scf_tests.log 2>&1; tail -12 gpurun_out/r2_s21_scf_tests.log | cut -c1-300
for c in Si8_kpt; do
  bash scripts/run_sparc_case.sh $c 2>&1 | sed "s/^/[$c] /" | grep -E "walltime|Total walltime|DP_Project|AAR|Lanczos calls|ChebyshevFiltering calls|Hamiltonian_vectors_mult calls|Free energy";
  bash scripts/run_sparc_case.sh $c CHEFSI_B200_NO_LANCZOS=1 2>&1 | sed "s/^/[$c NO_LANCZOS] /" | grep -E "walltime|Total walltime|DP_Project|AAR|Lanczos calls|ChebyshevFiltering calls|Hamiltonian_vectors_mult calls|Free energy";
done > gpurun_out/r2_s21_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s21_scf.log
