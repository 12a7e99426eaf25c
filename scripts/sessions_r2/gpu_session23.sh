mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rayleigh_ritz_gpu.py -m gpu -x -q > gpurun_out/r2_s23_tests.log 2>&1; tail -25 gpurun_out/r2_s23_tests.log | cut -c1-250
timeout 1200 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s -k rayleigh > gpurun_out/r2_s23_scf_tests.log 2>&1; tail -8 gpurun_out/r2_s23_scf_tests.log | cut -c1-300
for c in Si8_kpt Au_fcc211; do
  for v in A=1 CHEFSI_B200_NO_DENSITY=1; do
  bash scripts/run_sparc_case.sh $c $v 2>&1 | sed "s/^/[$c $v] /" | grep -E "Total walltime|subspace eigenproblems|Free energy per atom  ";
  done
done > gpurun_out/r2_s23_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s23_scf.log
