mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradient or lap_vec" > gpurun_out/r2_s26_tests.log 2>&1; tail -8 gpurun_out/r2_s26_tests.log | cut -c1-250
timeout 300 python scripts/gradient_bench.py > gpurun_out/r2_s26_gradient_bench.log 2>&1; cat gpurun_out/r2_s26_gradient_bench.log | cut -c1-200
timeout 900 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s -k "gradients or energy_matches" > gpurun_out/r2_s26_scf_tests.log 2>&1; tail -14 gpurun_out/r2_s26_scf_tests.log | cut -c1-300
for c in Si8 BaTiO3; do
  for v in A=1 CHEFSI_B200_NO_GRAD=1; do
  bash scripts/run_sparc_case.sh $c $v 2>&1 | sed "s/^/[$c $v] /" | grep -E "Total walltime|Gradient_vectors_dir|Free energy per atom  ";
  done
done > gpurun_out/r2_s26_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s26_scf.log
