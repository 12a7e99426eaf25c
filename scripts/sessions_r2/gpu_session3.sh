# round 2, session 3: Lap_vec_mult on the GPU (Poisson residual) -- parity + the SCF table of the four test systems
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lap_vec_mult" > gpurun_out/r2_s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_s3_tests.log
tail -4 gpurun_out/r2_s3_tests.log
timeout 1500 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s > gpurun_out/r2_s3_scf_tests.log 2>&1; echo "scf tests rc=$?" >> gpurun_out/r2_s3_scf_tests.log
tail -8 gpurun_out/r2_s3_scf_tests.log
for c in Si8 BaTiO3 Au_fcc211 Si8_kpt; do
  bash scripts/run_sparc_case.sh $c CHEFSI_B200_SHIM_VERBOSE=2 2>&1 | sed "s/^/[$c gpu] /" | tail -12
  bash scripts/run_sparc_case.sh $c CHEFSI_B200_NO_LAP=1 2>&1 | sed "s/^/[$c gpu, Lap_vec_mult on CPU] /" | tail -6
done > gpurun_out/r2_s3_scf_table.log 2>&1
cat gpurun_out/r2_s3_scf_table.log | cut -c1-400
