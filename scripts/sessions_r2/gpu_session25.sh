mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_subspace_gpu.py tests/test_rayleigh_ritz_gpu.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_s25_tests.log 2>&1; tail -15 gpurun_out/r2_s25_tests.log | cut -c1-250
timeout 600 python scripts/subspace_bench.py > gpurun_out/r2_s25_subspace_bench.log 2>&1; cat gpurun_out/r2_s25_subspace_bench.log | cut -c1-250
CHEFSI_B200_GEMM_BIG_TILES=0 timeout 600 python scripts/subspace_bench.py 96 512 > gpurun_out/r2_s25_subspace_bench_small_tiles.log 2>&1; cat gpurun_out/r2_s25_subspace_bench_small_tiles.log | cut -c1-250
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 5 -o gpurun_out/r2_s25_gemm -f python scripts/subspace_bench.py 96 512 > gpurun_out/r2_s25_ncu.log 2>&1; tail -3 gpurun_out/r2_s25_ncu.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/r2_s25_gemm.ncu-rep > gpurun_out/r2_s25_gemm_summary.txt 2>&1; grep -E "^---|time_duration|dram__bytes|fp64|issue_active|math_pipe|short_score|long_score|barrier|stalled_wait|registers_per|warps_active|hit_rate" gpurun_out/r2_s25_gemm_summary.txt | cut -c1-160
