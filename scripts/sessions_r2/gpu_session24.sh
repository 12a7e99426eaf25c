mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 5 -o gpurun_out/r2_s24_gemm -f python scripts/subspace_bench.py 96 512 > gpurun_out/r2_s24_ncu.log 2>&1; tail -5 gpurun_out/r2_s24_ncu.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/r2_s24_gemm.ncu-rep > gpurun_out/r2_s24_gemm_summary.txt 2>&1; cat gpurun_out/r2_s24_gemm_summary.txt | cut -c1-200
