mkdir -p gpurun_out
# ncu of the non-orthogonal streaming kernel (cell_typ 17)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stream_mixed" -s 4 -c 1 -o gpurun_out/prof_r2_mixed17 \
    python bench.py --cell-typ 17 --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --no-nloc --e2e-cols 8 > gpurun_out/prof_r2_mixed17.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_r2_mixed17.ncu-rep > gpurun_out/prof_r2_mixed17.txt 2>&1; cat gpurun_out/prof_r2_mixed17.txt
# sanitizers over the new kernels' parity tests
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_subspace_gpu.py tests/test_multi_gpu.py -m gpu -x -q -k "mixed_stream_many or (mixed_stream_kernel and 17) or subspace or multi_device_filter or overlapping_spheres or lap_vec" > gpurun_out/r2_s8_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s8_memcheck.log; tail -6 gpurun_out/r2_s8_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_subspace_gpu.py -m gpu -x -q -k "(mixed_stream_kernel and 17 and N0) or (mixed_stream_kernel and 11 and N1) or project_and_rotate_small" > gpurun_out/r2_s8_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s8_racecheck.log; tail -6 gpurun_out/r2_s8_racecheck.log
