mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_band_parallel_gpu.py -m gpu -x -q > gpurun_out/r2_s29_tests.log 2>&1; tail -30 gpurun_out/r2_s29_tests.log | cut -c1-300
