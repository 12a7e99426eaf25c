mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradient" > gpurun_out/r2_s27_tests.log 2>&1; tail -8 gpurun_out/r2_s27_tests.log | cut -c1-250
timeout 300 python scripts/gradient_bench.py > gpurun_out/r2_s27_gradient_bench.log 2>&1; cat gpurun_out/r2_s27_gradient_bench.log | cut -c1-200
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradient" > gpurun_out/r2_s27_memcheck.log 2>&1; tail -5 gpurun_out/r2_s27_memcheck.log | cut -c1-250
