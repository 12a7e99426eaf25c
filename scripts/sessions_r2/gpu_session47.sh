# multi-device context: Hermitian block pairs of Hp / Mp formed once (devices listed twice on a one-GPU box)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py tests/test_band_parallel_gpu.py tests/test_subspace_gpu.py tests/test_rayleigh_ritz_gpu.py -m gpu -x -q > gpurun_out/r2_s47_tests.log 2>&1; tail -4 gpurun_out/r2_s47_tests.log | cut -c1-300
timeout 300 python scripts/multi_ctx_bench.py > gpurun_out/r2_s47_multi_ctx.log 2>&1; tail -4 gpurun_out/r2_s47_multi_ctx.log | cut -c1-300
