# compute-sanitizer memcheck over the kernels touched in the last session (K1 / K1c consumer loops, shared rank projection, multi-device mirror)
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -k "dense_stream_kernel_vs_oracle or kpt_stream_kernel or fused_projector or empty_block or multi_device_subspace" > gpurun_out/r2_s53_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s53_memcheck.log; grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2_s53_memcheck.log | tail -5
