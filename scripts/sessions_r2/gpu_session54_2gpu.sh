# band-parallel projection: antipodal block of an even rank count shared half and half; tests + timing on 2 GPUs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_band_parallel_gpu.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_s54_tests.log 2>&1; tail -4 gpurun_out/r2_s54_tests.log | cut -c1-300

timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/band_parallel_bench.py 96 512 > gpurun_out/r2_s54_bp_n2.log 2>&1; tail -1 gpurun_out/r2_s54_bp_n2.log | cut -c1-500
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/band_parallel_bench.py 128 512 > gpurun_out/r2_s54_bp_n2_128.log 2>&1; tail -1 gpurun_out/r2_s54_bp_n2_128.log | cut -c1-500
