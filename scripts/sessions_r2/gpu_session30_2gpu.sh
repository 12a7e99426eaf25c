mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8 > gpurun_out/r2_s30_topo.txt
timeout 900 python -m pytest tests/test_band_parallel_gpu.py tests/test_multi_gpu.py tests/test_domain_split.py -m gpu -x -q > gpurun_out/r2_s30_tests.log 2>&1; tail -6 gpurun_out/r2_s30_tests.log | cut -c1-300
timeout 300 python scripts/band_parallel_bench.py 96 512 > gpurun_out/r2_s30_bp_n1.log 2>&1; tail -2 gpurun_out/r2_s30_bp_n1.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/band_parallel_bench.py 96 512 > gpurun_out/r2_s30_bp_n2.log 2>&1; tail -3 gpurun_out/r2_s30_bp_n2.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/band_parallel_bench.py 128 512 > gpurun_out/r2_s30_bp_n2_128.log 2>&1; tail -3 gpurun_out/r2_s30_bp_n2_128.log | cut -c1-400
