mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2; nproc
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_s17_bench_n8.json 2> gpurun_out/r2_s17_bench_n8.err; tail -c 2500 gpurun_out/r2_s17_bench_n8.json; tail -3 gpurun_out/r2_s17_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/r2_s17_ref_n8.json 2>&1; tail -c 600 gpurun_out/r2_s17_ref_n8.json
timeout 900 python scripts/multi_ctx_bench.py > gpurun_out/r2_s17_multi_ctx_bench.log 2>&1; cat gpurun_out/r2_s17_multi_ctx_bench.log
