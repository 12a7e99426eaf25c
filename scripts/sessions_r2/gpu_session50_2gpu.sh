# N = 2 bench line of the final tree (one process per GPU under torchrun, NCCL broadcast of Veff / projector tables)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_s50_bench_n2.json 2> gpurun_out/r2_s50_bench_n2.err; tail -c 600 gpurun_out/r2_s50_bench_n2.json; tail -2 gpurun_out/r2_s50_bench_n2.err | cut -c1-300
