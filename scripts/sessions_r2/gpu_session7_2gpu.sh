# two B200s: NCCL paths (multi-device context broadcast, domain split send/recv, bench band split), multi-GPU SCF
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_s7_gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_domain_split.py -m gpu -x -q > gpurun_out/r2_s7_tests_2gpu.log 2>&1; tail -5 gpurun_out/r2_s7_tests_2gpu.log
timeout 900 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s -k multi_device > gpurun_out/r2_s7_scf_multi.log 2>&1; tail -5 gpurun_out/r2_s7_scf_multi.log | cut -c1-300
for c in BaTiO3 Au_fcc211; do
  bash scripts/run_sparc_case.sh $c CHEFSI_B200_DEVICES=0,1 2>&1 | sed "s/^/[$c 2 GPUs] /" | grep -E "wall|walltime|devices|ChebyshevFiltering calls|Free energy"
done > gpurun_out/r2_s7_scf_2gpu.log 2>&1; cut -c1-260 gpurun_out/r2_s7_scf_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_s7_bench_n2.json 2> gpurun_out/r2_s7_bench_n2.err; tail -c 1800 gpurun_out/r2_s7_bench_n2.json; tail -3 gpurun_out/r2_s7_bench_n2.err
# one process, one multi-device context, host block of 512 columns: the drop-in's multi-GPU path on the headline workload
timeout 900 python scripts/multi_ctx_bench.py > gpurun_out/r2_s7_multi_ctx_bench.log 2>&1; cat gpurun_out/r2_s7_multi_ctx_bench.log
