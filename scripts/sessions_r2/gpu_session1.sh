# round 2, first GPU session: parity tests, bench line, ncu captures of K1 and K3 (traffic), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_s1_gpu.txt
nproc >> gpurun_out/r2_s1_gpu.txt; numactl -H >> gpurun_out/r2_s1_gpu.txt 2>&1; nvidia-smi topo -m >> gpurun_out/r2_s1_gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_s1_tests.log
tail -5 gpurun_out/r2_s1_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_s1_bench.json 2> gpurun_out/r2_s1_bench.err; tail -c 3000 gpurun_out/r2_s1_bench.json; tail -5 gpurun_out/r2_s1_bench.err
bash scripts/ncu_dense.sh r2_dense
bash scripts/ncu_kernel.sh r2_nloc "nloc_kernel" 6
python profiles/ncu_summary.py gpurun_out/prof_r2_dense.ncu-rep > gpurun_out/prof_r2_dense.txt 2>&1
