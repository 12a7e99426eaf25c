# round 2, session 2: mixed (non-orthogonal) streaming kernel -- parity, then the secondary synthetic config
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_stream or zmarch or all_cell_types" > gpurun_out/r2_s2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_s2_tests.log
tail -15 gpurun_out/r2_s2_tests.log
for ct in 17 14 11; do
timeout 600 python bench.py --cell-typ $ct --ncol 256 --steps 2 --warmup 2 --skip-cpu-baseline --e2e-cols 32 > gpurun_out/r2_s2_bench_typ$ct.json 2> gpurun_out/r2_s2_bench_typ$ct.err; tail -c 1500 gpurun_out/r2_s2_bench_typ$ct.json; tail -3 gpurun_out/r2_s2_bench_typ$ct.err
done
CHEFSI_B200_FORCE_GENERAL=1 timeout 600 python bench.py --cell-typ 17 --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 32 > gpurun_out/r2_s2_bench_typ17_zmarch.json 2>&1; tail -c 800 gpurun_out/r2_s2_bench_typ17_zmarch.json
