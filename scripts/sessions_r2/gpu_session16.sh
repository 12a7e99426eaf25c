mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s16_tests.log; tail -6 gpurun_out/r2_s16_tests.log | cut -c1-200
# launch list of the bench command (kernel shares of a step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches_bench_ncol256.csv python bench.py --ncol 256 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/r2_s16_launches.log 2>&1; tail -2 gpurun_out/r2_s16_launches.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_s16_bench.json 2> gpurun_out/r2_s16_bench.err; tail -c 2500 gpurun_out/r2_s16_bench.json; tail -3 gpurun_out/r2_s16_bench.err
timeout 600 python bench.py --kpt --ncol 256 --block 64 --steps 2 --warmup 2 --skip-cpu-baseline --e2e-cols 32 > gpurun_out/r2_s16_bench_kpt.json 2>&1; tail -c 600 gpurun_out/r2_s16_bench_kpt.json
