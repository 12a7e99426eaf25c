# validation of the tree after the K1 consumer-loop work: full GPU suite, smoke, ncu capture + launch list, N=1 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s43_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s43_tests.log; tail -4 gpurun_out/r2_s43_tests.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2_s43_smoke.log 2>&1; tail -2 gpurun_out/r2_s43_smoke.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_dense_kernel -s 6 -c 1 -o gpurun_out/prof_r2_dense_lean python bench.py --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 > gpurun_out/prof_r2_dense_lean.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_r2_dense_lean.ncu-rep > gpurun_out/prof_r2_dense_lean.txt 2>&1; head -12 gpurun_out/prof_r2_dense_lean.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches_bench_ncol256_lean.csv python bench.py --ncol 256 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 > gpurun_out/r2_s43_launches.log 2>&1; tail -1 gpurun_out/r2_s43_launches.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_s43_bench.json 2> gpurun_out/r2_s43_bench.err; tail -c 3000 gpurun_out/r2_s43_bench.json; tail -3 gpurun_out/r2_s43_bench.err
