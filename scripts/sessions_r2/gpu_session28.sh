mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s28_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s28_tests.log; tail -6 gpurun_out/r2_s28_tests.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_s28_smoke.log 2>&1; tail -2 gpurun_out/r2_s28_smoke.log
timeout 900 python bench.py > gpurun_out/r2_s28_bench.json 2> gpurun_out/r2_s28_bench.err; tail -c 3000 gpurun_out/r2_s28_bench.json; tail -3 gpurun_out/r2_s28_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_s28_bench_ref.json 2> gpurun_out/r2_s28_bench_ref.err; tail -c 1200 gpurun_out/r2_s28_bench_ref.json
