# final tree: full GPU suite, smoke, default N=1 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s49_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s49_tests.log; tail -4 gpurun_out/r2_s49_tests.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2_s49_smoke.log 2>&1; tail -1 gpurun_out/r2_s49_smoke.log
timeout 600 python bench.py > gpurun_out/r2_s49_bench.json 2> gpurun_out/r2_s49_bench.err; tail -c 1500 gpurun_out/r2_s49_bench.json | cut -c1-1500; tail -2 gpurun_out/r2_s49_bench.err
