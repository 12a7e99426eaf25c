# final tree: full GPU suite + smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_s58_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_s58_tests.log; tail -4 gpurun_out/r2_s58_tests.log | cut -c1-200
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2_s58_smoke.log 2>&1; tail -1 gpurun_out/r2_s58_smoke.log
