# new reference-made golden fixture (multi-tile grid with shifted tiles, disjoint spheres) on the GPU
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > gpurun_out/r2_s56_tests.log 2>&1; tail -2 gpurun_out/r2_s56_tests.log
