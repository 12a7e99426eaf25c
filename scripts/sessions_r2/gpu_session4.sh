mkdir -p gpurun_out
timeout 600 python scripts/small_call_latency.py > gpurun_out/r2_s4_latency_fast.log 2>&1; cat gpurun_out/r2_s4_latency_fast.log
CHEFSI_B200_FAST_SMALL=0 timeout 600 python scripts/small_call_latency.py > gpurun_out/r2_s4_latency_slow.log 2>&1; cat gpurun_out/r2_s4_latency_slow.log
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_s4_multi_tests.log 2>&1; tail -5 gpurun_out/r2_s4_multi_tests.log
timeout 900 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -k multi_device > gpurun_out/r2_s4_multi_scf.log 2>&1; tail -5 gpurun_out/r2_s4_multi_scf.log
for c in Si8 BaTiO3 Au_fcc211; do bash scripts/run_sparc_case.sh $c 2>&1 | sed "s/^/[$c gpu] /" | grep -E "wall|walltime|Lap_vec|ChebyshevFiltering calls|context creation"; done > gpurun_out/r2_s4_scf.log 2>&1; cut -c1-300 gpurun_out/r2_s4_scf.log
