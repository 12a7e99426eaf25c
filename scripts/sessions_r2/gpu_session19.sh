mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_s19_tests.log 2>&1; tail -6 gpurun_out/r2_s19_tests.log | cut -c1-220
timeout 900 python -m pytest tests/test_sparc_scf_gpu.py -m gpu -x -q -s -k multi_device > gpurun_out/r2_s19_scf_multi.log 2>&1; tail -4 gpurun_out/r2_s19_scf_multi.log | cut -c1-300
for c in BaTiO3 Au_fcc211; do bash scripts/run_sparc_case.sh $c CHEFSI_B200_DEVICES=0,0 2>&1 | sed "s/^/[$c devices 0,0] /" | grep -E "walltime|devices|DP_Project|AAR|Lanczos calls|ChebyshevFiltering calls|Free energy"; done > gpurun_out/r2_s19_scf.log 2>&1; cut -c1-260 gpurun_out/r2_s19_scf.log
