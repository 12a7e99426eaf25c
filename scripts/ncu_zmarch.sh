tag=$1; shift
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-stencil_zmarch}" -s 3 -c 1 -o gpurun_out/prof_${tag} \
    python bench.py --ncol 64 --block 64 --steps 1 --warmup 0 --skip-cpu-baseline --e2e-cols 8 --no-nloc "$@" > gpurun_out/prof_${tag}.log 2>&1
tail -1 gpurun_out/prof_${tag}.log | cut -c1-200
