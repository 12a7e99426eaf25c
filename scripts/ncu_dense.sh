# usage: bash scripts/ncu_dense.sh <tag> [env assignments...]   -- one ncu --set full capture of the dense streaming stencil kernel
tag=$1; shift
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stream_dense" -s 4 -c 1 -o gpurun_out/prof_${tag} \
    python bench.py --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --no-nloc --e2e-cols 8 > gpurun_out/prof_${tag}.log 2>&1
tail -1 gpurun_out/prof_${tag}.log | cut -c1-300
