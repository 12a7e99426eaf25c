# usage: bash scripts/ncu_kernel.sh <tag> <kernel-regex> <skip> [env assignments...] -- one ncu --set full capture of one launch
# of the named kernel inside a 128-column bench run (never a bench value: numbers printed under ncu are discarded)
tag=$1; regex=$2; skip=$3; shift 3
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c 1 -o gpurun_out/prof_${tag} \
    python bench.py --ncol 128 --steps 1 --warmup 1 --skip-cpu-baseline --e2e-cols 8 > gpurun_out/prof_${tag}.log 2>&1
tail -1 gpurun_out/prof_${tag}.log | cut -c1-200
python profiles/ncu_summary.py gpurun_out/prof_${tag}.ncu-rep > gpurun_out/prof_${tag}.txt 2>&1
