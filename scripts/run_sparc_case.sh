#!/bin/bash
# usage: scripts/run_sparc_case.sh <case> [env assignments]  -- run integration/_build/sparc_b200 on one staged case
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
work=$(mktemp -d)
cp -r $root/integration/_build/cases $work/
cd $work/cases/tests/$name/standard
start=$(date +%s%N)
env CHEFSI_B200_SHIM_VERBOSE=1 OMP_NUM_THREADS=1 "$@" $root/integration/_build/sparc_b200 -name $name > run.log 2> run.err || { tail -20 run.log run.err; exit 1; }
end=$(date +%s%N)
echo "== $name $* wall $(( (end - start) / 1000000 )) ms"
grep -E "Free energy per atom|Total number of SCF|Total walltime" $name.out
grep -E "Free energy per atom" $name.refout | sed 's/^/refout: /'
grep -F "[chefsi_b200 shim]" run.err || tail -3 run.err
