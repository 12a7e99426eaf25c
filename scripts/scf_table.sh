# CheFSI seconds per SCF for the four BASELINE.json test systems: SPARC + drop-in on the GPU vs the same executable with
# every call forwarded to the reference's own routines (CHEFSI_B200_DISABLE=1, one host core, np = 1)
for c in Si8 Si8_kpt BaTiO3 Au_fcc211; do
  for mode in gpu cpu; do
    if [ $mode = cpu ]; then extra="CHEFSI_B200_DISABLE=1"; else extra="CHEFSI_B200_X=0"; fi
    bash scripts/run_sparc_case.sh $c $extra 2>&1 | grep -v "bc: command" | sed "s/^/[$c $mode] /"
  done
done
