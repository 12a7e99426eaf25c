mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_stream" 2>&1 | tail -2
for v in 0 1; do echo "== SMALL_BRICK=$v"; CHEFSI_B200_SMALL_BRICK=$v timeout 600 python scripts/small_call_latency.py 2>&1 | grep -E "pageable"; done
for v in 0 1; do for c in Si8 Au_fcc211; do bash scripts/run_sparc_case.sh $c CHEFSI_B200_SMALL_BRICK=$v 2>&1 | sed "s/^/[$c brick=$v] /" | grep -E "walltime|AAR|Lanczos calls|Free energy per atom  "; done; done
